#!/bin/bash
# Quick tuning loop on the box: parity first, then bench variants.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for mb in 3 4 5; do
  echo "== AIE_STEP_MINB=$mb"
  AIE_STEP_MINB=$mb timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_minb$mb.json 2> gpurun_out/bench_minb$mb.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_minb$mb.json"))
print("value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"]*1e3,1) for k, v in d["roofline"]["kernels"].items()})
PY
done

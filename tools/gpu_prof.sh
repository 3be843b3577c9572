#!/bin/bash
# ncu captures of the two kernels + source-level hot spots (run on the box)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
if [[ ${1:-test} == test ]]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"]*1e3,1) for k, v in d["roofline"]["kernels"].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aie_step_kernel -s 12 -c 1 -f -o gpurun_out/prof_step \
    python bench.py --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"

#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_$label.json"))
print("$label value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"]*1e3,1) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["kernels"]["aie_step_kernel"].get("unfused_ms"))
PY
}
run base X=1





run minb3 AIE_STEP_MINB=3
run minb5 AIE_STEP_MINB=5

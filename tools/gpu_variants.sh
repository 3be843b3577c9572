#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$label.json"))
    print("$label value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"]*1e3,1) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["kernels"]["aie_step_kernel"].get("unfused_ms"))
except Exception as ex:
    print("$label FAILED", ex, open("gpurun_out/bench_$label.err").read()[-300:])
PY
}
run m4 AIE_STEP_MINB=4
run m4_ps AIE_STEP_MINB=4 AIE_PHASE_SYNC=1
run m3_ps AIE_STEP_MINB=3 AIE_PHASE_SYNC=1
run m5_ps AIE_STEP_MINB=5 AIE_PHASE_SYNC=1
run w16_ps AIE_STEP_WPB=16 AIE_PHASE_SYNC=1
run w32_ps AIE_STEP_WPB=32 AIE_PHASE_SYNC=1
run w32 AIE_STEP_WPB=32

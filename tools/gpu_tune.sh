#!/bin/bash
# Tuning sweep on the box: launch-shape knobs of the fused step kernel (AIE_STEP_WPB, AIE_STEP_MINB, AIE_PHASE_SYNC, AIE_MW, AIE_SPLIT).
# usage: bash tools/gpu_tune.sh [phase|wpb|c5]...
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
B="--no-cpu-baseline --no-extra-workloads --e2e-steps 3 --steps 100 --warmup 10"
run() {  # label, env assignments..., -- bench args
  local label=$1; shift
  local envs=(); while [[ $1 != -- ]]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py $B "$@" > gpurun_out/tune_$label.json 2> gpurun_out/tune_$label.err
  python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/tune_%s.json" % sys.argv[1]))
    k = d["roofline"]["kernels"]["aie_step_kernel"]
    print("%-22s ms/step %.4f sustained %.4f  step kernel %.1f us (dyn %.1f obs %.1f) frac %.3f" % (
        sys.argv[1], d["ms_per_step"], d["sustained"]["ms_per_step"], k["ms"] * 1e3, k["unfused_ms"]["dynamics_only"] * 1e3,
        k["unfused_ms"]["observe_only"] * 1e3, d["roofline"]["frac"]))
except Exception as ex:
    print(sys.argv[1], "failed:", ex)
PY
}
WHAT="${*:-phase}"
if [[ $WHAT == *phase* ]]; then
  timeout 300 env AIE_PHASE_SYNC=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > gpurun_out/pytest_tune.log 2>&1; tail -2 gpurun_out/pytest_tune.log
  run c2_base X=1 -- --workload c2
  run c2_phase_sync AIE_PHASE_SYNC=1 -- --workload c2
  run c3_base X=1 -- --workload c3 --preroll 300
  run c3_phase_sync AIE_PHASE_SYNC=1 -- --workload c3 --preroll 300
fi
if [[ $WHAT == *wpb* ]]; then
  for w in 8 7 6; do run c2_wpb$w AIE_STEP_WPB=$w -- --workload c2; done
  for w in 8 7 6 5; do run c3_wpb$w AIE_STEP_WPB=$w -- --workload c3 --preroll 300; done
fi
if [[ $WHAT == *c5* ]]; then
  run c5_mw4_split0 AIE_MW=4 AIE_SPLIT=0 -- --workload c5 --steps 40 --preroll 100
  run c5_mw4_split1 AIE_MW=4 AIE_SPLIT=1 -- --workload c5 --steps 40 --preroll 100
  run c5_mw1_split1 AIE_MW=1 AIE_SPLIT=1 -- --workload c5 --steps 40 --preroll 100
fi

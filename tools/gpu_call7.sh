#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
B="--no-cpu-baseline --no-extra-workloads --e2e-steps 3 --steps 100 --warmup 10"
run() {
  local label=$1; shift
  local envs=(); while [[ $1 != -- ]]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py $B "$@" > gpurun_out/tune_$label.json 2> gpurun_out/tune_$label.err
  python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/tune_%s.json" % sys.argv[1]))
    k = d["roofline"]["kernels"]["aie_step_kernel"]
    print("%-22s ms/step %.4f sustained %.4f  step kernel %.1f us (dyn %.1f obs %.1f) frac %.3f reset: %s" % (
        sys.argv[1], d["ms_per_step"], d["sustained"]["ms_per_step"], k["ms"] * 1e3, k["unfused_ms"]["dynamics_only"] * 1e3,
        k["unfused_ms"]["observe_only"] * 1e3, d["roofline"]["frac"], d["config"].get("device_reset")))
except Exception as ex:
    print(sys.argv[1], "failed:", ex)
PY
}
run c2 X=1 -- --workload c2
run c5_mw4_split0 AIE_MW=4 AIE_SPLIT=0 -- --workload c5 --steps 60
run c5_mw4_split1 AIE_MW=4 AIE_SPLIT=1 -- --workload c5 --steps 60
run c5_mw1_split1 AIE_MW=1 AIE_SPLIT=1 -- --workload c5 --steps 60
run c3 X=1 -- --workload c3 --preroll 300

#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_device_reset.py -m gpu -x -q > gpurun_out/pytest_part.log 2>&1; tail -2 gpurun_out/pytest_part.log
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
    print("%-22s ms/step %.4f sustained %.4f  step kernel %.1f us (dyn %.1f obs %.1f) frac %.3f" % (
        sys.argv[1], d["ms_per_step"], d["sustained"]["ms_per_step"], k["ms"] * 1e3, k["unfused_ms"]["dynamics_only"] * 1e3,
        k["unfused_ms"]["observe_only"] * 1e3, d["roofline"]["frac"]))
except Exception as ex:
    print(sys.argv[1], "failed:", ex)
PY
}
run c3 AIE_VERBOSE=1 -- --workload c3 --preroll 300
grep "\[aie\]" gpurun_out/tune_c3.err | head -3
run c3_minb4 AIE_STEP_MINB=4 -- --workload c3 --preroll 300
run c3_wpb4 AIE_STEP_WPB=4 -- --workload c3 --preroll 300
run c3_wpb6 AIE_STEP_WPB=6 -- --workload c3 --preroll 300
run c2 X=1 -- --workload c2
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_run.py > gpurun_out/racecheck.log 2>&1; tail -6 gpurun_out/racecheck.log

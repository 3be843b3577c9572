#!/bin/bash
# Scratch harness for one-off B200 experiments: bench.py under different env knobs / env counts.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { # label, bench-args (quoted), env...
  label=$1; shift; bargs=$1; shift
  env AIE_VERBOSE=1 "$@" timeout 600 python bench.py --no-cpu-baseline --e2e-steps 3 $bargs > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$label.json"))
    print("$label value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"]*1e3,1) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["frac"], d["clocks"]["sm_mhz"])
except Exception as ex:
    print("$label FAILED", ex, open("gpurun_out/bench_$label.err").read()[-300:])
PY
}
run c4 "--workload c4 --steps 300 --warmup 20" X=1
run c2 "" X=1

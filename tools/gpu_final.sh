#!/bin/bash
# Last check of HEAD on the box: GPU tests, smoke, the default bench invocation (both arms), COVID kernel capture.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("default bench: value %.3e, e2e %.3e, frac %.3f, clocks %s, cpu_baseline %.3e" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["value"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aie_covid_step_kernel -s 150 -c 1 -f -o gpurun_out/prof_covid python bench.py --workload c4 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_covid.log 2>&1; echo "ncu covid rc=$?"

#!/bin/bash
# Runs on the B200 box under gpurun: parity tests, smoke, bench, ncu launch list + full captures, racecheck.
# Everything of interest lands in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
PH=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [[ $PH == all || $PH == *test* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
if [[ $PH == all || $PH == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
  tail -3 gpurun_out/smoke.log
fi
if [[ $PH == all || $PH == *bench* ]]; then
  timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_ref.json
fi
if [[ $PH == all || $PH == *c3* ]]; then
  timeout 900 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
  cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
fi
if [[ $PH == all || $PH == *c4* ]]; then
  timeout 900 python bench.py --workload c4 --steps 100 --warmup 10 --e2e-steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
  cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
fi
if [[ $PH == all || $PH == *c5* ]]; then
  timeout 900 python bench.py --workload c5 --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
  cat gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
fi
if [[ $PH == all || $PH == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 90 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:aie_step_kernel -s 8 -c 2 -f -o gpurun_out/prof_step \
      python bench.py --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:aie_observe_kernel -s 8 -c 2 -f -o gpurun_out/prof_observe \
      python bench.py --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_observe.log 2>&1; echo "ncu observe rc=$?"
fi
if [[ $PH == all || $PH == *race* ]]; then
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_run.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
  tail -8 gpurun_out/racecheck.log
  timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_run.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
  tail -5 gpurun_out/memcheck.log
fi
ls -la gpurun_out

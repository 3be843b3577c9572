#!/bin/bash
# reference-CUDA COVID test + e2e probe + c4 bench (run on the box)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m pytest tests/test_covid_ref_cuda.py tests/test_compact_transfer.py -m gpu -x -q -s > gpurun_out/pytest_new.log 2>&1; tail -5 gpurun_out/pytest_new.log
timeout 400 python tools/e2e_probe.py c2 > gpurun_out/e2e_probe_c2.log 2>&1; tail -14 gpurun_out/e2e_probe_c2.log
timeout 300 python bench.py --workload c4 --steps 540 --warmup 10 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python tools/summarize_bench.py gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
python -c "import json; print(json.load(open('gpurun_out/bench_c4.json')).get('vs_reference_cuda'))"

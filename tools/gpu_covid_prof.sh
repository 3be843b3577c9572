cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aie_covid_step_kernel -s 150 -c 1 -f -o gpurun_out/prof_covid python bench.py --workload c4 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_covid.log 2>&1; echo "rc=$?"

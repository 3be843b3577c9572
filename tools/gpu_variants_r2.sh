#!/bin/bash
# Round-2 opener (one gpurun call): for the default library and every prebuilt tuning variant (tools/build_variants.py,
# built in the CPU container; the .so files travel with the snapshot) - the golden / batch parity tests of the
# BASELINE shapes, then a c2 and a c3 bench line.  Prints one summary line per variant.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
V=ai_economist_b200/csrc/variants
for lib in default $(ls $V/*.so 2>/dev/null); do
  name=$(basename $lib .so)
  if [[ $lib == default ]]; then unset AIE_LIB_PATH; else export AIE_LIB_PATH=$PWD/$lib; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(c1_tutorial or c3_paper_tax or c5_small or full_size_c2) and not c3_full_size and not c5_full_size" > gpurun_out/pytest_$name.log 2>&1
  echo "$name parity rc=$? $(tail -1 gpurun_out/pytest_$name.log)"
  for w in c2 c3; do
    timeout 300 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_${name}_$w.json 2> gpurun_out/bench_${name}_$w.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${name}_$w.json"))
    print("$name $w value %.4e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"] * 1e3, 1) for k, v in d["roofline"]["kernels"].items()}, "frac %.3f" % d["roofline"]["frac"], d["clocks"]["sm_mhz"])
except Exception as ex:
    print("$name $w FAILED", ex)
PY
  done
done
# end-to-end leg: plain D2H copies vs the compacted transfer (same host bytes), default library
unset AIE_LIB_PATH
timeout 200 python -m pytest tests/test_compact_transfer.py -m gpu -x -q > gpurun_out/pytest_compact.log 2>&1; echo "compact parity rc=$? $(tail -1 gpurun_out/pytest_compact.log)"
for mode in plain compact; do
  for th in 0 16 64; do
    [[ $mode == plain && $th != 0 ]] && continue
    timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 20 --e2e-mode $mode --e2e-threads $th > gpurun_out/bench_e2e_${mode}_$th.json 2> gpurun_out/bench_e2e_${mode}_$th.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_e2e_${mode}_$th.json"))
    print("e2e $mode threads=$th: %.4e agent-env-steps/s, d2h %.1f MB/step" % (d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"] / 1e6))
except Exception as ex:
    print("e2e $mode $th FAILED", ex)
PY
  done
done
# COVID (c4): history scan vs the persistent change list
timeout 300 python -m pytest tests/test_covid.py -m gpu -x -q > gpurun_out/pytest_covid.log 2>&1; echo "covid parity rc=$? $(tail -1 gpurun_out/pytest_covid.log)"
for cl in 0 1; do
  AIE_COVID_CHANGE_LIST=$cl timeout 300 python bench.py --workload c4 --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_c4_cl$cl.json 2> gpurun_out/bench_c4_cl$cl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c4_cl$cl.json"))
    print("c4 change_list=$cl value %.4e ms/step %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
except Exception as ex:
    print("c4 change_list=$cl FAILED", ex)
PY
done
# source-level captures of the step kernel on the two heavier gather-trade-build workloads (c3: paper config with taxes,
# c5: 64 agents / deep book), read offline with tools/ncu_by_line.py: round 1 only profiled c2 by line
for w in c3 c5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:aie_step_kernel -s 8 -c 1 -f -o gpurun_out/prof_step_$w \
      python bench.py --workload $w --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_step_$w.log 2>&1; echo "ncu $w rc=$?"
done

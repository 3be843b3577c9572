#!/bin/bash
# e2e leg of the bench under combinations of the transfer knobs.  Each argument after the workload is one combination of
# VAR=value settings joined by commas, e.g.  bash tools/gpu_e2e_chunks.sh c2 AIE_E2E_CHUNKS=4,AIE_E2E_ITEM_ENVS=16 AIE_E2E_CHUNKS=2
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
[[ -n "${SKIP_TESTS:-}" ]] || timeout 600 python -m pytest tests/test_compact_transfer.py -m gpu -x -q > gpurun_out/pytest_compact.log 2>&1; [[ -n "${SKIP_TESTS:-}" ]] || tail -2 gpurun_out/pytest_compact.log
w=$1; shift
i=0
for combo in "$@"; do
  i=$((i+1))
  env $(echo "$combo" | tr ',' ' ') timeout 300 python bench.py --workload $w --no-cpu-baseline --no-extra-workloads --e2e-steps 20 --steps 50 --warmup 10 ${BENCH_ARGS:-} \
      > gpurun_out/e2e_${w}_$i.json 2> gpurun_out/e2e_${w}_$i.err
  python - $w $i "$combo" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/e2e_%s_%s.json" % (sys.argv[1], sys.argv[2])))
    e = d["e2e"]; t = e["last_call_timing_ms"]
    print("%s e2e %.4e  ms/step mean %.3f min %.3f max %.3f | first slice dev %.3f last %.3f host-last %.3f expanded %.3f wait %.1f busy %.1f staging node %s | %s" % (
        sys.argv[1], e["value"], e["ms_per_step"]["mean"], e["ms_per_step"]["min"], e["ms_per_step"]["max"], t["first_slice_dev"],
        t["last_slice_dev"], t["last_slice"], t["expanded"], t["wait_sum"], t["busy_sum"], t.get("staging_node"), sys.argv[3]))
except Exception as ex:
    print(sys.argv[1], sys.argv[3], "failed:", ex)
PY
done

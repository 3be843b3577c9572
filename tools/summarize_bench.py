"""One-screen summary of a bench.py JSON line (for the gpurun tail)."""
import json
import sys

try:
    d = json.load(open(sys.argv[1]))
except Exception as ex:
    print("no bench line:", ex)
    sys.exit(0)
r = d.get("roofline", {})
print("value %.4e  ms/step %.4f  sustained %.4f ms (%s reps)  e2e %.4e [%s]  launches %s  clocks %s" % (
    d["value"], d["ms_per_step"], d.get("sustained", {}).get("ms_per_step", float("nan")), d.get("sustained", {}).get("repeats"),
    d["e2e"]["value"], d["e2e"].get("mode", ""), d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
print("  roofline %s frac %.3f (survey bytes %.3f) traffic %s; kernels %s" % (
    r.get("kernel"), r.get("frac", float("nan")), (r.get("frac_on_survey_8d_bytes") or {}).get("frac", float("nan")), r.get("traffic"),
    {k: round(v["ms"] * 1e3, 1) for k, v in r.get("kernels", {}).items()}))
un = r.get("kernels", {}).get("aie_step_kernel", {}).get("unfused_ms")
if un:
    print("  unfused: dynamics %.1f us, observe %.1f us" % (un["dynamics_only"] * 1e3, un["observe_only"] * 1e3))
if "cpu_baseline" in d:
    print("  cpu_baseline %.4e on %s cores (%s)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"]))
for k, w in (d.get("workloads") or {}).items():
    if "error" in w:
        print("  %s: ERROR %s" % (k, w["error"]))
    else:
        print("  %s: value %.4e ms/step %.4f (sustained %.4f) frac %.3f e2e %.4e kernels %s" % (
            k, w["value"], w["ms_per_step"], w["sustained"]["ms_per_step"], w["roofline"]["frac"], w["e2e"]["value"],
            {kk: round(v * 1e3, 1) for kk, v in w["roofline"]["kernel_ms"].items()}))
        if "vs_reference_cuda" in w:
            v = w["vs_reference_cuda"]
            print("      vs reference CUDA:", v if "error" in v else "ref %.4f ms/step (%d launches) -> x%.2f whole step, x%.2f step kernel" % (
                v["ms_per_step"], v["launches_per_step"], v["speedup_whole_step"], v["speedup_step_kernel_only"]))

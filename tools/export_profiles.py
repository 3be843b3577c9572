"""gpurun_out/ of `bash tools/gpu_round.sh test,smoke,bench,ncu,ncx` -> committed artefacts under profiles/<tag>_*:
bench lines, logs, the ncu launch list, and for every `ncu --set full` report the raw metrics of its first launch (CSV: the
file bench.py reads `roofline.traffic` from) plus the source page aggregated per line / function.   python tools/export_profiles.py r02z"""
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
ENVS = {"c2": 8192, "c3": 8192, "c4": 4096, "c5": 2048}

for src, dst in [("bench.json", "bench_c2_all_workloads.json"), ("bench_ref.json", "bench_reference_arm.json"),
                 ("pytest_gpu.log", "pytest_gpu.log"), ("smoke.log", "smoke.log"), ("launches.csv", "launches_c2.csv"),
                 ("racecheck.log", "racecheck.log"), ("memcheck.log", "memcheck.log"), ("bench.err", "bench.err")]:
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, "%s_%s" % (tag, dst)))
for w in ("c2", "c3", "c4", "c5"):
    raw_csv, src_csv = os.path.join(OUT, "prof_step_%s_raw.csv" % w), os.path.join(OUT, "prof_step_%s_src.csv" % w)
    if not os.path.exists(raw_csv):
        continue
    kernel = "aie_covid_step_kernel" if w == "c4" else "aie_step_kernel"
    rows = list(csv.reader(open(raw_csv)))
    name = "%s_ncu_full_%s%s_raw.csv" % (tag, kernel, "" if w == "c2" else "_" + w)
    with open(os.path.join(PROF, name), "w", newline="") as f:
        csv.writer(f).writerows(rows[:3])
    d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"]
    print(w, {k: (d.get(k), u.get(k)) for k in keys})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), src_csv, str(ENVS[w]), "40"],
                         capture_output=True, text=True, cwd=ROOT).stdout
    open(os.path.join(PROF, "%s_step_kernel_by_line_%s.txt" % (tag, w)), "w").write(out)

"""Aggregate an ncu source page (cuda,sass view) per CUDA source line and per enclosing function.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_by_line.py src.csv [n_envs] [top]

Prints executed warp-instructions per env, stall samples and the dominant stall reasons."""
import csv, sys, re, collections

path = sys.argv[1]
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path, newline="")))
hdr = None
cur_file, cur_line, cur_src = None, None, ""
per_line = collections.OrderedDict()
stall_cols = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        i_samp = hdr.index("# Samples"); i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None: continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0])); cur_src = r[1].strip()
        per_line.setdefault(cur_line, {"src": cur_src, "samp": 0, "inst": 0, "stalls": collections.Counter(), "sass": 0})
        continue
    d = per_line[cur_line]
    try:
        d["samp"] += int(r[i_samp]); d["inst"] += int(r[i_inst]); d["sass"] += 1
        for i, h in stall_cols:
            v = int(r[i]) if r[i] not in ("", "-") else 0
            if v: d["stalls"][h[6:]] += v
    except (ValueError, IndexError):
        pass
tot_s = sum(d["samp"] for d in per_line.values()); tot_i = sum(d["inst"] for d in per_line.values())
print("total samples %d, warp-instructions %d (%.0f per env)" % (tot_s, tot_i, tot_i / n_envs))
allst = collections.Counter()
for d in per_line.values(): allst.update(d["stalls"])
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(allst.values()))) for k, v in allst.most_common(10)))
print("\n== top lines by samples")
for (f, ln), d in sorted(per_line.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = ", ".join("%s %d" % kv for kv in d["stalls"].most_common(3))
    print("%5.1f%% samp %6d  inst/env %6.1f  sass %4d  %s:%d  %s   [%s]" % (100.0 * d["samp"] / max(1, tot_s), d["samp"], d["inst"] / n_envs, d["sass"], f, ln, d["src"][:70], st))
# per-function: nearest preceding line in the same file that looks like a function header
func_of = {}
files = collections.defaultdict(list)
for (f, ln) in per_line: files[f].append(ln)
src_cache = {}
def func_for(f, ln):
    import os
    p = None
    for root in ("ai_economist_b200/csrc", "."):
        cand = os.path.join(root, f)
        if os.path.exists(cand): p = cand; break
    if p is None: return f
    if p not in src_cache: src_cache[p] = open(p).read().splitlines()
    L = src_cache[p]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.match(r"^(?:template.*>\s*)?(?:AIE_DEV(?:_NOINLINE)?|__global__|__device__|static|inline)\b.*?(\w+)\s*\(", L[i])
        if m and not L[i].startswith(" "): return m.group(1)
    return f
agg = collections.defaultdict(lambda: {"samp": 0, "inst": 0, "stalls": collections.Counter()})
for (f, ln), d in per_line.items():
    fn = func_for(f, ln)
    agg[fn]["samp"] += d["samp"]; agg[fn]["inst"] += d["inst"]; agg[fn]["stalls"].update(d["stalls"])
print("\n== by function (source attribution after inlining)")
for fn, d in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = ", ".join("%s %d" % kv for kv in d["stalls"].most_common(4))
    print("%5.1f%% samp %6d  inst/env %7.1f  %-24s [%s]" % (100.0 * d["samp"] / max(1, tot_s), d["samp"], d["inst"] / n_envs, fn, st))

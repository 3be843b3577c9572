"""Instruction count per kernel of the built library (cuobjdump -sass), plus a few mnemonic counts."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "ai_economist_b200/csrc/libaie_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, cnt, mn = None, collections.Counter(), collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: name = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        cnt[name] += 1
        mn[name][m.group(1).split(".")[0]] += 1
for n, c in sorted(cnt.items(), key=lambda x: x[1]):
    top = ", ".join("%s %d" % kv for kv in mn[n].most_common(8))
    print("%6d  %s\n        %s" % (c, n, top))

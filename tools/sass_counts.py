"""Per-kernel SASS opcode counts of libpmk_b200.so (cuobjdump -sass): the evidence that the FP64 tensor path (DMMA.8x8x4),
1-D TMA bulk copies (UBLKCP) and mbarriers (SYNCS) are what the shipped kernels execute.  python tools/sass_counts.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "patchmixturekriging_b200", "libpmk_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "ATOM", "RED", "UTCQMMA", "UTCHMMA", "HMMA"]
cur, counts, arch = None, collections.OrderedDict(), ""
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
        if op == "DMMA":
            counts[cur]["DMMA" + (m.group(2) or "")] += 1
print(f"# {os.path.basename(so)}  arch {arch}  ({len(counts)} kernels)   columns: instructions, then opcode counts (prefix match)")
tot = collections.Counter()
for fn, c in counts.items():
    name = demangle(fn)
    name = re.sub(r"\(.*$", "", name)
    cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    shapes = " ".join(f"{k}={v}" for k, v in c.items() if k.startswith("DMMA."))
    print(f"{name[:110]:<70s} {c['_total']:>6d}  {cols}  {shapes}")
    tot.update(c)
print("# total: " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))

"""Static census of the tensor-core / TMA / TMEM instructions in the SHIPPED library: `cuobjdump -sass libdlv3p.so`,
counted per kernel.  Evidence that the GEMM / implicit-convolution / depthwise kernels are tcgen05 + TMEM + TMA code
(UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce,
UTCBAR = tcgen05.commit, SYNCS = mbarrier) and that no legacy HMMA (mma.sync) path exists.
Usage: python scripts/sass_census.py [path/to/libdlv3p.so] > profiles/r2_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deeplabv3plus_keras_b200", "libdlv3p.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA",
         "IMMA", "FFMA2", "HFMA2", "ATOMS", "RED", "ATOMG", "LDGSTS"]
per = collections.OrderedDict()
cur = None
archs = set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        archs.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1


def demangle(names):
    try:
        r = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, r))
    except Exception:
        return {n: n for n in names}


dm = demangle(list(per))
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; architectures: {', '.join(sorted(archs))})\n")
print(f"{len(per)} kernels, {tot['_total']} instructions.  Totals: " +
      ", ".join(f"{w} {tot[w]}" for w in WATCH if tot[w] or w in ("HMMA", "IMMA")) + "\n")
print("| kernel | instr | UTCHMMA | UTCBAR | LDTM | UTMALDG | UTMASTG | UTMAREDG | SYNCS | FFMA2 |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
rows = [(n, c) for n, c in per.items() if any(c[w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"))]
rows.sort(key=lambda r: -(r[1]["UTCHMMA"] * 1000 + r[1]["UTMALDG"]))
for n, c in rows:
    name = re.sub(r"\((int|bool|unsigned int)\)", "", dm[n]).split("(")[0].replace("dlv3p::", "").replace("void ", "")
    print(f"| `{name[:90]}` | {c['_total']} | {c['UTCHMMA']} | {c['UTCBAR']} | {c['LDTM']} | {c['UTMALDG']} | "
          f"{c['UTMASTG']} | {c['UTMAREDG']} | {c['SYNCS']} | {c['FFMA2']} |")
print(f"\n{len(per) - len(rows)} further kernels (element-wise / reductions / SIMT fp32 parity path) use none of the above.")

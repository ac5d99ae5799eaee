"""Opcode histogram of an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source sass`): executed
warp-instructions and stall samples per SASS opcode, per kernel.  Usage: python scripts/sass_hist.py file.csv [N]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
ops, samp, tot, name = collections.Counter(), collections.Counter(), 0, ""


def flush():
    if tot:
        print(f"== {name[:100]}  executed warp-instructions: {tot}")
        for op, n in ops.most_common(top):
            print(f"  {op:12s} {n:11d} {100 * n / tot:5.1f}%   stall samples {samp[op]}")


for r in rows:
    if r and r[0] == "Kernel Name":
        flush()
        ops, samp, tot, name = collections.Counter(), collections.Counter(), 0, r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ia:
        continue
    n, s = int(r[ia] or 0), int(r[ist] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += n
    samp[op] += s
    tot += n
flush()

"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>`) into a per-kernel
table: launches, total device time and share.  Per-launch times under ncu are cold-cache and serialised, so the
SHARES are what is comparable with the CUDA-event numbers of bench.py, not the absolutes (B200_PROFILING.md).

    python scripts/summarize_launches.py gpurun_out/launches_r1.csv [--last N] > profiles/r1_launches.md

--last N keeps only the final N launches of the capture (e.g. the last full training step).
"""
import argparse
import collections
import csv
import re
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--last", type=int, default=0)
ap.add_argument("--first", type=int, default=0)
args = ap.parse_args()

rows = []
with open(args.csv, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    ns = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit", "ns") in ("us", "usecond"):
        ns *= 1e3
    rows.append((r["Kernel Name"], ns, r["Grid Size"], r["Block Size"]))
if args.last:
    rows = rows[-args.last:]
if args.first:
    rows = rows[:args.first]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("dlv3p::", "").replace("<unnamed>::", "")
    return name[:90]


agg = collections.OrderedDict()
for name, ns, grid, block in rows:
    a = agg.setdefault(short(name), [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"launches: {len(rows)}  total device time: {total / 1e6:.3f} ms (cold-cache, serialised under ncu)\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |")

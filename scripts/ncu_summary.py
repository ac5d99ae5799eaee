"""Summarise an .ncu-rep (read here with `ncu -i`): per captured launch the headline counters the roofline argument
rests on.  Usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X.md ; also prints a JSON line per
launch on stderr-free stdout tail for profiles/ncu_traffic.json."""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem % of peak (active)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_dim_x", "cluster x"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# ncu --set full summary of `{rep}` (counters read with `ncu -i ... --page raw --csv`)\n")
traffic = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    print(f"## `{name[:110]}`\n")
    print("| counter | value | unit |\n|---|---:|---|")
    vals = {}
    for key, label in want:
        if key in idx:
            print(f"| {label} (`{key}`) | {r[idx[key]]} | {units[idx[key]]} |")
            vals[key] = (r[idx[key]], units[idx[key]])
    print()

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if "dram__bytes_read.sum" in vals:
        t = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
        traffic.setdefault(name.split("(")[0], []).append(t)
print("<!-- traffic " + json.dumps({k: sum(v) / len(v) for k, v in traffic.items()}) + " -->")

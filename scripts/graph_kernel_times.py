"""True in-graph kernel durations of the replayed training step (CUPTI activity records through torch.profiler): the
instrumented eager pass of bench.py brackets every launch with events and overstates the 15-30 us kernels.
Prints per-kernel-name totals per step and writes profiles-style markdown to stdout."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import warnings

import bench as B
from deeplabv3plus_keras_b200 import SemanticSegmentation, keras
from deeplabv3plus_keras_b200.trainer import Trainer

STEPS = 3
conf = B.make_conf("bfloat16")
keras.reset_uids()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    ss = SemanticSegmentation(conf)
B.he_init(ss.model)
tr = Trainer(ss.model, 16)
x, y = B.synthetic(conf, 16, tr.plan.out_shape[1:3], 1024)
xs, ys = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
tr.stage_inputs(xs, ys)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        tr.step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t0, t1 = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    name = ev.name
    dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    agg[name][0] += 1
    agg[name][1] += dur
total = sum(v[1] for v in agg.values()) / STEPS
print(f"# In-graph kernel time of the replayed training step (CUPTI via torch.profiler, {STEPS} steps)\n")
print(f"sum of kernel durations per step: {total / 1e3:.3f} ms (kernels on the side stream overlap the main chain)\n")
print("| kernel | launches/step | us/step | avg us | share |")
print("|---|---:|---:|---:|---:|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    short = name.replace("dlv3p::", "").split("(")[0][:90]
    print(f"| `{short}` | {n / STEPS:.1f} | {us / STEPS:.1f} | {us / n:.1f} | {100 * us / STEPS / total:.1f}% |")

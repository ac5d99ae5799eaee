"""Probe: does torch symmetric memory (CUDA VMM peer mappings over NVLink) work on this box, and what do the library
all-reduces cost at the gradient-arena size?  torchrun --nproc-per-node N scripts/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n = 16_109_824
group = dist.group.WORLD
try:
    t = symm.empty(n, dtype=torch.float32, device="cuda")
    hdl = symm.rendezvous(t, group.group_name)
    if rank == 0:
        print("symm ok: world", hdl.world_size, "rank", hdl.rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4],
              "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:2], "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0),
              "signal_pad_size", getattr(hdl, "signal_pad_size", None), flush=True)
except Exception as e:      # noqa: BLE001
    print(rank, "symm FAILED:", repr(e)[:400], flush=True)
    dist.destroy_process_group()
    raise SystemExit(0)


def timed(fn, name, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(f"{name}: {ms * 1e3:.1f} us for {nbytes / 1e6:.0f} MB", flush=True)


g = torch.randn(n, device="cuda")
timed(lambda: dist.all_reduce(g), "nccl all_reduce fp32", n * 4)
g16 = torch.randn(n, device="cuda").bfloat16()
timed(lambda: dist.all_reduce(g16), "nccl all_reduce bf16", n * 2)
for opname in ("two_shot_all_reduce_", "one_shot_all_reduce", "multimem_all_reduce_"):
    try:
        op = getattr(torch.ops.symm_mem, opname)
        t.copy_(g)
        if opname == "one_shot_all_reduce":
            fn = lambda: op(t, "sum", group.group_name)
        else:
            fn = lambda: op(t, "sum", group.group_name)
        timed(fn, f"symm_mem.{opname} fp32", n * 4)
    except Exception as e:      # noqa: BLE001
        if rank == 0:
            print(opname, "failed:", repr(e)[:300], flush=True)
try:
    t16 = symm.empty(n, dtype=torch.bfloat16, device="cuda")
    symm.rendezvous(t16, group.group_name)
    for opname in ("two_shot_all_reduce_", "multimem_all_reduce_"):
        op = getattr(torch.ops.symm_mem, opname)
        timed(lambda: op(t16, "sum", group.group_name), f"symm_mem.{opname} bf16", n * 2)
except Exception as e:      # noqa: BLE001
    if rank == 0:
        print("bf16 symm failed:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()

"""Data-parallel plumbing (new here — the reference has no multi-GPU path).

The batch axis shards across ranks with no data-path collective (every image is independent through the graph and
BatchNormalization statistics stay per replica, as in a single-GPU Keras run); the only exchange is one sum
all-reduce of the flat fp32 gradient arena per step, cut into contiguous buckets so that NCCL can pipeline them.
The mean over replicas (each replica's loss is the mean over its own shard) is applied as `grad_scale = 1/world`
inside the fused Adam kernel rather than as a separate pass over the arena.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of the global batch owned by `rank`; requires an even split (Keras DP semantics)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not divide over {world} replicas")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def bucket_ranges(n: int, buckets: int, align: int = 8) -> List[Tuple[int, int]]:
    """Cut [0, n) into `buckets` contiguous ranges whose boundaries are multiples of `align` elements."""
    buckets = max(1, min(buckets, n // align if n >= align else 1))
    step = -(-n // buckets)
    step = -(-step // align) * align
    out, lo = [], 0
    while lo < n:
        hi = min(lo + step, n)
        out.append((lo, hi))
        lo = hi
    return out


def allreduce_gradients(g: torch.Tensor, n: int, group=None, buckets: int = 4) -> None:
    """In-place SUM all-reduce of g[:n] over the group (asynchronous per bucket, then waited)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    works = [dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)
             for lo, hi in bucket_ranges(n, buckets)]
    for w in works:
        w.wait()


def allreduce_ranges(g: torch.Tensor, ranges, group=None) -> None:
    """In-place SUM all-reduce of the given [lo, hi) slices of the flat gradient arena, enqueued on the current
    stream (the trainer's communication stream): the slices are the arena prefixes that became final during the
    backward segment just launched, so the exchange overlaps the rest of backward."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for lo, hi in ranges:
        if hi > lo:
            dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=group)


def exchange_schedule(plan, n_buckets: int):
    """How the Trainer overlaps the gradient exchange with backward: the backward launch list is cut into `n_buckets`
    contiguous segments; after segment i the arena slices that became FINAL during it (the gradient arena is laid out
    in backward order, Plan.final_prefixes) are all-reduced while segment i+1 runs.  Returns (cuts, ranges):
    cuts[i]..cuts[i+1] = launch range of segment i, ranges[i] = [(lo, hi), ...] to exchange after it; the last entry
    completes both arena regions, so the union of all ranges is exactly [0, n_train)."""
    P = plan.params
    n = len(plan.bwd)
    cuts = [round(i * n / n_buckets) for i in range(n_buckets + 1)]
    ranges, pa, pb = [], 0, P.n_reg
    for i, c in enumerate(cuts[1:]):
        a, b = (P.n_reg, P.n_train) if i == n_buckets - 1 else plan.final_prefixes(c)
        ranges.append([(pa, a), (pb, b)])
        pa, pb = a, b
    return cuts, ranges


def run_step_with_exchange(plan, n_buckets: int, group=None) -> None:
    """Reference (synchronous) execution of the Trainer's data-parallel step on the batch resident in `plan`: head
    (zero, forward, loss), then per segment: launch it, all-reduce what it finished.  The Trainer does the same with
    CUDA graphs per segment and the all-reduces on a communication stream."""
    cuts, ranges = exchange_schedule(plan, n_buckets)
    plan.finalize()
    plan.zero_grads()
    plan.forward()
    plan.loss_forward_backward()
    for (a, b), rg in zip(zip(cuts[:-1], cuts[1:]), ranges):
        plan.run_bwd_range(a, b)
        allreduce_ranges(plan.params.g, rg, group)

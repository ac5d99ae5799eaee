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


class PeerExchange:
    """SUM all-reduce of slices of the flat gradient arena WITHOUT SM-resident collective kernels: a reduce-scatter
    by peer-to-peer pushes into NVLink-mapped (symmetric-memory) staging rows, a local reduction of this rank's slice,
    and an all-gather by peer-to-peer pulls — the transfers are `cudaMemcpyAsync` between peer-mapped buffers (copy
    engines over NVLink / NVSwitch), the only kernels are three one-CTA signal barriers and one small row sum.

    Why: every hot kernel of the training step is a persistent one-CTA-per-SM launch with a static tile stride.  An NCCL
    all-reduce running beside backward puts its CTAs on SMs, the displaced CTAs of the compute kernel run after the
    others and that kernel's tail doubles: overlapping hid a quarter of the exchange (profiles/r2_scaling.md).  Copy
    engines take no SM; what is left beside backward is ~120 MB of extra HBM traffic per step.

    Every slice is reduced by exactly ONE rank in a fixed order (rank 0 .. world-1) and then copied to all the others, so
    the exchanged gradient is bit-identical on every rank, as after an NCCL all-reduce.

    Layout of the symmetric buffer (fp32, per rank): `world` staging rows of `cap` elements (row r = what rank r pushed
    for MY slices), then one row of `cap` elements with my reduced slices (what the peers pull)."""

    def __init__(self, n: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.cap = -(-n // self.world) + 64                      # two ranges per exchange, each chunk rounded up to 8
        self.buf = symm.empty((self.world + 1) * self.cap, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        self.rows = self.buf[:self.world * self.cap].view(self.world, self.cap)
        self.red = self.buf[self.world * self.cap:]

    def _slices(self, ranges):
        out, off = [], 0
        for lo, hi in ranges:
            L = hi - lo
            if L <= 0:
                continue
            ch = (-(-L // self.world) + 7) // 8 * 8
            out.append((lo, lo + L, ch, off))
            off += ch
        assert off <= self.cap, (off, self.cap)
        return out

    def all_reduce_(self, g: torch.Tensor, ranges) -> None:
        """In place on the current stream: g[lo:hi] <- sum over ranks, for every (lo, hi) of `ranges`."""
        segs = self._slices(ranges)
        if not segs:
            return
        # (no barrier in front: a peer pushes into my rows only after it has passed the closing barrier of the previous
        # exchange, which I entered after my reduction had read them)
        self._push(g, segs)
        self.hdl.barrier()                               # every push into my rows has landed
        self._reduce(g, segs)
        self.hdl.barrier()                               # every rank's reduced slices are in its `red` row
        self._pull(g, segs)
        self.hdl.barrier()                               # nobody still reads my `red` row when the next exchange starts

    def _push(self, g, segs):
        """reduce-scatter, push: my contribution to rank r's slices into row `me` of rank r's staging rows"""
        W, me, cap = self.world, self.rank, self.cap
        for step in range(W):
            r = (me + step) % W
            dst = self.hdl.get_buffer(r, (cap,), torch.float32, me * cap)
            for lo, end, ch, o in segs:
                a, b = lo + r * ch, min(lo + (r + 1) * ch, end)
                if b > a:
                    dst[o:o + b - a].copy_(g[a:b])

    def _reduce(self, g, segs):
        """my slices: sum of the staging rows in rank order -> `red` row and my own arena"""
        me = self.rank
        for lo, end, ch, o in segs:
            a, b = lo + me * ch, min(lo + (me + 1) * ch, end)
            if b > a:
                torch.sum(self.rows[:, o:o + b - a], dim=0, out=self.red[o:o + b - a])
                g[a:b].copy_(self.red[o:o + b - a])

    def _pull(self, g, segs):
        """all-gather, pull: every other rank's reduced slices out of its `red` row"""
        W, me, cap = self.world, self.rank, self.cap
        for step in range(1, W):
            r = (me + step) % W
            src = self.hdl.get_buffer(r, (cap,), torch.float32, W * cap)
            for lo, end, ch, o in segs:
                a, b = lo + r * ch, min(lo + (r + 1) * ch, end)
                if b > a:
                    g[a:b].copy_(src[o:o + b - a])


class GatherExchange:
    """SUM all-reduce of the gradient arena with NO kernel beside backward: every rank PUSHES each arena slice, as soon as
    the backward segment that completed it has run, into row `rank` of every peer's NVLink-mapped staging buffer
    (`cudaMemcpyAsync` on the copy engines, its own row included); after backward one signal barrier and ONE pass
    `g = sum_r rows[r]` in rank order (bit-identical on every rank) replace the all-reduce.  An all-gather moves
    (world-1) x the bytes of a reduce-scatter + all-gather, but NVLink has them to spare behind 5 ms of backward (66 MB x 7
    at N=8), and nothing SM-resident runs next to the persistent compute kernels — whose CTAs an overlapped NCCL kernel,
    or PeerExchange's row sum, displaces for as long as it runs.  Exposed per step: the barrier and the row sum
    (world x 66 MB read at HBM speed: 25 us at N=2, 100 us at N=8).

    Staging buffer (symmetric memory, fp32): [world, n] per rank."""

    def __init__(self, n: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = n
        self.buf = symm.empty(self.world * n, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        self.rows = self.buf.view(self.world, n)
        self._mine = [self.hdl.get_buffer(r, (n,), torch.float32, self.rank * n) for r in range(self.world)]

    def push_(self, g: torch.Tensor, ranges) -> None:
        """Current stream: my g[lo:hi] -> row `rank` of every rank's staging buffer (peers first, staggered)."""
        for step in range(1, self.world + 1):
            dst = self._mine[(self.rank + step) % self.world]
            for lo, hi in ranges:
                if hi > lo:
                    dst[lo:hi].copy_(g[lo:hi])

    def finish_(self, g: torch.Tensor) -> None:
        """Current stream, after the last push_: wait until every rank's pushes have landed, then g[:n] = sum of the rows;
        a closing barrier keeps the peers' next pushes out of the rows until this sum has read them."""
        self.hdl.barrier()
        torch.sum(self.rows, dim=0, out=g[:self.n])
        self.hdl.barrier()

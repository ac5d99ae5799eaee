"""Training-step driver: CUDA-graph replay of the lowered forward+backward, data-parallel gradient exchange.

One process per GPU.  The reference has no multi-GPU path at all (`multi_gpu`/`num_gpus` in conf.json:6-7 feed dead
code, ss.py:1222-1223); data parallelism over the batch axis is new here: weights, Adam moments and
BatchNormalization moving statistics are replicated, BN batch statistics stay per replica (the reference has no
cross-replica BN), each replica's loss is the mean over its shard, so the exchanged gradient is the mean over
replicas — one NCCL all-reduce over NVLink per bucket of the flat fp32 gradient arena, issued on a side stream as
soon as the backward segment that produced the bucket has been launched, overlapping the rest of backward.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import dp, ops
from .engine import Plan


class Trainer:
    def __init__(self, model, batch_size: int, dtype: Optional[str] = None, use_graph: bool = True,
                 buckets: int = 4, process_group=None, fused_tail: bool = True, overlap_wgrad: bool = True,
                 exchange: str = "overlap", grad_dtype: str = "float32", cut_events: bool = False):
        """`exchange` (world > 1): "overlap" = the arena prefixes each backward segment finished are all-reduced on a
        communication stream while the next segment runs (`buckets` segments); "tail" = ONE all-reduce of the whole arena
        after backward on the training stream.  Every hot kernel here is a persistent one-CTA-per-SM launch, so NCCL's
        CTAs displace CTAs of whatever runs beside them and that kernel's tail doubles: overlapping hides nothing
        beyond what it costs, and the tail exchange of a bf16 copy of the gradients (`grad_dtype="bfloat16"`, half
        the bytes, summed in fp32 per pair by NCCL's bf16 reduction) is the cheaper one (profiles/r2_scaling.md).
        "peer" = the overlap schedule with dp.PeerExchange instead of NCCL: peer-to-peer copies over NVLink-mapped
        symmetric memory on the copy engines (no collective CTAs beside the persistent compute kernels)."""
        if exchange not in ("overlap", "tail", "peer", "gather", "none") or grad_dtype not in ("float32", "bfloat16"):
            raise ValueError("exchange: 'overlap' | 'tail' | 'peer' | 'gather' | 'none'; grad_dtype: 'float32' | 'bfloat16'")
        # "gather" = dp.GatherExchange: copy-engine pushes of every finished slice to all peers behind backward, one barrier
        # and one row sum after it — nothing SM-resident beside the persistent compute kernels
        # ("none": independent replicas, NO gradient exchange — a measurement aid only: what N ranks stepping side by side
        # cost by themselves, i.e. the slowest GPU of the box and the max-over-ranks of the timing, before any collective)
        if grad_dtype == "bfloat16" and exchange != "tail":
            raise ValueError("grad_dtype='bfloat16' needs exchange='tail'")
        self.exchange, self.grad_dtype = exchange, grad_dtype
        # cut_events (world > 1, overlap / peer, CUDA graphs; off by default): ONE graph for the whole step as on one GPU —
        # no per-segment graph launches, no side-stream join at the cuts; the segment ends are marked inside it by EXTERNAL
        # events (event-record nodes) on the training and the filter-gradient streams, and the communication stream, outside
        # the graph, waits for them before it exchanges the slices that segment completed.  Correct (`bench.py --check
        # --cut-events`) but no faster than one graph per segment (8.34 vs 8.29-8.32 ms at N=2; 8.13 on one GPU or with
        # two replicas that never exchange): the cost of the data-parallel step is neither the graph cuts nor — the peer
        # exchange runs on copy engines and costs the same — NCCL's CTAs; it was not located this round
        self.cut_events = cut_events
        # (measured at N=2: NCCL kernels and copy engines cost the same 0.18 ms per step — 8.32 / 8.31 ms against 8.13 on one
        # GPU — so what the data-parallel step pays is the segmentation itself: a graph launch and a side-stream join per
        # segment.  Capturing the whole step with the exchanges forked inside ONE graph was tried and dropped: slower
        # (8.47 / 8.42 ms) and the replayed exchange did not reproduce the eager gradients in `bench.py --check`.)
        self.model = model
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        kw = {"dtype": dtype} if dtype else {}
        if self.world > 1:
            kw["dropout_seed"] = 1024 + 104729 * self.rank     # independent dropout masks per replica
        self.plan: Plan = model.plan(batch_size, training=True, fused_tail=fused_tail, **kw)
        p = self.plan
        loss = model.loss
        p.set_loss(loss.pos_weights, loss.neg_weights, loss.epsilon)
        self.opt = model.optimizer
        if self.world > 1:
            # replicas start from rank 0's weights and BN moving statistics whatever each process initialised with
            torch.distributed.broadcast(p.params.w, 0, group=process_group)
            torch.distributed.broadcast(p.params.f, 0, group=process_group)
            p.params.mark_updated()
        self.use_graph = use_graph
        self.buckets = buckets
        self.stream = torch.cuda.Stream()
        if overlap_wgrad:
            p.side_stream = torch.cuda.Stream()
        self.comm_stream = torch.cuda.Stream(priority=-1) if (self.world > 1 and exchange in ("overlap", "peer", "gather")) else None
        self.gather = dp.GatherExchange(p.params.n_train, p.device, process_group) \
            if (self.world > 1 and exchange == "gather") else None
        self.peer = dp.PeerExchange(p.params.n_train, p.device, process_group) \
            if (self.world > 1 and exchange == "peer") else None
        self._g16 = torch.empty(p.params.n_train, dtype=torch.bfloat16, device=p.device) \
            if (self.world > 1 and grad_dtype == "bfloat16") else None
        self._segments: List = []          # (graph or callable, grad-arena range completed by it)
        self._prep_graph = None
        # pinned host staging for the end-to-end path
        self.host_x = torch.empty(p.x_in.shape, dtype=torch.float32).pin_memory()
        self.host_y = torch.empty(tuple(p.labels.shape), dtype=torch.int32).pin_memory()
        self.host_loss = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.dev_x32 = torch.empty(p.x_in.shape, dtype=torch.float32, device=p.device) \
            if p.x_in.buf.dtype != torch.float32 else None
        # input double buffering (prefetch): the H2D copy of batch t+1 runs on a copy stream into staging buffers while
        # step t computes; the step then starts with a device-side cast/copy (~15 us) out of the staging buffers
        self.copy_stream = torch.cuda.Stream()
        self.stage_x = torch.empty(p.x_in.shape, dtype=torch.float32, device=p.device)
        self.stage_y = torch.empty(tuple(p.labels.shape), dtype=torch.int32, device=p.device)
        self._staged = None                # event: staging buffers hold a complete batch
        self._stage_free = None            # event: the training stream has consumed the staging buffers
        self._build(buckets if (self.world > 1 and exchange in ("overlap", "peer", "gather")) else 1)

    # ------------------------------------------------------------------------------------------------------
    def _build(self, n_buckets: int):
        p = self.plan
        # the backward launch list in n_buckets contiguous segments; after segment i the gradient-arena slices that
        # became final during it are exchanged while segment i+1 runs (dp.exchange_schedule; with one bucket this
        # degenerates to one all-reduce after backward)
        cuts, ranges = dp.exchange_schedule(p, n_buckets)

        def head():
            p.zero_grads()
            p.forward()
            p.loss_forward_backward()

        parts = [head] + [(lambda a=a, b=b: p.run_bwd_range(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
        self._ranges: List = [[]] + ranges             # nothing to exchange after the head
        self._cuts = None
        if self.use_graph and n_buckets > 1 and self.cut_events:
            segs = list(zip(cuts[:-1], cuts[1:]))
            marks = [(torch.cuda.Event(external=True),
                      torch.cuda.Event(external=True) if p.side_stream is not None else None) for _ in segs[:-1]]

            def whole():
                head()
                pend = None
                for i, (a, b) in enumerate(segs):
                    last = i == len(segs) - 1
                    pend = p.run_bwd_range(a, b, join=last, pending=pend)
                    if not last:
                        em, es = marks[i]
                        em.record(self.stream)                       # everything the training stream launched so far
                        if es is not None:
                            es.record(p.side_stream)                 # ... and the filter-gradient kernels launched so far
            parts = [whole]
            self._cuts = (marks, ranges)
            self._ranges = [[]]

        if self.use_graph:
            # warm-up outside capture (lazy module loading, cudaFuncSetAttribute) on the capture stream
            with torch.cuda.stream(self.stream):
                for part in parts:
                    part()
                p.run_prep()
            torch.cuda.synchronize()
            self._graphs = []
            for part in parts:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    part()
                self._graphs.append(g)
            self._prep_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._prep_graph, stream=self.stream):
                p.run_prep()
            self._parts = [g.replay for g in self._graphs]
            self._prep = self._prep_graph.replay
        else:
            self._parts = parts
            self._prep = p.run_prep
        # + fused decoder tail, gradient / statistics / loss memsets, L2 sum, two Adam launches, weight re-quantisation
        self.launches_per_step = p.launches_fwd + p.launches_bwd + 8

    # ------------------------------------------------------------------------------------------------------
    def stage_inputs(self, images: torch.Tensor, labels: torch.Tensor):
        """Host (pinned) -> device copies of one batch, asynchronous on the training stream."""
        p = self.plan
        with torch.cuda.stream(self.stream):
            if self.dev_x32 is not None:
                self.dev_x32.copy_(images, non_blocking=True)
                ops.cast(self.dev_x32, p.x_in.buf)
            else:
                p.x_in.buf.copy_(images, non_blocking=True)
            p.labels.copy_(labels, non_blocking=True)

    def prefetch(self, images: torch.Tensor, labels: torch.Tensor):
        """Start the H2D copy of the NEXT batch (pinned host tensors) on the copy stream; returns immediately.  The next
        train_step_e2e() call consumes it.  This is the input double buffering every training loop does (the
        reference's Keras enqueuer prepares the next batches on worker threads, ss.py:1063-1073)."""
        if self._stage_free is not None:
            self.copy_stream.wait_event(self._stage_free)
        with torch.cuda.stream(self.copy_stream):
            self.stage_x.copy_(images, non_blocking=True)
            self.stage_y.copy_(labels, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(self.copy_stream)

    def _consume_staged(self):
        p = self.plan
        self.stream.wait_event(self._staged)
        with torch.cuda.stream(self.stream):
            if p.x_in.buf.dtype != torch.float32:
                ops.cast(self.stage_x, p.x_in.buf)
            else:
                p.x_in.buf.copy_(self.stage_x, non_blocking=True)
            p.labels.copy_(self.stage_y, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(self.stream)
        self._staged = None

    def step(self, optimizer_step: bool = True):
        """One training step on the batch resident in the plan's input buffers."""
        p = self.plan
        with torch.cuda.stream(self.stream):
            p.ensure_current()                  # another plan of the model stepped, or weights were set / loaded
            for part, ranges in zip(self._parts, self._ranges):
                part()
                if self.world > 1 and ranges and self.exchange in ("overlap", "peer", "gather"):
                    self._exchange(ranges)      # all-reduce what this segment finished while the next one runs
            if self._cuts is not None:
                marks, ranges = self._cuts
                for (em, es), rg in zip(marks, ranges[:-1]):
                    self._exchange(rg, after=(em, es))      # waits for the cut marks recorded inside the running graph
                self._exchange(ranges[-1])                  # the rest, once the graph has completed
            if self.world > 1 and self.exchange == "gather":
                with torch.cuda.stream(self.comm_stream):
                    self.gather.finish_(self.plan.params.g)
                    self._comm_done = torch.cuda.Event()
                    self._comm_done.record(self.comm_stream)
            if self.world > 1 and self.exchange in ("overlap", "peer", "gather"):
                self.stream.wait_event(self._comm_done)
            elif self.world > 1 and self.exchange == "tail":
                self._exchange_tail()
            if optimizer_step:
                p.regularization()
                self._adam()
                self._prep()

    def _exchange(self, ranges, after=None):
        g = self.plan.params.g
        if after is None:
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.comm_stream.wait_event(ev)
        else:
            for ev in after:
                if ev is not None:
                    self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            if self.peer is not None:
                self.peer.all_reduce_(g, ranges)
            elif self.gather is not None:
                self.gather.push_(g, ranges)
            else:
                dp.allreduce_ranges(g, ranges, self.pg)
            self._comm_done = torch.cuda.Event()
            self._comm_done.record(self.comm_stream)

    def _exchange_tail(self):
        """One all-reduce of the complete gradient arena behind backward, on the training stream."""
        P = self.plan.params
        g = P.g[:P.n_train]
        if self._g16 is not None:
            ops.cast(g, self._g16)
            torch.distributed.all_reduce(self._g16, op=torch.distributed.ReduceOp.SUM, group=self.pg)
            ops.cast(self._g16, g)
        else:
            torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.SUM, group=self.pg)

    def _adam(self):
        P, opt = self.plan.params, self.opt
        lr_t = opt.step_size()
        gs = 1.0 / self.world
        if P.n_reg:
            ops.adam(P.w, P.g, P.m, P.v, P.n_reg, lr_t, opt.beta_1, opt.beta_2, opt.epsilon, gs, P.l2)
        if P.n_train > P.n_reg:
            ops.adam(P.w, P.g, P.m, P.v, P.n_train - P.n_reg, lr_t, opt.beta_1, opt.beta_2, opt.epsilon, gs, 0.0,
                     w_off=P.n_reg)
        opt.iterations += 1
        self.plan.step_counter.add_(1)
        P.mark_updated()
        self.plan._prep_version = P.version     # step() re-derives the bf16 operands right after (self._prep)

    def read_loss(self) -> float:
        """Device -> host read of the step's loss (data term + L2 term)."""
        p = self.plan
        with torch.cuda.stream(self.stream):
            self.host_loss[0:1].copy_(p.loss_sum, non_blocking=True)
            self.host_loss[1:2].copy_(p.reg_sum, non_blocking=True)
        self.stream.synchronize()
        P = p.N * p.out_shape[1] * p.out_shape[2]
        return float(self.host_loss[0]) / P + p.params.l2 * float(self.host_loss[1])

    def train_step_e2e(self, images: Optional[torch.Tensor], labels: Optional[torch.Tensor],
                       prefetch_next=None) -> float:
        """The user-facing call: pinned host batch in, its loss out (H2D + step + D2H).
        `prefetch_next=(images, labels)`: the batch of the NEXT call; its H2D copy is started behind this step's launch
        and overlaps the step's compute.  A call that follows one with `prefetch_next` trains on that prefetched batch
        (pass the same tensors, or None)."""
        if self._staged is not None:
            self._consume_staged()                     # prefetched by the previous call
        else:
            self.stage_inputs(images, labels)
        self.step()
        if prefetch_next is not None:
            self.prefetch(*prefetch_next)
        return self.read_loss()


class Predictor:
    """Serving-side driver (the reference's `segment()` / `Model.predict` use, ss.py:1207-1227): CUDA-graph replay of the
    inference plan — forward, bilinear up-sampling of the logits, channel argmax — with pinned host staging.

    `step()` runs on the batch resident in the plan's input buffer; `segment_e2e(images)` is the user-facing call:
    pinned host fp32 images in, int32 label maps out (H2D + graph + D2H, synchronous)."""

    def __init__(self, model, batch_size: int, dtype: Optional[str] = None, use_graph: bool = True):
        self.model = model
        self.plan: Plan = model.plan(batch_size, training=False, **({"dtype": dtype} if dtype else {}))
        p = self.plan
        self.stream = torch.cuda.Stream()
        small = p.out_shape[3] <= 256           # label ids fit a byte: a quarter of the label-map bytes (HBM and D2H)
        self.labels = torch.empty(p.out_shape[:3], dtype=torch.uint8 if small else torch.int32, device=p.device)
        self.host_x = torch.empty(p.x_in.shape, dtype=torch.float32).pin_memory()
        self.host_labels = torch.empty(p.out_shape[:3], dtype=self.labels.dtype).pin_memory()
        self._copy_stream = self._stage_x = self._staged = self._stage_free = None
        self.dev_x32 = torch.empty(p.x_in.shape, dtype=torch.float32, device=p.device) \
            if p.x_in.buf.dtype != torch.float32 else None
        self.use_graph = use_graph
        self._graph = None
        with torch.cuda.stream(self.stream):
            self._run()                                     # warm-up outside capture
        torch.cuda.synchronize()
        if use_graph:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, stream=self.stream):
                self._run()
        self.launches_per_step = p.launches_fwd + 1

    def _run(self):
        self.plan.segment_device(self.labels)

    def step(self):
        with torch.cuda.stream(self.stream):
            self.plan.ensure_current()                      # weights trained / loaded since the last call
            if self._graph is not None:
                self._graph.replay()
            else:
                self._run()

    def stage_inputs(self, images: torch.Tensor):
        p = self.plan
        with torch.cuda.stream(self.stream):
            if self.dev_x32 is not None:
                self.dev_x32.copy_(images, non_blocking=True)
                ops.cast(self.dev_x32, p.x_in.buf)
            else:
                p.x_in.buf.copy_(images, non_blocking=True)

    def prefetch(self, images: torch.Tensor):
        """Start the H2D copy of the NEXT batch on the copy stream (input double buffering, as Trainer.prefetch)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage_x = torch.empty(self.plan.x_in.shape, dtype=torch.float32, device=self.plan.device)
        if self._stage_free is not None:
            self._copy_stream.wait_event(self._stage_free)
        with torch.cuda.stream(self._copy_stream):
            self._stage_x.copy_(images, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(self._copy_stream)

    def _consume_staged(self):
        p = self.plan
        self.stream.wait_event(self._staged)
        with torch.cuda.stream(self.stream):
            if p.x_in.buf.dtype != torch.float32:
                ops.cast(self._stage_x, p.x_in.buf)
            else:
                p.x_in.buf.copy_(self._stage_x, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(self.stream)
        self._staged = None

    def segment_e2e(self, images: Optional[torch.Tensor], prefetch_next: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pinned host images [B,H,W,3] fp32 -> pinned host label maps [B,Ho,Wo] (uint8 when the class count fits,
        else int32): H2D + graph + D2H, synchronous.  `prefetch_next`: the NEXT call's batch, copied behind this call's
        compute; the call that follows uses it (pass None or the same tensor)."""
        if self._staged is not None:
            self._consume_staged()
        else:
            self.stage_inputs(images)
        self.step()
        with torch.cuda.stream(self.stream):
            self.host_labels.copy_(self.labels, non_blocking=True)
        if prefetch_next is not None:
            self.prefetch(prefetch_next)
        self.stream.synchronize()
        return self.host_labels

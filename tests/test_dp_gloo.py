"""N>1 host logic on CPU: two gloo ranks, each running the engine (through the tests/fake_ops.py test double) on its
shard of the batch, exchange gradients with the Trainer's segment-wise prefix schedule (dp.run_step_with_exchange); the averaged gradient must equal the mean of the
oracle's per-shard gradients (BatchNormalization statistics stay per replica)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util
from tests.test_ops_gpu import NW, PW


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from deeplabv3plus_keras_b200 import dp, engine
    from tests import fake_ops
    engine.ops = fake_ops
    conf = util.make_conf(base="mobilenetv2", image_size=65, width=32, aspp=util.DEFAULT_ASPP)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    G = 4
    lo, hi = dp.shard_bounds(G, rank, world)
    plan = engine.Plan(ss.model, hi - lo, training=True)
    x, y = util.synthetic_batch(conf, G, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x[lo:hi], y[lo:hi])
    # the Trainer's own data-parallel schedule: backward in 3 segments, the arena prefixes each segment finished are
    # all-reduced right behind it (dp.exchange_schedule / Plan.final_prefixes / dp.allreduce_ranges)
    dp.run_step_with_exchange(plan, 3, None)
    n = plan.params.n_train
    g = (plan.params.g[:n] / world).numpy().copy()
    # ... and it must equal ONE all-reduce of the complete arena after a plain backward
    plan.step_fwd_bwd()
    dp.allreduce_gradients(plan.params.g, n, None, buckets=3)
    g2 = (plan.params.g[:n] / world).numpy()
    assert np.allclose(g, g2, rtol=1e-5, atol=1e-7), float(np.abs(g - g2).max())
    cuts, ranges = dp.exchange_schedule(plan, 3)
    covered = sorted(r for rg in ranges for r in rg if r[1] > r[0])
    assert covered[0][0] == 0 and covered[-1][1] == n and sum(hi - lo for lo, hi in covered) == n
    names = {k: v for k, v in plan.gradients().items()}
    if rank == 0:
        q.put((g, {k: v / world for k, v in names.items()}, x, y))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_matches_mean_of_shards():
    from oracle import model as OM
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g, named, x, y = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    conf = util.make_conf(base="mobilenetv2", image_size=65, width=32, aspp=util.DEFAULT_ASPP)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    w = util.torch_weights(ss.model)
    lam = conf["hps"]["weight_decay"]
    regularised = {f"{l.name}/kernel" for l in ss.model.flat_layers() if getattr(l, "kernel_regularizer", None) is not None}
    acc = None
    for lo, hi in ((0, 2), (2, 4)):
        _, _, grads, _ = OM.loss_and_grads(conf, w, torch.from_numpy(x[lo:hi]).double(), torch.from_numpy(y[lo:hi]), PW, NW)
        acc = grads if acc is None else {k: acc[k] + grads[k] for k in grads}
    for k, v in acc.items():
        want = (v / 2).numpy().copy()
        if k in regularised:
            want -= 2 * lam * w[k].numpy()
        scale = max(np.abs(want).max(), 1e-3)
        err = np.abs(named[k] - want) / scale
        assert (err > 3e-2).mean() <= 2e-3 and err.max() < 0.3, (k, float(err.max()))


def test_bucket_ranges_cover_and_align():
    from deeplabv3plus_keras_b200 import dp
    for n in (8, 1000, 16_540_000):
        for b in (1, 3, 8):
            r = dp.bucket_ranges(n, b)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == c[0] for a, c in zip(r, r[1:]))
            assert all(lo % 8 == 0 for lo, _ in r)
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)


def test_peer_exchange_schedule_in_lockstep():
    """dp.PeerExchange (copy-engine all-reduce over symmetric memory): its push / reduce / pull phases executed in
    lockstep for W simulated ranks over plain tensors — every element of every range ends up as the sum over ranks on
    every rank (bit-identical across ranks: one rank reduces each slice), elements outside the ranges are untouched,
    ragged ranges (length not a multiple of W or 8, shorter than W) included."""
    import torch
    from deeplabv3plus_keras_b200 import dp

    for W in (2, 3, 8):
        n = 1000
        cap = -(-n // W) + 64
        bufs = [torch.full(((W + 1) * cap,), float("nan")) for _ in range(W)]

        class Hdl:
            def get_buffer(self, r, sizes, dtype, off):
                return bufs[r][off:off + sizes[0]]

            def barrier(self):
                pass

        xs = []
        for r in range(W):
            x = object.__new__(dp.PeerExchange)
            x.world, x.rank, x.cap, x.hdl, x.buf = W, r, cap, Hdl(), bufs[r]
            x.rows, x.red = bufs[r][:W * cap].view(W, cap), bufs[r][W * cap:]
            xs.append(x)
        gen = torch.Generator().manual_seed(W)
        gs = [torch.randn(n, generator=gen) for _ in range(W)]
        ref = torch.stack(gs).sum(0)
        for ranges in ([(0, 13), (500, 777)], [(13, 500), (777, 1000)], [(3, 5), (0, 0)], [(0, 0), (990, 997)]):
            before = [g.clone() for g in gs]
            segs = [x._slices(ranges) for x in xs]
            for x, g, sg in zip(xs, gs, segs):
                x._push(g, sg)
            for x, g, sg in zip(xs, gs, segs):
                x._reduce(g, sg)
            for x, g, sg in zip(xs, gs, segs):
                x._pull(g, sg)
            inside = torch.zeros(n, dtype=torch.bool)
            for lo, hi in ranges:
                inside[lo:hi] = True
            want = torch.stack(before).sum(0)
            for r in range(W):
                assert torch.equal(gs[r][~inside], before[r][~inside])
                assert torch.equal(gs[r][inside], gs[0][inside])                       # identical on every rank
                assert torch.allclose(gs[r][inside], want[inside], rtol=1e-6, atol=1e-6)


def test_gather_exchange_schedule_in_lockstep():
    """dp.GatherExchange: every rank pushes its slices into row `rank` of every rank's staging buffer, then each sums the
    rows in rank order — simulated for W ranks over plain tensors: the arena ends up as the sum over ranks, bit-identical
    on every rank, whatever the slicing of the pushes."""
    import torch
    from deeplabv3plus_keras_b200 import dp

    for W in (2, 5):
        n = 777
        bufs = [torch.full((W * n,), float("nan")) for _ in range(W)]

        class Hdl:
            def barrier(self):
                pass

        xs = []
        for r in range(W):
            x = object.__new__(dp.GatherExchange)
            x.world, x.rank, x.n, x.hdl, x.buf = W, r, n, Hdl(), bufs[r]
            x.rows = bufs[r].view(W, n)
            x._mine = [bufs[q][r * n:(r + 1) * n] for q in range(W)]
            xs.append(x)
        gen = torch.Generator().manual_seed(W)
        gs = [torch.randn(n + 5, generator=gen) for _ in range(W)]           # 5 trailing elements outside the arena
        want = torch.stack([g[:n] for g in gs]).sum(0)
        tails = [g[n:].clone() for g in gs]
        for ranges in ([(0, 100), (400, 500)], [(100, 400), (500, 700)], [(0, 0), (700, 777)]):
            for x, g in zip(xs, gs):
                x.push_(g, ranges)
        for x, g in zip(xs, gs):
            x.finish_(g)
        for r in range(W):
            assert torch.equal(gs[r][:n], gs[0][:n])
            assert torch.allclose(gs[r][:n], want, rtol=1e-6, atol=1e-6)
            assert torch.equal(gs[r][n:], tails[r])

"""Fused depthwise backward (dlv3p_dwconv3x3_bwd) against the two launches it replaces, at the cfg-2 shapes, timed
inside a replayed CUDA graph (20 back-to-back launches, median of 5 replays) like the training step runs them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
REP = 20


def timeit(fn, reps=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / REP)
    ts.sort()
    return ts[len(ts) // 2] * 1e3


for shape in [(16, 32, 32, 736), (16, 64, 64, 736), (16, 127, 127, 256), (16, 254, 254, 128), (16, 254, 254, 64)]:
    N, H, W, C = shape
    x = torch.randn(shape, device=dev).to(bf)
    g = torch.randn(shape, device=dev).to(bf)
    add = torch.randn(shape, device=dev).to(bf)
    dx = torch.empty_like(x)
    w = torch.randn(3, 3, C, device=dev)
    dwg = torch.zeros(3, 3, C, device=dev)
    sc, sh, mu, isd = (torch.rand(C, device=dev) + 0.5 for _ in range(4))
    red = torch.zeros(2 * C, device=dev)
    nbytes = x.numel() * 2
    t_d = timeit(lambda: ops.dwconv3x3_dgrad_bnred(g, w, x.shape, x, sc, sh, 1, mu, isd, red, out=dx))
    t_w = timeit(lambda: ops.dwconv3x3_wgrad(x, g, dwg, 1, (1, 1), in_scale=sc, in_shift=sh, in_act=1))
    t_f = timeit(lambda: ops.dwconv3x3_bwd(g, x, w, dwg, in_scale=sc, in_shift=sh, in_act=1, bn_mean=mu, bn_invstd=isd,
                                           bn_red=red, out=dx))
    print(f"{shape} bnred : dgrad_bnred {t_d:7.1f} + wgrad {t_w:7.1f} = {t_d + t_w:7.1f} us | fused {t_f:7.1f} us "
          f"({3 * nbytes / t_f / 1e3:6.0f} GB/s)", flush=True)
    t_d = timeit(lambda: ops.dwconv3x3_dgrad(g, w, x.shape, 1, (1, 1), x_pre=x, in_act=1, addend=add, out=dx))
    t_w = timeit(lambda: ops.dwconv3x3_wgrad(x, g, dwg, 1, (1, 1), in_act=1))
    t_f = timeit(lambda: ops.dwconv3x3_bwd(g, x, w, dwg, in_act=1, addend=add, out=dx))
    print(f"{shape} preact+add: dgrad {t_d:7.1f} + wgrad {t_w:7.1f} = {t_d + t_w:7.1f} us | fused {t_f:7.1f} us "
          f"({4 * nbytes / t_f / 1e3:6.0f} GB/s)", flush=True)
    t_d = timeit(lambda: ops.dwconv3x3_dgrad(g, w, x.shape, 1, (1, 1), x_pre=x, in_act=1, out=dx))
    t_f = timeit(lambda: ops.dwconv3x3_bwd(g, x, w, dwg, in_act=1, out=dx))
    print(f"{shape} preact    : dgrad {t_d:7.1f} + wgrad {t_w:7.1f} = {t_d + t_w:7.1f} us | fused {t_f:7.1f} us "
          f"({3 * nbytes / t_f / 1e3:6.0f} GB/s)", flush=True)

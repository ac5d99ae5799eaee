"""Isolated-kernel roofline measurements at BASELINE cfg-2 shapes (Xception OS16 513^2, batch 16, bf16).

Each kernel is timed with CUDA events on the launching stream after warm-up, with an L2 flush (a 256 MB write)
between repetitions; achieved = algorithmic bytes or flops (profiler.COSTS) / time.  `--only NAME` restricts the
set (used to keep ncu captures short).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import _lib, ops
from deeplabv3plus_keras_b200.profiler import cost

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--json", default="")
args = ap.parse_args()

dev = "cuda"
bf = torch.bfloat16
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
captured = {}


class Cap:
    def before(self, name, a):
        captured["last"] = (name, a)
        return None

    def after(self, tok):
        pass


def timeit(label, fn):
    if args.only and args.only not in label:
        return
    _lib.PROFILER = Cap()
    fn()
    _lib.PROFILER = None
    name, a = captured["last"]
    b, f = cost(name, a)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        if not args.no_flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    row = dict(label=label, kernel=name, ms=ms, GBps=b / ms / 1e6, TFLOPs=f / ms / 1e9, bytes=b, flops=f)
    results.append(row)
    print(f"{label:46s} {ms*1e3:9.1f} us  {row['GBps']:8.0f} GB/s  {row['TFLOPs']:8.1f} TFLOP/s", flush=True)


results = []
N = 16


def t(shape, dtype=bf):
    return torch.randn(shape, device=dev, dtype=torch.float32).to(dtype)


# ---- K1 depthwise: entry flow (HBM-resident) and middle flow (L2-resident) ----
for (H, C, dil) in [(254, 128, 1), (127, 256, 1), (64, 728, 1), (32, 728, 1), (32, 256, 12)]:
    x, w = t((N, H, H, C)), t((3, 3, C), torch.float32)
    y, dx, res = torch.empty_like(x), torch.empty_like(x), t((N, H, H, C))
    dwg = torch.zeros((3, 3, C), device=dev)
    tag = f"{H}x{H}x{C} r{dil}"
    timeit(f"dw_fwd {tag}", lambda: ops.dwconv3x3_fwd(x, w, 1, (dil, dil), in_act=1, out=y))
    timeit(f"dw_dgrad {tag}", lambda: ops.dwconv3x3_dgrad(y, w, x.shape, 1, (dil, dil), x_pre=x, in_act=1, out=dx))
    timeit(f"dw_dgrad+add {tag}", lambda: ops.dwconv3x3_dgrad(y, w, x.shape, 1, (dil, dil), x_pre=x, in_act=1,
                                                              addend=res, out=dx))
    timeit(f"dw_wgrad {tag}", lambda: ops.dwconv3x3_wgrad(x, y, dwg, 1, (dil, dil), in_act=1))

# ---- K2 GEMMs ----
for (M, Nn, K) in [(16384, 728, 728), (65536, 728, 728), (1032256, 128, 128), (1032256, 64, 288), (258064, 256, 256),
                   (16384, 256, 1280), (16384, 1024, 728)]:
    a, b = t((M, K)), t((Nn, K))
    c = torch.empty((M, Nn), device=dev, dtype=bf)
    stats = torch.zeros((2, Nn), device=dev)
    dwt = torch.zeros((K, Nn), device=dev)
    dy = t((M, Nn))
    timeit(f"gemm M{M} N{Nn} K{K}", lambda: ops.gemm_bf16(a, b, M, Nn, K, c))
    timeit(f"gemm+stats M{M} N{Nn} K{K}", lambda: ops.gemm_bf16(a, b, M, Nn, K, c, col_stats=stats))
    timeit(f"gemm_wgrad M{M} K{K} N{Nn}", lambda: ops.gemm_wgrad_bf16(a, dy, dwt, M, K, Nn))

# ---- K3 ----
for (M, C) in [(1032256, 128), (16384, 728)]:
    y, dz, out = t((M, C)), t((M, C)), torch.empty((M, C), device=dev, dtype=bf)
    sc, sh, mu, isd = (torch.rand(C, device=dev) + 0.5 for _ in range(4))
    red = torch.zeros((2, C), device=dev)
    timeit(f"affine_act M{M} C{C}", lambda: ops.affine_act(y, M, C, out, sc, sh, 1))
    timeit(f"bn_stats M{M} C{C}", lambda: ops.bn_stats(y, M, C, red))
    timeit(f"bn_bwd_reduce M{M} C{C}", lambda: ops.bn_bwd_reduce(dz, y, sc, sh, mu, isd, 1, M, C, red))
    timeit(f"bn_bwd_apply M{M} C{C}", lambda: ops.bn_bwd_apply(dz, y, sc, sh, mu, isd, 1, red, M, C, out))

x = t((N, 254, 254, 128))
am = torch.empty((N, 127, 127, 128), device=dev, dtype=torch.uint8)
yo = torch.empty((N, 127, 127, 128), device=dev, dtype=bf)
timeit("maxpool_fwd 254x254x128", lambda: ops.maxpool3x3s2_fwd(x, out=yo, argmax=am))
timeit("maxpool_bwd 254x254x128", lambda: ops.maxpool3x3s2_bwd(yo, am, x.shape, out=x))
x1 = t((N, 256, 256, 32))
colb = torch.empty((N * 254 * 254, 288), device=dev, dtype=bf)
timeit("im2col 256x256x32 -> K288", lambda: ops.im2col3x3(x1, 1, 1, 254, 254, 0, 0, 288, out=colb))
timeit("col2im K288 -> 256x256x32", lambda: ops.col2im3x3(colb, x1.shape, 1, 1, 254, 254, 0, 0, 288, out=x1))
xs = torch.empty((N, 127, 127, 64), device=dev, dtype=bf)
xb = t((N, 254, 254, 64))
timeit("subsample_fwd 254x254x64 s2", lambda: ops.subsample_fwd(xb, 2, out=xs))
timeit("subsample_bwd 254x254x64 s2", lambda: ops.subsample_bwd(xs, xb.shape, 2, out=xb))
del colb
zl = t((N, 32, 32, 21), torch.float32)
lab = torch.randint(0, 21, (N, 512, 512), device=dev, dtype=torch.int32)
pw, nw = torch.rand(21, device=dev), torch.rand(21, device=dev)
ls, dzl = torch.zeros(1, device=dev), torch.zeros_like(zl)
timeit("fused_loss_fwd 512x512x21", lambda: ops.upsample_softmax_cbloss_fwd(zl, lab, pw, nw, 1e-7, N, 32, 32, 21, 16, ls))
timeit("fused_loss_bwd 512x512x21", lambda: ops.upsample_softmax_cbloss_bwd(zl, lab, pw, nw, 1e-7, N, 32, 32, 21, 16,
                                                                              1.0, dzl))
timeit("fused_loss_fwd_bwd 512x512x21", lambda: ops.upsample_softmax_cbloss_fwd_bwd(zl, lab, pw, nw, 1e-7, N, 32, 32, 21,
                                                                                      16, 1.0, ls, dzl))
if args.json:
    json.dump(results, open(args.json, "w"), indent=1)

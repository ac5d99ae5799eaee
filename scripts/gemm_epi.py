"""Diagnostics: where does the epilogue of the 2-CTA GEMM spend its time?  20 back-to-back launches per timing."""
import os
import sys

os.environ["DLV3P_GEMM_DBG_ENABLE"] = "1"
os.environ.setdefault("DLV3P_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "deeplabv3plus_keras_b200", "libdlv3p_diag.so"))   # DLV3P_DIAG=1 bash csrc/build.sh
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
REP = 20


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(REP):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / REP)
    ts.sort()
    return ts[len(ts) // 2] * 1e3


MODES = [(0, "full"), (5, "epi only"), (13, "epi, no store"), (21, "epi, no tmem ld"), (29, "epi, neither"), (7, "empty"),
         (3, "MMA only"), (6, "loads only"), (2, "loads+MMA"), (1, "MMA+epi"), (8, "full, no store")]
for (M, N, K) in [(16384, 728, 728), (65536, 728, 728)]:
    a = torch.randn((M, K), device=dev).to(bf)
    b = torch.randn((N, K), device=dev).to(bf)
    c = torch.empty((M, N), device=dev, dtype=bf)
    dy = torch.randn((M, N), device=dev).to(bf)
    dw = torch.zeros((K, N), device=dev)
    stats = torch.zeros((2, N), device=dev)
    for mode, name in MODES:
        os.environ["DLV3P_GEMM_DBG"] = str(mode)
        u1 = timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c))
        u2 = timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c, col_stats=stats))
        u3 = timeit(lambda: ops.gemm_wgrad_bf16(a, dy, dw, M, K, N))
        print(f"M{M} N{N} K{K} dbg={mode:2d} {name:16s} gemm {u1:6.2f} | +stats {u2:6.2f} | wgrad {u3:6.2f} us", flush=True)
    os.environ["DLV3P_GEMM_DBG"] = "0"

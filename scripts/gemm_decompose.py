"""Bottleneck decomposition of the 2-CTA tcgen05 GEMM (diagnostics, not a bench): the same launch with operand loads,
epilogue work and/or MMAs switched off (DLV3P_GEMM_DBG bit mask 1 / 2 / 4), next to cuBLAS (torch.matmul) on the same
shape.  Run on the GPU box: DLV3P_GEMM_DBG_ENABLE=1 python scripts/gemm_decompose.py"""
import os
import sys

os.environ["DLV3P_GEMM_DBG_ENABLE"] = "1"
os.environ.setdefault("DLV3P_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "deeplabv3plus_keras_b200", "libdlv3p_diag.so"))   # DLV3P_DIAG=1 bash csrc/build.sh
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, flush=True, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


MODES = [(0, "full"), (1, "no loads"), (2, "no epilogue"), (3, "MMA only"), (4, "loads+epilogue"), (6, "loads only"),
         (5, "epilogue only")]
for (M, N, K) in [(16384, 728, 728), (65536, 728, 728), (16384, 1024, 728), (16384, 256, 1280)]:
    a = torch.randn((M, K), device=dev).to(bf)
    b = torch.randn((N, K), device=dev).to(bf)
    c = torch.empty((M, N), device=dev, dtype=bf)
    dy = torch.randn((M, N), device=dev).to(bf)
    dw = torch.zeros((K, N), device=dev)
    stats = torch.zeros((2, N), device=dev)
    fl = 2.0 * M * N * K
    for flush in (True, False):
        tag = "L2 flushed" if flush else "L2 warm"
        us = timeit(lambda: torch.matmul(a, b.t(), out=c), flush)
        print(f"M{M} N{N} K{K} [{tag}] cuBLAS a@b.T            {us:8.1f} us {fl/us/1e6:8.1f} TFLOP/s", flush=True)
        us = timeit(lambda: torch.matmul(a.t(), dy, out=dw.to(bf)), flush)
        print(f"M{M} N{N} K{K} [{tag}] cuBLAS a.T@dy (wgrad)   {us:8.1f} us {fl/us/1e6:8.1f} TFLOP/s", flush=True)
        for mode, name in MODES:
            os.environ["DLV3P_GEMM_DBG"] = str(mode)
            us = timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c), flush)
            us2 = timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c, col_stats=stats), flush)
            us3 = timeit(lambda: ops.gemm_wgrad_bf16(a, dy, dw, M, K, N), flush)
            print(f"M{M} N{N} K{K} [{tag}] dbg={mode} {name:15s} gemm {us:7.1f} us {fl/us/1e6:7.1f} TF | +stats {us2:7.1f} us"
                  f" {fl/us2/1e6:7.1f} TF | wgrad {us3:7.1f} us {fl/us3/1e6:7.1f} TF", flush=True)
        os.environ["DLV3P_GEMM_DBG"] = "0"

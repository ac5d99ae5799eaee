"""Diagnostics: does the 1456-byte row pitch of the 728-channel tensors (rows start mid-sector / straddle two L2 lines)
limit the TMA operand feed?  Same logical GEMM with leading dimensions 728 / 736 (sector aligned) / 768 (line aligned)."""
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


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


for (M, N, K) in [(65536, 728, 728), (16384, 728, 728), (65536, 768, 768), (16384, 768, 768), (262144, 728, 728)]:
    fl = 2.0 * M * N * K
    for ld in (728, 736, 768):
        if ld < max(N, K):
            continue
        a = torch.randn((M, ld), device=dev).to(bf)
        b = torch.randn((N, ld), device=dev).to(bf)
        c = torch.empty((M, ld), device=dev, dtype=bf)
        dy = torch.randn((M, ld), device=dev).to(bf)
        dw = torch.zeros((K, ld), device=dev)
        stats = torch.zeros((2, N), device=dev)
        av, bv, cv = a[:, :K], b[:, :K], c[:, :N]
        us = timeit(lambda: torch.matmul(av, bv.t(), out=cv)) if ld == max(N, K) else float("nan")
        row = f"M{M} N{N} K{K} ld={ld} cuBLAS {us:7.1f} us |"
        for mode, name in [(0, "full"), (6, "loads only"), (3, "MMA only")]:
            os.environ["DLV3P_GEMM_DBG"] = str(mode)
            u1 = timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c, lda=ld, ldb=ld, ldc=ld))
            u2 = timeit(lambda: ops.gemm_bf16(a, b, M, N, K, c, lda=ld, ldb=ld, ldc=ld, col_stats=stats))
            u3 = timeit(lambda: ops.gemm_wgrad_bf16(a, dy, dw, M, K, N, ldx=ld, ldy=ld, ldw=ld))
            row += f" {name}: {u1:6.1f} / {u2:6.1f} / {u3:6.1f} us ({fl/u1/1e6:6.0f} TF) |"
        os.environ["DLV3P_GEMM_DBG"] = "0"
        print(row, flush=True)

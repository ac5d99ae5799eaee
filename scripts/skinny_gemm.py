"""Narrow GEMMs of MobileNetV2 at 512x1024 (BASELINE cfg-5): time per shape with CUDA events (L2 flushed by the 0.9 GB
operand itself), for an ncu capture of gemm_tc_kernel<32,*> / <64,*>.  python scripts/skinny_gemm.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

torch.manual_seed(0)
for M, K, N in ((4194304, 96, 16), (4194304, 16, 96), (1048576, 144, 24), (1048576, 24, 144), (262144, 192, 32)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, (K + 7) // 8 * 8, device="cuda").bfloat16()
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sc, sh = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    for _ in range(3):
        ops.gemm_bf16(a, w, M, N, K, c, lda=K, ldb=w.shape[1], ldc=N, col_scale=sc, col_shift=sh, act=ops.ACT_RELU6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm_bf16(a, w, M, N, K, c, lda=K, ldb=w.shape[1], ldc=N, col_scale=sc, col_shift=sh, act=ops.ACT_RELU6)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (M * K + M * N) * 2 / 1e9
    print(f"M={M} K={K} N={N}: {ms * 1e3:.1f} us, {gb / ms * 1e3:.0f} GB/s ({gb / ms * 1e3 / 6542.7 * 100:.0f} % of HBM)")
    ref = torch.clamp(a[:4096].float() @ w[:, :K].float().t(), 0, 6)
    err = (c[:4096].float() - ref).abs().max().item()
    assert err < 0.1 * ref.abs().max().item() + 0.05, err

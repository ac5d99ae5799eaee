"""Bottleneck decomposition of the 1-CTA tcgen05 GEMM on MobileNetV2's narrow shapes (diagnostics build only:
DLV3P_DIAG=1 bash deeplabv3plus_keras_b200/csrc/build.sh): the same launch with operand loads (1), epilogue work (2)
and / or MMAs (4) switched off."""
import os
import sys

os.environ["DLV3P_GEMM_DBG_ENABLE"] = "1"
os.environ.setdefault("DLV3P_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "deeplabv3plus_keras_b200", "libdlv3p_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

MODES = [(0, "full"), (1, "no loads"), (2, "no epilogue"), (4, "no MMA"), (3, "MMA only"), (6, "loads only"), (5, "epilogue only"),
         (7, "empty")]
for M, K, N in ((4194304, 96, 16), (4194304, 16, 96), (1048576, 144, 24), (1048576, 24, 144), (4194304, 32, 32)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, (K + 7) // 8 * 8, device="cuda").bfloat16()
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sc, sh = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    gb = (M * K + M * N) * 2 / 1e9
    out = []
    for mode, name in MODES:
        os.environ["DLV3P_GEMM_DBG"] = str(mode)
        for _ in range(2):
            ops.gemm_bf16(a, w, M, N, K, c, lda=K, ldb=w.shape[1], ldc=N, col_scale=sc, col_shift=sh, act=ops.ACT_RELU6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gemm_bf16(a, w, M, N, K, c, lda=K, ldb=w.shape[1], ldc=N, col_scale=sc, col_shift=sh, act=ops.ACT_RELU6)
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 5 * 1e3:.0f}")
    os.environ["DLV3P_GEMM_DBG"] = "0"
    print(f"M={M} K={K} N={N} ({gb:.2f} GB, HBM floor {gb / 6.5427 * 1e3:.0f} us): " + " | ".join(out) + " us", flush=True)

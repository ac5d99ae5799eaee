"""One short process that launches the step's three hot kernel families at BASELINE cfg-2 shapes, for a single
`ncu --set full` capture (profiles/): 2-CTA tcgen05 GEMM 16384x728x728 (+BN statistics epilogue), its filter-gradient
variant, and the TMA depthwise forward / input-gradient / filter-gradient on the entry-flow [16,254,254,128] tensor."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
M, N, K = 16384, 728, 728
a = torch.randn(M, K, device=dev).to(bf)
b = torch.randn(N, K, device=dev).to(bf)
c = torch.empty(M, N, device=dev, dtype=bf)
dy = torch.randn(M, N, device=dev).to(bf)
stats = torch.zeros(2, N, device=dev)
dw = torch.zeros(K, N, device=dev)
x = torch.randn(16, 254, 254, 128, device=dev).to(bf)
y, dx = torch.empty_like(x), torch.empty_like(x)
w = torch.randn(3, 3, 128, device=dev)
dwg = torch.zeros(3, 3, 128, device=dev)
for _ in range(3):
    ops.gemm_bf16(a, b, M, N, K, c, col_stats=stats)
    ops.gemm_wgrad_bf16(a, dy, dw, M, K, N)
    ops.dwconv3x3_fwd(x, w, 1, (1, 1), in_act=1, out=y)
    ops.dwconv3x3_dgrad(y, w, x.shape, 1, (1, 1), x_pre=x, in_act=1, out=dx)
    ops.dwconv3x3_wgrad(x, y, dwg, 1, (1, 1), in_act=1)
torch.cuda.synchronize()
print("ok")

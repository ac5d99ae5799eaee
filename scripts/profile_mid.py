"""One short process that launches the middle-flow kernels of BASELINE cfg-2 ([16,32,32,736] tensors, L2-resident,
20-30 us each — 54 % of the step) for a single `ncu --set full` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
N, H, C = 16, 32, 736
M = N * H * H
x = torch.randn(N, H, H, C, device=dev).to(bf)
y, dx, g = torch.empty_like(x), torch.empty_like(x), torch.randn(N, H, H, C, device=dev).to(bf)
w = torch.randn(3, 3, C, device=dev)
dwg = torch.zeros(3, 3, C, device=dev)
sc, sh, mu, isd = (torch.rand(C, device=dev) + 0.5 for _ in range(4))
red = torch.zeros(2 * C, device=dev)
a = torch.randn(M, C, device=dev).to(bf)
b = torch.randn(C, C, device=dev).to(bf)
c = torch.empty(M, C, device=dev, dtype=bf)
stats = torch.zeros(2, C, device=dev)
dwt = torch.zeros(C, C, device=dev)
for _ in range(3):
    ops.dwconv3x3_fwd(x, w, 1, (1, 1), in_scale=sc, in_shift=sh, in_act=1, out=y)
    ops.dwconv3x3_dgrad_bnred(g, w, x.shape, x, sc, sh, 1, mu, isd, red, out=dx)
    ops.dwconv3x3_wgrad(x, g, dwg, 1, (1, 1), in_scale=sc, in_shift=sh, in_act=1)
    ops.bn_bwd_apply(g, x, sc, sh, mu, isd, 0, red, M, C, y)
    ops.gemm_bf16(a, b, M, C, C, c, col_stats=stats)
    ops.gemm_wgrad_bf16(a, g.view(M, C), dwt, M, C, C)
torch.cuda.synchronize()
print("ok")

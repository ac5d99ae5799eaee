"""One short process that launches the kernels added in round 2 at their BASELINE shapes for a single `ncu --set full`
capture: the implicit-GEMM SAME 3x3 convolution of cfg-4 (304 channels at 256 x 256 -> 21 classes; batch 4 here),
the fused up-sampling + argmax of cfg-5 (19 classes, x16 to 1024 x 2048; batch 2), the narrow-N GEMM epilogue
(MobileNetV2 projection, M = 1 M, K = 144, N = 24) and, as the reference point, the middle-flow GEMM of cfg-2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
N, H, W, Cin, Cout = 4, 256, 256, 304, 21
x = torch.randn(N, H, W, Cin, device=dev).to(bf)
wt = torch.randn(Cout, 9 * Cin, device=dev).to(bf)
kp = 64
wd = torch.zeros(Cin, 9 * kp, device=dev, dtype=bf)
y = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float32)
dy = torch.zeros(N, H, W, 24, device=dev, dtype=bf)
dy[..., :Cout] = torch.randn(N, H, W, Cout, device=dev).to(bf)
dx = torch.empty(N, H, W, Cin, device=dev, dtype=bf)
dw = torch.zeros(3, 3, Cin, Cout, device=dev)
z = torch.randn(2, 64, 128, 19, device=dev)
lab = torch.empty(2, 1024, 2048, dtype=torch.uint8, device=dev)
M2, K2, N2 = 1048576, 144, 24
a2 = torch.randn(M2, K2, device=dev).to(bf)
w2 = torch.randn(N2, K2, device=dev).to(bf)
c2 = torch.empty(M2, N2, device=dev, dtype=bf)
sc2, sh2 = torch.ones(N2, device=dev), torch.zeros(N2, device=dev)
M, C = 16384, 736
a = torch.randn(M, C, device=dev).to(bf)
b = torch.randn(C, C, device=dev).to(bf)
c = torch.empty(M, C, device=dev, dtype=bf)
stats = torch.zeros(2, C, device=dev)
for _ in range(2):
    ops.conv3x3_same_fwd(x, wt, y, Cout)
    ops.conv3x3_same_dgrad(dy, 24, wd, (N, H, W, Cin), Cout, dx)
    ops.conv3x3_same_wgrad(x, dy, 24, dw, Cout)
    ops.upsample_argmax(z, 16, 16, lab)
    ops.gemm_bf16(a2, w2, M2, N2, K2, c2, lda=K2, ldb=K2, ldc=N2, col_scale=sc2, col_shift=sh2, act=ops.ACT_RELU6)
    ops.gemm_bf16(a, b, M, C, C, c, col_stats=stats)
torch.cuda.synchronize()
print("ok")

"""Fused decoder tail (x16 upsample + softmax + class-balanced loss + gradient) at BASELINE cfg-2 for one ncu capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev = "cuda"
N = 16
zl = torch.randn(N, 32, 32, 21, device=dev)
lab = torch.randint(0, 21, (N, 512, 512), device=dev, dtype=torch.int32)
pw, nw = torch.rand(21, device=dev), torch.rand(21, device=dev)
ls, dzl = torch.zeros(1, device=dev), torch.zeros_like(zl)
for _ in range(3):
    ops.upsample_softmax_cbloss_fwd_bwd(zl, lab, pw, nw, 1e-7, N, 32, 32, 21, 16, 1.0, ls, dzl)
torch.cuda.synchronize()
print("ok")

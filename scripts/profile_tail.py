"""Fused decoder tail (upsample + softmax + class-balanced loss + gradient) for one ncu capture: x16 at BASELINE cfg-2
(logits [16,32,32,21]) and x2 at cfg-4 (boundary refinement, logits [16,256,256,21]); 3 launches of each, x16 first.
  ncu --set full --clock-control none --import-source on -k regex:tail_pixel -s 2 -c 1 ...   (third x16 launch)
  ncu ... -k regex:tail_pixel -s 5 -c 1 ...                                                    (third x2 launch)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev = "cuda"
N = 16
lab = torch.randint(0, 21, (N, 512, 512), device=dev, dtype=torch.int32)
pw, nw = torch.rand(21, device=dev), torch.rand(21, device=dev)
for hw, f in ((32, 16), (256, 2)):
    zl = torch.randn(N, hw, hw, 21, device=dev)
    ls, dzl = torch.zeros(1, device=dev), torch.zeros_like(zl)
    for _ in range(3):
        ops.upsample_softmax_cbloss_fwd_bwd(zl, lab, pw, nw, 1e-7, N, hw, hw, 21, f, 1.0, ls, dzl)
    torch.cuda.synchronize()
print("ok")

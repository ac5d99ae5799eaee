"""ncu driver for the fused depthwise backward: a few launches at the middle-flow and the largest entry-flow shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeplabv3plus_keras_b200 import ops

dev, bf = "cuda", torch.bfloat16
for shape in [(16, 32, 32, 736), (16, 254, 254, 128)]:
    N, H, W, C = shape
    x = torch.randn(shape, device=dev).to(bf)
    g = torch.randn(shape, device=dev).to(bf)
    dx = torch.empty_like(x)
    w = torch.randn(3, 3, C, device=dev)
    dwg = torch.zeros(3, 3, C, device=dev)
    sc, sh, mu, isd = (torch.rand(C, device=dev) + 0.5 for _ in range(4))
    red = torch.zeros(2 * C, device=dev)
    for _ in range(3):
        ops.dwconv3x3_bwd(g, x, w, dwg, in_scale=sc, in_shift=sh, in_act=1, bn_mean=mu, bn_invstd=isd, bn_red=red, out=dx)
        ops.dwconv3x3_dgrad_bnred(g, w, x.shape, x, sc, sh, 1, mu, isd, red, out=dx)
    torch.cuda.synchronize()

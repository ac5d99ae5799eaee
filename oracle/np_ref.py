"""ORACLE cross-check (test infrastructure only): pure-numpy, loop-level restatement of the TF op formulas used on
the hot path, written directly from the published semantics and sharing no code with oracle/tf_ops.py
(which leans on torch's conv/pool primitives).  Small cases only — it is O(N*H*W*C*k*k) Python loops.

Used by tests/test_oracle.py to pin tf_ops.py: asymmetric SAME padding, dilation, max-pool padding that never wins,
half-pixel bilinear resize, fused-batch-norm moving-variance convention, the class-balanced loss.
"""
from __future__ import annotations

import math

import numpy as np


def same_geometry(n, k, s, d=1):
    k_eff = (k - 1) * d + 1
    out = -(-n // s)
    total = max((out - 1) * s + k_eff - n, 0)
    return out, total // 2


def conv2d(x, w, stride=1, padding="same", dilation=(1, 1)):
    N, H, W, C = x.shape
    kh, kw, _, Co = w.shape
    if padding == "same":
        Ho, pt = same_geometry(H, kh, stride, dilation[0])
        Wo, pl = same_geometry(W, kw, stride, dilation[1])
    else:
        Ho = (H - ((kh - 1) * dilation[0] + 1)) // stride + 1
        Wo = (W - ((kw - 1) * dilation[1] + 1)) // stride + 1
        pt = pl = 0
    y = np.zeros((N, Ho, Wo, Co))
    for ho in range(Ho):
        for wo in range(Wo):
            for i in range(kh):
                for j in range(kw):
                    hi = ho * stride - pt + i * dilation[0]
                    wi = wo * stride - pl + j * dilation[1]
                    if 0 <= hi < H and 0 <= wi < W:
                        y[:, ho, wo, :] += x[:, hi, wi, :] @ w[i, j]
    return y


def depthwise_conv2d(x, w, stride=1, padding="same", dilation=(1, 1)):
    N, H, W, C = x.shape
    kh, kw = w.shape[:2]
    if padding == "same":
        Ho, pt = same_geometry(H, kh, stride, dilation[0])
        Wo, pl = same_geometry(W, kw, stride, dilation[1])
    else:
        Ho = (H - ((kh - 1) * dilation[0] + 1)) // stride + 1
        Wo = (W - ((kw - 1) * dilation[1] + 1)) // stride + 1
        pt = pl = 0
    y = np.zeros((N, Ho, Wo, C))
    for ho in range(Ho):
        for wo in range(Wo):
            for i in range(kh):
                for j in range(kw):
                    hi = ho * stride - pt + i * dilation[0]
                    wi = wo * stride - pl + j * dilation[1]
                    if 0 <= hi < H and 0 <= wi < W:
                        y[:, ho, wo, :] += x[:, hi, wi, :] * w[i, j, :, 0]
    return y


def max_pool_3x3_s2_same(x):
    N, H, W, C = x.shape
    Ho, pt = same_geometry(H, 3, 2)
    Wo, pl = same_geometry(W, 3, 2)
    y = np.full((N, Ho, Wo, C), -np.inf)
    for ho in range(Ho):
        for wo in range(Wo):
            for i in range(3):
                for j in range(3):
                    hi, wi = ho * 2 - pt + i, wo * 2 - pl + j
                    if 0 <= hi < H and 0 <= wi < W:
                        y[:, ho, wo, :] = np.maximum(y[:, ho, wo, :], x[:, hi, wi, :])
    return y


def avg_pool_valid(x, k):
    N, H, W, C = x.shape
    y = np.zeros((N, H // k, W // k, C))
    for ho in range(H // k):
        for wo in range(W // k):
            y[:, ho, wo, :] = x[:, ho * k:(ho + 1) * k, wo * k:(wo + 1) * k, :].mean(axis=(1, 2))
    return y


def resize_bilinear(x, fh, fw):
    N, H, W, C = x.shape
    y = np.zeros((N, H * fh, W * fw, C))
    for yo in range(H * fh):
        sy = (yo + 0.5) / fh - 0.5
        y0, y1, ly = max(math.floor(sy), 0), min(math.ceil(sy), H - 1), sy - math.floor(sy)
        for xo in range(W * fw):
            sx = (xo + 0.5) / fw - 0.5
            x0, x1, lx = max(math.floor(sx), 0), min(math.ceil(sx), W - 1), sx - math.floor(sx)
            top = x[:, y0, x0] + (x[:, y0, x1] - x[:, y0, x0]) * lx
            bot = x[:, y1, x0] + (x[:, y1, x1] - x[:, y1, x0]) * lx
            y[:, yo, xo] = top + (bot - top) * ly
    return y


def batch_norm_train(x, gamma, beta, mm, mv, eps, momentum):
    n = x.shape[0] * x.shape[1] * x.shape[2]
    mean = x.reshape(-1, x.shape[-1]).mean(0)
    var = ((x.reshape(-1, x.shape[-1]) - mean) ** 2).mean(0)
    y = (x - mean) / np.sqrt(var + eps) * gamma + beta
    unbiased = var * n / (n - 1)
    return y, momentum * mm + (1 - momentum) * mean, momentum * mv + (1 - momentum) * unbiased


def class_balanced_loss(y_true, y_pred, pw, nw, eps=1e-7):
    total = 0.0
    C = y_true.shape[-1]
    flat_t, flat_p = y_true.reshape(-1, C), y_pred.reshape(-1, C)
    for px in range(flat_t.shape[0]):
        for i in range(C):
            total += -(pw[i] * flat_t[px, i] * math.log(flat_p[px, i] + eps)
                       + nw[i] * (1.0 - flat_t[px, i]) * math.log(1.0 - flat_p[px, i] + eps))
    return total / flat_t.shape[0]

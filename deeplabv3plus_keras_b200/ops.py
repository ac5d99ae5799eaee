"""Operator layer: one Python function per C-ABI entry point, on torch CUDA tensors (NHWC).

torch is used for device memory and streams only; every function forwards raw device pointers to libdlv3p.so on
``torch.cuda.current_stream()``.  Outputs are allocated when not supplied (tests); the engine passes its own
pre-planned buffers so nothing is allocated inside a training step (CUDA-graph safe).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6, BF16, F32, call

Tensor = torch.Tensor


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise ValueError(f"unsupported tensor dtype {t.dtype}")


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous NHWC tensor")
    return t


def same_pad(n: int, k: int, stride: int, dil: int = 1) -> Tuple[int, int]:
    """TF 'SAME': out = ceil(n/s); total pad = max((out-1)*s + k_eff - n, 0); before = total // 2."""
    k_eff = (k - 1) * dil + 1
    out = -(-n // stride)
    total = max((out - 1) * stride + k_eff - n, 0)
    return out, total // 2


def valid_out(n: int, k: int, stride: int, dil: int = 1) -> int:
    k_eff = (k - 1) * dil + 1
    return (n - k_eff) // stride + 1


def conv_geometry(h: int, w: int, k: int, stride: int, dil: Tuple[int, int], padding: str):
    if padding == "same":
        ho, pt = same_pad(h, k, stride, dil[0])
        wo, pl = same_pad(w, k, stride, dil[1])
    elif padding == "valid":
        ho, pt = valid_out(h, k, stride, dil[0]), 0
        wo, pl = valid_out(w, k, stride, dil[1]), 0
    else:
        raise ValueError(f"padding must be 'same' or 'valid', got {padding!r}")
    return ho, wo, pt, pl


# ---------------------------------------------------------------------------------------------- K1
def dwconv3x3_fwd(x: Tensor, w: Tensor, stride=1, dil=(1, 1), padding="same", in_scale=None, in_shift=None,
                  in_act=ACT_NONE, out: Optional[Tensor] = None, pad: Optional[Tuple[int, int, int, int]] = None):
    """x [N,H,W,C]; w [3,3,C] fp32.  pad=(ho,wo,pad_t,pad_l) overrides `padding` (MobileNetV2 ZeroPadding2D)."""
    _chk(x, "x")
    N, H, W, Cc = x.shape
    ho, wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=x.dtype, device=x.device)
    call("dlv3p_dwconv3x3_fwd", _p(x), _p(w), _p(out), N, H, W, Cc, stride, dil[0], dil[1], pt, pl, ho, wo,
         _p(in_scale), _p(in_shift), in_act, _dt(x), _stream())
    return out


def dwconv3x3_fwd_epi(x: Tensor, w: Tensor, out_scale: Tensor, out_shift: Tensor, out_act: int, stride=1, dil=(1, 1),
                      padding="same", out: Optional[Tensor] = None, pad: Optional[Tuple[int, int, int, int]] = None):
    """Inference DepthwiseConv2D -> BatchNormalization -> ReLU/ReLU6: act(out_scale * conv(x) + out_shift) in one launch."""
    _chk(x, "x")
    N, H, W, Cc = x.shape
    ho, wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=x.dtype, device=x.device)
    call("dlv3p_dwconv3x3_fwd_epi", _p(x), _p(w), _p(out), N, H, W, Cc, stride, dil[0], dil[1], pt, pl, ho, wo,
         _p(out_scale), _p(out_shift), out_act, _dt(x), _stream())
    return out


def dwconv3x3_dgrad(dy: Tensor, w: Tensor, x_shape, stride=1, dil=(1, 1), padding="same", x_pre=None, in_scale=None,
                    in_shift=None, in_act=ACT_NONE, addend=None, out: Optional[Tensor] = None, pad=None):
    _chk(dy, "dy")
    N, H, W, Cc = x_shape
    ho, wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    assert tuple(dy.shape) == (N, ho, wo, Cc), (dy.shape, (N, ho, wo, Cc))
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    call("dlv3p_dwconv3x3_dgrad", _p(dy), _p(w), _p(out), N, H, W, Cc, stride, dil[0], dil[1], pt, pl, ho, wo,
         _p(x_pre), _p(in_scale), _p(in_shift), in_act, _p(addend), _dt(dy), _stream())
    return out


def dwconv3x3_bn_fwd(x: Tensor, w: Tensor, bn_sums, gamma, beta, moving_mean, moving_var, count, eps, momentum,
                     updates: int, in_act: int, scale, shift, mean, invstd, out: Optional[Tensor] = None, pad=None):
    """dwconv3x3_fwd on act(BN_train(x)) with bn_finalize folded in (stride 1, dilation 1, bf16)."""
    _chk(x, "x")
    N, H, W, Cc = x.shape
    ho, wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, 1, (1, 1), "same")
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=x.dtype, device=x.device)
    call("dlv3p_dwconv3x3_bn_fwd", _p(x), _p(w), _p(out), N, H, W, Cc, pt, pl, ho, wo, _p(bn_sums), _p(gamma), _p(beta),
         _p(moving_mean), _p(moving_var), float(count), eps, momentum, updates, in_act, _p(scale), _p(shift), _p(mean),
         _p(invstd), _dt(x), _stream())
    return out


def dwconv3x3_dgrad_bnred(dy: Tensor, w: Tensor, x_shape, x_pre: Tensor, in_scale: Tensor, in_shift: Tensor, in_act: int,
                          bn_mean: Tensor, bn_invstd: Tensor, bn_red: Tensor, out: Optional[Tensor] = None, pad=None):
    """dwconv3x3_dgrad (stride 1, dilation 1, bf16) + the BN-backward reductions of x_pre's layer into bn_red[2C]."""
    _chk(dy, "dy")
    N, H, W, Cc = x_shape
    ho, wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, 1, (1, 1), "same")
    assert tuple(dy.shape) == (N, ho, wo, Cc), (dy.shape, (N, ho, wo, Cc))
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    call("dlv3p_dwconv3x3_dgrad_bnred", _p(dy), _p(w), _p(out), N, H, W, Cc, pt, pl, ho, wo, _p(x_pre), _p(in_scale),
         _p(in_shift), in_act, _p(bn_mean), _p(bn_invstd), _p(bn_red), _dt(dy), _stream())
    return out


def dwconv3x3_wgrad(x: Tensor, dy: Tensor, dw: Tensor, stride=1, dil=(1, 1), padding="same", in_scale=None,
                    in_shift=None, in_act=ACT_NONE, pad=None):
    """dw [3,3,C] fp32 is ACCUMULATED into."""
    _chk(x, "x"); _chk(dy, "dy")
    N, H, W, Cc = x.shape
    ho, wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    call("dlv3p_dwconv3x3_wgrad", _p(x), _p(dy), _p(dw), N, H, W, Cc, stride, dil[0], dil[1], pt, pl, ho, wo,
         _p(in_scale), _p(in_shift), in_act, _dt(x), _stream())
    return dw


def dwconv3x3_bwd(dy: Tensor, x: Tensor, w: Tensor, dw: Tensor, in_scale=None, in_shift=None, in_act=ACT_NONE,
                  addend=None, bn_mean=None, bn_invstd=None, bn_red=None, bn_y=None, out: Optional[Tensor] = None):
    """Input gradient + filter gradient (+ BN-backward reductions: of x's layer, or with bn_y of the layer whose raw
    output is bn_y, taken on the final gradient) of a stride-1 dilation-1 SAME depthwise conv in one launch; dw [3,3,C]
    fp32 and bn_red [2C] are ACCUMULATED into.  Returns dx."""
    _chk(dy, "dy"); _chk(x, "x")
    N, H, W, Cc = x.shape
    assert tuple(dy.shape) == (N, H, W, Cc), (dy.shape, x.shape)
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    call("dlv3p_dwconv3x3_bwd", _p(dy), _p(x), _p(w), _p(out), _p(dw), N, H, W, Cc, _p(in_scale), _p(in_shift), in_act,
         _p(addend), _p(bn_mean), _p(bn_invstd), _p(bn_red), _p(bn_y), _dt(dy), _stream())
    return out


# ---------------------------------------------------------------------------------------------- K2
def gemm_bf16(a: Tensor, b: Tensor, M: int, N: int, K: int, out: Tensor, lda=None, ldb=None, ldc=None,
              col_scale=None, col_shift=None, act=ACT_NONE, addend=None, ld_addend=0, col_stats=None):
    """out[M,N] = epi(a[M,K] @ b[N,K]^T); a,b bf16 with K contiguous."""
    lda = K if lda is None else lda
    ldb = K if ldb is None else ldb
    ldc = N if ldc is None else ldc
    call("dlv3p_gemm_bf16", _p(a), lda, _p(b), ldb, _p(out), ldc, M, N, K, _dt(out), _p(col_scale), _p(col_shift),
         act, _p(addend), ld_addend, _p(col_stats), _stream())
    return out


def gemm_wgrad_bf16(x: Tensor, dy: Tensor, dw: Tensor, M: int, K: int, N: int, ldx=None, ldy=None, ldw=None):
    """dw[K,N] (fp32) += x[M,K]^T @ dy[M,N]"""
    call("dlv3p_gemm_wgrad_bf16", _p(x), K if ldx is None else ldx, _p(dy), N if ldy is None else ldy, _p(dw),
         N if ldw is None else ldw, M, K, N, _stream())
    return dw


def conv3x3_valid_supported(Cin: int, Cout: int) -> bool:
    """Shapes all three implicit-GEMM entry points accept (see include/dlv3p.h)."""
    return Cin % 8 == 0 and Cout == 64 and 64 < 3 * Cin <= 96


def conv3x3_valid_fwd(x: Tensor, wt: Tensor, out: Tensor, Cout: int, ldw=None, col_scale=None, col_shift=None,
                      act=ACT_NONE, col_stats=None):
    """out[N,H-2,W-2,Cout] = epi(conv3x3_valid(x[N,H,W,Cin])), implicit GEMM; wt bf16 [Cout, 9*Cin] K-major (pitch ldw)."""
    _chk(x, "x")
    N, H, W, Cin = x.shape
    call("dlv3p_conv3x3_valid_fwd_bf16", _p(x), _p(wt), 9 * Cin if ldw is None else ldw, _p(out), N, H, W, Cin, Cout,
         _p(col_scale), _p(col_shift), act, _p(col_stats), _stream())
    return out


def conv3x3_valid_dgrad(dy: Tensor, wd: Tensor, x_shape, Cout: int, out: Tensor):
    """out[N,H,W,Cin] = conv_transpose(dy[N,H-2,W-2,Cout]); wd bf16 [Cin, 9*Cout]."""
    _chk(dy, "dy")
    N, H, W, Cin = x_shape
    call("dlv3p_conv3x3_valid_dgrad_bf16", _p(dy), _p(wd), _p(out), N, H, W, Cin, Cout, _stream())
    return out


def conv3x3_valid_wgrad(x: Tensor, dy: Tensor, dw: Tensor, Cout: int):
    """dw (fp32 HWIO [3,3,Cin,Cout], accumulated) += x-windows^T dy."""
    _chk(x, "x"); _chk(dy, "dy")
    N, H, W, Cin = x.shape
    call("dlv3p_conv3x3_valid_wgrad_bf16", _p(x), _p(dy), _p(dw), N, H, W, Cin, Cout, _stream())
    return dw


def upsample_argmax(z: Tensor, fh: int, fw: int, labels: Tensor):
    """labels[N,H*fh,W*fw] (int32 or uint8) = argmax_c bilinear_upsample(z[N,H,W,C] fp32): the inference tail without the
    high-resolution logits / probabilities."""
    _chk(z, "z")
    if z.dtype != torch.float32 or labels.dtype not in (torch.int32, torch.uint8):
        raise ValueError("upsample_argmax: fp32 logits, int32 or uint8 labels")
    N, H, W, Cc = z.shape
    call("dlv3p_upsample_argmax", _p(z), _p(labels), labels.element_size(), N, H, W, Cc, fh, fw, _stream())
    return labels


def conv3x3_same_supported(Cin: int, Cout: int) -> bool:
    """Shapes the implicit-GEMM SAME convolution entry points accept (see include/dlv3p.h)."""
    return Cin % 8 == 0 and Cin >= 16 and Cout <= 256


def conv3x3_same_fwd(x: Tensor, wt: Tensor, out: Tensor, Cout: int, ldw=None, col_scale=None, col_shift=None,
                     act=ACT_NONE, col_stats=None):
    """out[N,H,W,Cout] (bf16 or fp32) = epi(conv3x3_same(x[N,H,W,Cin])); wt bf16 [Cout, 9*Cin] K-major (pitch ldw)."""
    _chk(x, "x")
    N, H, W, Cin = x.shape
    call("dlv3p_conv3x3_same_fwd_bf16", _p(x), _p(wt), 9 * Cin if ldw is None else ldw, _p(out), _dt(out), N, H, W, Cin,
         Cout, _p(col_scale), _p(col_shift), act, _p(col_stats), _stream())
    return out


def conv3x3_same_dgrad(dy: Tensor, ld_dy: int, wd: Tensor, x_shape, Cout: int, out: Tensor):
    """out[N,H,W,Cin] (bf16) = SAME conv_transpose(dy[N,H,W,:Cout] with channel pitch ld_dy); wd bf16 [Cin, 9*kp]."""
    _chk(dy, "dy")
    N, H, W, Cin = x_shape
    call("dlv3p_conv3x3_same_dgrad_bf16", _p(dy), ld_dy, _p(wd), _p(out), N, H, W, Cin, Cout, _stream())
    return out


def conv3x3_same_wgrad(x: Tensor, dy: Tensor, ld_dy: int, dw: Tensor, Cout: int):
    """dw (fp32 HWIO [3,3,Cin,Cout], accumulated) += shifted x windows^T dy."""
    _chk(x, "x"); _chk(dy, "dy")
    N, H, W, Cin = x.shape
    call("dlv3p_conv3x3_same_wgrad_bf16", _p(x), _p(dy), ld_dy, _p(dw), N, H, W, Cin, Cout, _stream())
    return dw


def gemm_simt(a: Tensor, sam: int, sak: int, b: Tensor, sbk: int, sbn: int, out: Tensor, ldc: int, M: int, N: int,
              K: int, col_scale=None, col_shift=None, act=ACT_NONE, addend=None, ld_addend=0, accumulate=False):
    call("dlv3p_gemm_simt", _p(a), sam, sak, _p(b), sbk, sbn, _p(out), ldc, M, N, K, _dt(a), _dt(out),
         _p(col_scale), _p(col_shift), act, _p(addend), ld_addend, int(accumulate), _stream())
    return out


def im2col3x3(x: Tensor, stride: int, dil: int, ho: int, wo: int, pt: int, pl: int, ld_col: int,
              out: Optional[Tensor] = None):
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((N * ho * wo, ld_col), dtype=x.dtype, device=x.device)
    call("dlv3p_im2col3x3", _p(x), _p(out), N, H, W, Cc, stride, dil, pt, pl, ho, wo, ld_col, _dt(x), _stream())
    return out


def col2im3x3(col: Tensor, x_shape, stride: int, dil: int, ho: int, wo: int, pt: int, pl: int, ld_col: int,
              addend=None, out: Optional[Tensor] = None):
    N, H, W, Cc = x_shape
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=col.dtype, device=col.device)
    call("dlv3p_col2im3x3", _p(col), _p(out), N, H, W, Cc, stride, dil, pt, pl, ho, wo, ld_col, _p(addend),
         _dt(col), _stream())
    return out


def subsample_fwd(x: Tensor, stride: int, out: Optional[Tensor] = None):
    N, H, W, Cc = x.shape
    ho, wo = -(-H // stride), -(-W // stride)
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=x.dtype, device=x.device)
    call("dlv3p_subsample_fwd", _p(x), _p(out), N, H, W, Cc, stride, ho, wo, _dt(x), _stream())
    return out


def subsample_bwd(dy: Tensor, x_shape, stride: int, addend=None, out: Optional[Tensor] = None):
    N, H, W, Cc = x_shape
    _, ho, wo, _ = dy.shape
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    call("dlv3p_subsample_bwd", _p(dy), _p(out), N, H, W, Cc, stride, ho, wo, _p(addend), _dt(dy), _stream())
    return out


def weight_prep(w: Tensor, K: int, N: int, wt: Tensor, ldt: int, wn: Optional[Tensor] = None, ldn: int = 0):
    call("dlv3p_weight_prep", _p(w), K, N, _p(wt), ldt, _p(wn), ldn, _stream())


def weight_prep_table(entries, device) -> Tensor:
    """Device table for weight_prep_batch from [(w, K, N, wt, ldt, wn, ldn), ...] (include/dlv3p.h layout)."""
    import numpy as np
    rec = np.zeros(len(entries), dtype=np.dtype([("w", "<u8"), ("wt", "<u8"), ("wn", "<u8"), ("ldt", "<i8"),
                                                  ("ldn", "<i8"), ("K", "<i4"), ("N", "<i4")]))
    for i, (w, K, N, wt, ldt, wn, ldn) in enumerate(entries):
        rec[i] = (w.data_ptr(), wt.data_ptr(), wn.data_ptr() if wn is not None else 0, ldt, ldn, K, N)
    return torch.from_numpy(rec.view(np.uint8).copy()).to(device)


def weight_prep_batch(table: Tensor, count: int, blocks_per_entry: int = 32):
    call("dlv3p_weight_prep_batch", _p(table), count, blocks_per_entry, _stream())


# ---------------------------------------------------------------------------------------------- K3
def bn_stats(y: Tensor, M: int, Cc: int, sums: Tensor, ld=None):
    call("dlv3p_bn_stats", _p(y), Cc if ld is None else ld, M, Cc, _p(sums), _dt(y), _stream())
    return sums


def bn_finalize(sums, gamma, beta, moving_mean, moving_var, Cc, count, eps, momentum, scale, shift, mean, invstd,
                update_moving=True):
    call("dlv3p_bn_finalize", _p(sums), _p(gamma), _p(beta), _p(moving_mean), _p(moving_var), Cc, float(count),
         eps, momentum, _p(scale), _p(shift), _p(mean), _p(invstd), int(update_moving), _stream())


def bn_train_apply(y, M, Cc, sums, gamma, beta, moving_mean, moving_var, count, eps, momentum, updates, act, out,
                   scale, shift, mean, invstd, addend=None, ld_out=None):
    """Training-mode BatchNormalization (+activation, +residual add) forward from the batch sums, in one launch.
    `ld_out`: row pitch of `out` (a channel slice of a Concatenate buffer)."""
    call("dlv3p_bn_train_apply", _p(y), Cc, _p(sums), _p(gamma), _p(beta), _p(moving_mean), _p(moving_var),
         float(count), eps, momentum, int(updates), act, _p(addend), Cc, _p(out), Cc if ld_out is None else ld_out, M, Cc,
         _p(scale), _p(shift), _p(mean), _p(invstd), _dt(y), _stream())
    return out


def bn_fold(gamma, beta, moving_mean, moving_var, Cc, eps, scale, shift):
    call("dlv3p_bn_fold", _p(gamma), _p(beta), _p(moving_mean), _p(moving_var), Cc, eps, _p(scale), _p(shift),
         _stream())


def affine_act(y: Tensor, M: int, Cc: int, out: Tensor, scale=None, shift=None, act=ACT_NONE, addend=None,
               ld_y=None, ld_out=None, ld_addend=None):
    call("dlv3p_affine_act", _p(y), Cc if ld_y is None else ld_y, _p(scale), _p(shift), act, _p(addend),
         Cc if ld_addend is None else ld_addend, _p(out), Cc if ld_out is None else ld_out, M, Cc, _dt(y), _stream())
    return out


def bn_bwd_reduce(dz, y, scale, shift, mean, invstd, act, M, Cc, red, ld_dz=None, ld_y=None):
    call("dlv3p_bn_bwd_reduce", _p(dz), Cc if ld_dz is None else ld_dz, _p(y), Cc if ld_y is None else ld_y,
         _p(scale), _p(shift), _p(mean), _p(invstd), act, M, Cc, _p(red), _dt(y), _stream())


def bn_bwd_apply(dz, y, scale, shift, mean, invstd, act, red, M, Cc, dy, ld_dz=None, ld_y=None, ld_dy=None):
    call("dlv3p_bn_bwd_apply", _p(dz), Cc if ld_dz is None else ld_dz, _p(y), Cc if ld_y is None else ld_y,
         _p(scale), _p(shift), _p(mean), _p(invstd), act, _p(red), M, Cc, _p(dy), Cc if ld_dy is None else ld_dy,
         _dt(y), _stream())


def act_bwd(dy: Tensor, x: Tensor, act: int, out: Tensor, addend=None):
    call("dlv3p_act_bwd", _p(dy), _p(x), _p(out), act, _p(addend), dy.numel(), _dt(dy), _stream())
    return out


def add(a: Tensor, b: Tensor, out: Tensor):
    call("dlv3p_add", _p(a), _p(b), _p(out), a.numel(), _dt(a), _stream())
    return out


def copy2d(x: Tensor, ld_x: int, y: Tensor, ld_y: int, M: int, Cc: int, addend=None, ld_addend=0, x_off=0, y_off=0):
    """Row-strided copy; x_off/y_off are element offsets into the base pointers (channel slices of a concat)."""
    esz = x.element_size()
    call("dlv3p_copy2d", x.data_ptr() + x_off * esz, ld_x, y.data_ptr() + y_off * esz, ld_y, M, Cc, _p(addend),
         ld_addend, _dt(x), _stream())


def maxpool3x3s2_fwd(x: Tensor, out: Optional[Tensor] = None, argmax: Optional[Tensor] = None, addend=None):
    N, H, W, Cc = x.shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=x.dtype, device=x.device)
    call("dlv3p_maxpool3x3s2_fwd", _p(x), _p(out), _p(argmax), N, H, W, Cc, pt, pl, ho, wo, _p(addend), _dt(x),
         _stream())
    return out


def maxpool3x3s2_bwd(dy: Tensor, argmax: Tensor, x_shape, addend=None, out: Optional[Tensor] = None):
    N, H, W, Cc = x_shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    call("dlv3p_maxpool3x3s2_bwd", _p(dy), _p(argmax), _p(out), N, H, W, Cc, pt, pl, ho, wo, _p(addend), _dt(dy),
         _stream())
    return out


def maxpool3x3s2_bn_fwd(x: Tensor, scale: Tensor, shift: Tensor, out: Tensor, ymax: Tensor, argmax: Tensor, addend=None):
    """out = maxpool3x3s2(scale*x + shift) (+ addend) without writing the BN output; ymax = raw x of the winners."""
    _chk(x, "x")
    N, H, W, Cc = x.shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    call("dlv3p_maxpool3x3s2_bn_fwd", _p(x), _p(scale), _p(shift), _p(out), _p(ymax), _p(argmax), N, H, W, Cc, pt, pl, ho,
         wo, _p(addend), _dt(x), _stream())
    return out


def maxpool3x3s2_bn_bwd(dy: Tensor, argmax: Tensor, x: Tensor, scale, mean, invstd, red, count, out: Tensor):
    """out = BN input gradient of the layer whose raw output x fed maxpool3x3s2_bn_fwd (pool backward + bn_bwd_apply)."""
    _chk(dy, "dy")
    N, H, W, Cc = x.shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    call("dlv3p_maxpool3x3s2_bn_bwd", _p(dy), _p(argmax), _p(x), _p(scale), _p(mean), _p(invstd), _p(red), float(count),
         _p(out), N, H, W, Cc, pt, pl, ho, wo, _dt(dy), _stream())
    return out


def avgpool_fwd(x: Tensor, k: int, out: Optional[Tensor] = None):
    N, H, W, Cc = x.shape
    ho, wo = H // k, W // k
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=x.dtype, device=x.device)
    call("dlv3p_avgpool_fwd", _p(x), _p(out), N, H, W, Cc, k, ho, wo, _dt(x), _stream())
    return out


def avgpool_bwd(dy: Tensor, x_shape, k: int, addend=None, out: Optional[Tensor] = None):
    N, H, W, Cc = x_shape
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    call("dlv3p_avgpool_bwd", _p(dy), _p(out), N, H, W, Cc, k, H // k, W // k, _p(addend), _dt(dy), _stream())
    return out


def bilinear_fwd(x: Tensor, fh: int, fw: int, out: Optional[Tensor] = None, ld_x=None, ld_y=None, C=None,
                 y_off=0, out_dtype=None):
    N, H, W, Cx = x.shape
    Cc = Cx if C is None else C
    if out is None:
        out = torch.empty((N, H * fh, W * fw, Cc), dtype=out_dtype or x.dtype, device=x.device)
    call("dlv3p_bilinear_fwd", _p(x), Cx if ld_x is None else ld_x, out.data_ptr() + y_off * out.element_size(),
         Cc if ld_y is None else ld_y, N, H, W, Cc, fh, fw, _dt(x), _dt(out), _stream())
    return out


def bilinear_bwd(dy: Tensor, x_shape, fh: int, fw: int, out: Optional[Tensor] = None, addend=None, ld_dy=None,
                 ld_dx=None, dy_off=0, out_dtype=None):
    N, H, W, Cc = x_shape
    if out is None:
        out = torch.empty((N, H, W, Cc), dtype=out_dtype or dy.dtype, device=dy.device)
    call("dlv3p_bilinear_bwd", dy.data_ptr() + dy_off * dy.element_size(), Cc if ld_dy is None else ld_dy, _p(out),
         Cc if ld_dx is None else ld_dx, N, H, W, Cc, fh, fw, _p(addend), _dt(dy), _dt(out), _stream())
    return out


def softmax_cbloss_fwd(z, labels, pw, nw, eps, P, Cc, loss_sum, probs=None):
    call("dlv3p_softmax_cbloss_fwd", _p(z), _p(labels), _p(pw), _p(nw), eps, P, Cc, _p(loss_sum), _p(probs), _stream())


def softmax_cbloss_bwd(z, labels, pw, nw, eps, P, Cc, grad_scale, dz):
    call("dlv3p_softmax_cbloss_bwd", _p(z), _p(labels), _p(pw), _p(nw), eps, P, Cc, grad_scale, _p(dz), _stream())


def upsample_softmax_cbloss_fwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, loss_sum):
    call("dlv3p_upsample_softmax_cbloss_fwd", _p(zl), _p(labels), _p(pw), _p(nw), eps, N, H, W, Cc, f, _p(loss_sum),
         _stream())


def upsample_softmax_cbloss_bwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, grad_scale, dzl):
    call("dlv3p_upsample_softmax_cbloss_bwd", _p(zl), _p(labels), _p(pw), _p(nw), eps, N, H, W, Cc, f, grad_scale,
         _p(dzl), _stream())


def upsample_softmax_cbloss_fwd_bwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, grad_scale, loss_sum, dzl):
    """loss_sum += sum of per-pixel losses and dzl += grad_scale * d(loss)/d(zl) in one pass (caller zeroes both)."""
    call("dlv3p_upsample_softmax_cbloss_fwd_bwd", _p(zl), _p(labels), _p(pw), _p(nw), eps, N, H, W, Cc, f, grad_scale,
         _p(loss_sum), _p(dzl), _stream())


def softmax_argmax(z, P, Cc, probs=None, labels=None):
    call("dlv3p_softmax_argmax", _p(z), P, Cc, _p(probs), _p(labels), _stream())


def cbloss_dense_fwd(y_true, y_pred, pw, nw, eps, P, Cc, loss_sum):
    call("dlv3p_cbloss_dense_fwd", _p(y_true), _p(y_pred), _p(pw), _p(nw), eps, P, Cc, _p(loss_sum), _stream())


def cbloss_dense_bwd(y_true, y_pred, pw, nw, eps, P, Cc, grad_scale, dy_pred):
    call("dlv3p_cbloss_dense_bwd", _p(y_true), _p(y_pred), _p(pw), _p(nw), eps, P, Cc, grad_scale, _p(dy_pred),
         _stream())


def softmax_bwd(p, dp, P, Cc, dz):
    call("dlv3p_softmax_bwd", _p(p), _p(dp), P, Cc, _p(dz), _stream())


def confusion_matrix(y_true, y_pred, P, Cc, cm):
    call("dlv3p_confusion_matrix", _p(y_true), _p(y_pred), P, Cc, _p(cm), _stream())


def dropout(x: Tensor, rate: float, seed: int, out: Tensor, addend=None, seed_offset=None):
    call("dlv3p_dropout", _p(x), _p(out), x.numel(), rate, seed, _p(seed_offset), _p(addend), _dt(x), _stream())
    return out


def adam(w, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale=1.0, l2=0.0, w_off=0):
    o = w_off * 4
    call("dlv3p_adam", w.data_ptr() + o, g.data_ptr() + o, m.data_ptr() + o, v.data_ptr() + o, n, lr_t, beta1, beta2,
         eps, grad_scale, l2, _stream())


def sumsq(w, n, out, w_off=0):
    call("dlv3p_sumsq", w.data_ptr() + w_off * 4, n, _p(out), _stream())


def cast(x: Tensor, out: Tensor):
    call("dlv3p_cast", _p(x), _dt(x), _p(out), _dt(out), x.numel(), _stream())
    return out


def cast2d(x: Tensor, ld_x: int, out: Tensor, ld_y: int, M: int, Cc: int):
    call("dlv3p_cast2d", _p(x), ld_x, _dt(x), _p(out), ld_y, _dt(out), M, Cc, _stream())
    return out

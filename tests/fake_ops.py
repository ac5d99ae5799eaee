"""TEST DOUBLE for deeplabv3plus_keras_b200.ops — plain torch-CPU restatements of the kernel semantics declared in
include/dlv3p.h, with the same Python signatures.

Purpose: let the CPU test-suite (`-m "not gpu"`) exercise the HOST logic of engine.Plan — graph flattening,
fusion patterns, buffer planning, the hand-written backward schedule and gradient accumulation — against the
oracle, without a GPU.  It is installed by the `cpu_engine` fixture (tests/conftest_engine.py) via monkeypatching
and is never importable from the product package: the product's ops module still raises on CPU tensors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from deeplabv3plus_keras_b200 import ops as real

ACT_NONE, ACT_RELU, ACT_RELU6 = real.ACT_NONE, real.ACT_RELU, real.ACT_RELU6
same_pad, valid_out, conv_geometry = real.same_pad, real.valid_out, real.conv_geometry
FAKE = True


def _act(v, act):
    if act == ACT_RELU:
        return torch.relu(v)
    if act == ACT_RELU6:
        return torch.clamp(v, 0, 6)
    return v


def _mask(v, act):
    if act == ACT_RELU:
        return (v > 0).to(v.dtype)
    if act == ACT_RELU6:
        return ((v > 0) & (v < 6)).to(v.dtype)
    return torch.ones_like(v)


def _rows(t, M, ld, C, off=0):
    """[M,C] view of a buffer with row pitch ld starting `off` elements into it."""
    if t.dim() == 2 and tuple(t.shape) == (M, C) and t.stride() == (ld, 1) and off == 0:
        return t                                  # already the strided slice (a channel slice of a concat buffer)
    flat = t.reshape(-1)
    return flat[off:off + (M - 1) * ld + C].as_strided((M, C), (ld, 1))


def _pre(x, in_scale, in_shift, in_act):
    v = x.float()
    if in_scale is not None:
        v = v * in_scale + in_shift
    return _act(v, in_act)


def _dw(xf, w, stride, dil, Ho, Wo, pt, pl):
    N, H, W, C = xf.shape
    pb = max((Ho - 1) * stride + 2 * dil[0] + 1 - H - pt, 0)
    pr = max((Wo - 1) * stride + 2 * dil[1] + 1 - W - pl, 0)
    xp = F.pad(xf.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xp, w.permute(2, 0, 1).unsqueeze(1), stride=stride, dilation=dil, groups=C)
    return y[:, :, :Ho, :Wo].permute(0, 2, 3, 1)


def dwconv3x3_fwd(x, w, stride=1, dil=(1, 1), padding="same", in_scale=None, in_shift=None, in_act=ACT_NONE, out=None,
                  pad=None):
    N, H, W, C = x.shape
    Ho, Wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    y = _dw(_pre(x, in_scale, in_shift, in_act), w, stride, dil, Ho, Wo, pt, pl)
    if out is None:
        return y.to(x.dtype).contiguous()
    out.copy_(y)
    return out


def dwconv3x3_fwd_epi(x, w, out_scale, out_shift, out_act, stride=1, dil=(1, 1), padding="same", out=None, pad=None):
    N, H, W, C = x.shape
    Ho, Wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    y = _act(_dw(x.float(), w, stride, dil, Ho, Wo, pt, pl) * out_scale + out_shift, out_act)
    if out is None:
        return y.to(x.dtype).contiguous()
    out.copy_(y)
    return out


def dwconv3x3_dgrad(dy, w, x_shape, stride=1, dil=(1, 1), padding="same", x_pre=None, in_scale=None, in_shift=None,
                    in_act=ACT_NONE, addend=None, out=None, pad=None):
    N, H, W, C = x_shape
    Ho, Wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    z = torch.zeros((N, H, W, C), dtype=torch.float32, requires_grad=True)
    _dw(z, w, stride, dil, Ho, Wo, pt, pl).backward(dy.float())
    g = z.grad
    if in_act != ACT_NONE:
        pre = x_pre.float()
        if in_scale is not None:
            pre = pre * in_scale + in_shift
        g = g * _mask(pre, in_act)
    if addend is not None:
        g = g + addend.float()
    out.copy_(g)
    return out


def dwconv3x3_bn_fwd(x, w, bn_sums, gamma, beta, moving_mean, moving_var, count, eps, momentum, updates, in_act, scale,
                     shift, mean, invstd, out=None, pad=None):
    Cc = x.shape[3]
    for u in range(max(updates, 1)):
        bn_finalize(bn_sums, gamma, beta, moving_mean, moving_var, Cc, count, eps, momentum, scale, shift, mean, invstd,
                    update_moving=u < updates)
    return dwconv3x3_fwd(x, w, 1, (1, 1), in_scale=scale, in_shift=shift, in_act=in_act, out=out, pad=pad)


def dwconv3x3_dgrad_bnred(dy, w, x_shape, x_pre, in_scale, in_shift, in_act, bn_mean, bn_invstd, bn_red, out=None,
                          pad=None):
    if out is None:
        out = torch.empty(tuple(x_shape), dtype=dy.dtype)
    dwconv3x3_dgrad(dy, w, x_shape, 1, (1, 1), x_pre=x_pre, in_scale=in_scale, in_shift=in_shift, in_act=in_act,
                    out=out, pad=pad)
    Cc = x_shape[3]
    g = out.float().reshape(-1, Cc)
    Y = x_pre.float().reshape(-1, Cc)
    bn_red[:Cc] += g.sum(0)
    bn_red[Cc:2 * Cc] += (g * (Y - bn_mean) * bn_invstd).sum(0)
    return out


def dwconv3x3_wgrad(x, dy, dw, stride=1, dil=(1, 1), padding="same", in_scale=None, in_shift=None, in_act=ACT_NONE,
                    pad=None):
    N, H, W, C = x.shape
    Ho, Wo, pt, pl = pad if pad is not None else conv_geometry(H, W, 3, stride, dil, padding)
    wz = torch.zeros((3, 3, C), dtype=torch.float32, requires_grad=True)
    _dw(_pre(x, in_scale, in_shift, in_act), wz, stride, dil, Ho, Wo, pt, pl).backward(dy.float())
    dw.add_(wz.grad)
    return dw


def dwconv3x3_bwd(dy, x, w, dw, in_scale=None, in_shift=None, in_act=ACT_NONE, addend=None, bn_mean=None,
                  bn_invstd=None, bn_red=None, bn_y=None, out=None):
    pad = conv_geometry(x.shape[1], x.shape[2], 3, 1, (1, 1), "same")
    if bn_red is not None and bn_y is None:
        out = dwconv3x3_dgrad_bnred(dy, w, tuple(x.shape), x, in_scale, in_shift, in_act, bn_mean, bn_invstd, bn_red,
                                    out=out, pad=pad)
    else:
        out = dwconv3x3_dgrad(dy, w, tuple(x.shape), 1, (1, 1), x_pre=x if in_act != ACT_NONE else None,
                              in_scale=in_scale, in_shift=in_shift, in_act=in_act, addend=addend, out=out, pad=pad)
    if bn_y is not None:
        Cc = x.shape[3]
        g = out.float().reshape(-1, Cc)
        bn_red[:Cc] += g.sum(0)
        bn_red[Cc:2 * Cc] += (g * (bn_y.float().reshape(-1, Cc) - bn_mean) * bn_invstd).sum(0)
    dwconv3x3_wgrad(x, dy, dw, 1, (1, 1), in_scale=in_scale, in_shift=in_shift, in_act=in_act, pad=pad)
    return out


def _epi(acc, col_scale, col_shift, act, addend):
    if col_scale is not None:
        acc = acc * col_scale + col_shift
    acc = _act(acc, act)
    if addend is not None:
        acc = acc + addend
    return acc


def gemm_bf16(a, b, M, N, K, out, lda=None, ldb=None, ldc=None, col_scale=None, col_shift=None, act=ACT_NONE,
              addend=None, ld_addend=0, col_stats=None):
    lda, ldb, ldc = lda or K, ldb or K, ldc or N
    A, B = _rows(a, M, lda, K).float(), _rows(b, N, ldb, K).float()
    acc = A @ B.t()
    if col_stats is not None:
        col_stats[:N] += acc.sum(0)
        col_stats[N:2 * N] += (acc * acc).sum(0)
    add = _rows(addend, M, ld_addend, N).float() if addend is not None else None
    _rows(out, M, ldc, N).copy_(_epi(acc, col_scale, col_shift, act, add))
    return out


def gemm_wgrad_bf16(x, dy, dw, M, K, N, ldx=None, ldy=None, ldw=None):
    X, G = _rows(x, M, ldx or K, K).float(), _rows(dy, M, ldy or N, N).float()
    _rows(dw, K, ldw or N, N).add_(X.t() @ G)
    return dw


def conv3x3_valid_supported(Cin, Cout):
    return Cin % 8 == 0 and Cout == 64 and 64 < 3 * Cin <= 96


def conv3x3_valid_fwd(x, wt, out, Cout, ldw=None, col_scale=None, col_shift=None, act=ACT_NONE, col_stats=None):
    N, H, W, Cin = x.shape
    ldw = ldw or 9 * Cin
    w = _rows(wt, Cout, ldw, 9 * Cin).float().reshape(Cout, 3, 3, Cin)                 # [o, i, j, c]
    acc = F.conv2d(x.float().permute(0, 3, 1, 2), w.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).reshape(-1, Cout)
    if col_stats is not None:
        col_stats[:Cout] += acc.sum(0)
        col_stats[Cout:2 * Cout] += (acc * acc).sum(0)
    out.view(-1, Cout).copy_(_epi(acc, col_scale, col_shift, act, None))
    return out


def conv3x3_valid_dgrad(dy, wd, x_shape, Cout, out):
    N, H, W, Cin = x_shape
    w = wd.float().view(Cin, 3, 3, Cout).permute(1, 2, 0, 3)                  # [3,3,Cin,Cout]
    g = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    out.copy_(g)
    return out


def conv3x3_valid_wgrad(x, dy, dw, Cout):
    N, H, W, Cin = x.shape
    wz = torch.zeros((Cout, Cin, 3, 3), dtype=torch.float32, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wz).backward(dy.float().permute(0, 3, 1, 2))
    dw.view(3, 3, Cin, Cout).add_(wz.grad.permute(2, 3, 1, 0))
    return dw


def upsample_argmax(z, fh, fw, labels):
    labels.copy_(_resize(z.float(), fh, fw).argmax(-1).to(labels.dtype))
    return labels


def conv3x3_same_supported(Cin, Cout):
    return Cin % 8 == 0 and Cin >= 16 and Cout <= 256


def conv3x3_same_fwd(x, wt, out, Cout, ldw=None, col_scale=None, col_shift=None, act=ACT_NONE, col_stats=None):
    N, H, W, Cin = x.shape
    ldw = ldw or 9 * Cin
    w = _rows(wt, Cout, ldw, 9 * Cin).float().reshape(Cout, 3, 3, Cin)                 # [o, i, j, c]
    acc = F.conv2d(x.float().permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    if col_stats is not None:
        col_stats[:Cout] += acc.sum(0)
        col_stats[Cout:2 * Cout] += (acc * acc).sum(0)
    out.view(-1, Cout).copy_(_epi(acc, col_scale, col_shift, act, None))
    return out


def conv3x3_same_dgrad(dy, ld_dy, wd, x_shape, Cout, out):
    N, H, W, Cin = x_shape
    kp = wd.shape[1] // 9
    w = wd.float().view(Cin, 9, kp)[:, :, :Cout].reshape(Cin, 3, 3, Cout)          # [c, i, j, o]
    g = dy.float().reshape(N, H, W, ld_dy)[..., :Cout]
    dx = F.conv_transpose2d(g.permute(0, 3, 1, 2), w.permute(3, 0, 1, 2), padding=1).permute(0, 2, 3, 1)
    out.copy_(dx)
    return out


def conv3x3_same_wgrad(x, dy, ld_dy, dw, Cout):
    N, H, W, Cin = x.shape
    g = dy.float().reshape(N, H, W, ld_dy)[..., :Cout]
    wz = torch.zeros((Cout, Cin, 3, 3), dtype=torch.float32, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wz, padding=1).backward(g.permute(0, 3, 1, 2))
    dw.view(3, 3, Cin, Cout).add_(wz.grad.permute(2, 3, 1, 0))
    return dw


def gemm_simt(a, sam, sak, b, sbk, sbn, out, ldc, M, N, K, col_scale=None, col_shift=None, act=ACT_NONE, addend=None,
              ld_addend=0, accumulate=False):
    A = a.reshape(-1).as_strided((M, K), (sam, sak)).float()
    B = b.reshape(-1).as_strided((K, N), (sbk, sbn)).float()
    add = _rows(addend, M, ld_addend, N).float() if addend is not None else None
    r = _epi(A @ B, col_scale, col_shift, act, add)
    o = _rows(out, M, ldc, N)
    o.copy_(r + o.float() if accumulate else r)
    return out


def im2col3x3(x, stride, dil, ho, wo, pt, pl, ld_col, out=None):
    N, H, W, C = x.shape
    pb = max((ho - 1) * stride + 2 * dil + 1 - H - pt, 0)
    pr = max((wo - 1) * stride + 2 * dil + 1 - W - pl, 0)
    xp = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    cols = [xp[:, :, i * dil:i * dil + stride * (ho - 1) + 1:stride, j * dil:j * dil + stride * (wo - 1) + 1:stride]
            .permute(0, 2, 3, 1) for i in range(3) for j in range(3)]
    col = torch.cat(cols, dim=-1).reshape(N * ho * wo, 9 * C)
    if out is None:
        out = torch.zeros((N * ho * wo, ld_col), dtype=x.dtype)
    out.zero_()
    out[:, :9 * C] = col
    return out


def col2im3x3(col, x_shape, stride, dil, ho, wo, pt, pl, ld_col, addend=None, out=None):
    N, H, W, C = x_shape
    z = torch.zeros(x_shape, dtype=torch.float32, requires_grad=True)
    c = im2col3x3(z, stride, dil, ho, wo, pt, pl, 9 * C, out=None) if False else None
    pb = max((ho - 1) * stride + 2 * dil + 1 - H - pt, 0)
    pr = max((wo - 1) * stride + 2 * dil + 1 - W - pl, 0)
    xp = F.pad(z.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    cols = [xp[:, :, i * dil:i * dil + stride * (ho - 1) + 1:stride, j * dil:j * dil + stride * (wo - 1) + 1:stride]
            .permute(0, 2, 3, 1) for i in range(3) for j in range(3)]
    torch.cat(cols, dim=-1).reshape(N * ho * wo, 9 * C).backward(_rows(col, N * ho * wo, ld_col, 9 * C).float())
    g = z.grad + (addend.float() if addend is not None else 0)
    out.copy_(g)
    return out


def subsample_fwd(x, stride, out=None):
    y = x[:, ::stride, ::stride]
    if out is None:
        return y.contiguous()
    out.copy_(y)
    return out


def subsample_bwd(dy, x_shape, stride, addend=None, out=None):
    g = torch.zeros(x_shape, dtype=torch.float32)
    g[:, ::stride, ::stride] = dy.float()
    if addend is not None:
        g = g + addend.float()
    out.copy_(g)
    return out


def weight_prep(w, K, N, wt, ldt, wn=None, ldn=0):
    wt[:, :K] = w.t().to(wt.dtype)
    if wn is not None:
        wn[:, :N] = w.to(wn.dtype)


def bn_stats(y, M, Cc, sums, ld=None):
    Y = _rows(y, M, ld or Cc, Cc).float()
    sums[:Cc] += Y.sum(0)
    sums[Cc:2 * Cc] += (Y * Y).sum(0)
    return sums


def bn_finalize(sums, gamma, beta, moving_mean, moving_var, Cc, count, eps, momentum, scale, shift, mean, invstd,
                update_moving=True):
    m = sums[:Cc].double() / count
    var = torch.clamp(sums[Cc:2 * Cc].double() / count - m * m, min=0)
    is_ = 1.0 / torch.sqrt(var + eps)
    g = gamma.double() if gamma is not None else torch.ones(Cc, dtype=torch.float64)
    b = beta.double() if beta is not None else torch.zeros(Cc, dtype=torch.float64)
    scale.copy_(g * is_)
    shift.copy_(b - m * g * is_)
    mean.copy_(m)
    invstd.copy_(is_)
    if update_moving:
        unb = var * count / (count - 1) if count > 1 else var
        moving_mean.copy_(momentum * moving_mean + (1 - momentum) * m.float())
        moving_var.copy_(momentum * moving_var + (1 - momentum) * unb.float())


def bn_train_apply(y, M, Cc, sums, gamma, beta, moving_mean, moving_var, count, eps, momentum, updates, act, out,
                   scale, shift, mean, invstd, addend=None, ld_out=None):
    for u in range(max(updates, 1)):
        bn_finalize(sums, gamma, beta, moving_mean, moving_var, Cc, count, eps, momentum, scale, shift, mean, invstd,
                    update_moving=u < updates)
    return affine_act(y, M, Cc, out, scale, shift, act, addend=addend, ld_out=ld_out)


def weight_prep_table(entries, device):
    return list(entries)


def weight_prep_batch(table, count, blocks_per_entry=32):
    for w, K, N, wt, ldt, wn, ldn in table:
        weight_prep(w, K, N, wt, ldt, wn, ldn)


def bn_fold(gamma, beta, moving_mean, moving_var, Cc, eps, scale, shift):
    s = (gamma if gamma is not None else 1.0) / torch.sqrt(moving_var + eps)
    scale.copy_(s)
    shift.copy_((beta if beta is not None else 0.0) - moving_mean * s)


def affine_act(y, M, Cc, out, scale=None, shift=None, act=ACT_NONE, addend=None, ld_y=None, ld_out=None,
               ld_addend=None):
    v = _rows(y, M, ld_y or Cc, Cc).float()
    if scale is not None:
        v = v * scale + shift
    v = _act(v, act)
    if addend is not None:
        v = v + _rows(addend, M, ld_addend or Cc, Cc).float()
    _rows(out, M, ld_out or Cc, Cc).copy_(v)
    return out


def _bn_g(dz, y, scale, shift, act, M, Cc, ld_dz, ld_y):
    Y = _rows(y, M, ld_y or Cc, Cc).float()
    return _rows(dz, M, ld_dz or Cc, Cc).float() * _mask(Y * scale + shift, act), Y


def bn_bwd_reduce(dz, y, scale, shift, mean, invstd, act, M, Cc, red, ld_dz=None, ld_y=None):
    g, Y = _bn_g(dz, y, scale, shift, act, M, Cc, ld_dz, ld_y)
    red[:Cc] += g.sum(0)
    red[Cc:2 * Cc] += (g * (Y - mean) * invstd).sum(0)


def bn_bwd_apply(dz, y, scale, shift, mean, invstd, act, red, M, Cc, dy, ld_dz=None, ld_y=None, ld_dy=None):
    g, Y = _bn_g(dz, y, scale, shift, act, M, Cc, ld_dz, ld_y)
    if mean is not None:
        xh = (Y - mean) * invstd
        r = scale * (g - red[:Cc] / M - xh * red[Cc:2 * Cc] / M)
    else:
        r = scale * g
    _rows(dy, M, ld_dy or Cc, Cc).copy_(r)


def act_bwd(dy, x, act, out, addend=None):
    g = dy.float() * _mask(x.float(), act)
    if addend is not None:
        g = g + addend.float()
    out.copy_(g)
    return out


def add(a, b, out):
    out.copy_(a.float() + b.float())
    return out


def copy2d(x, ld_x, y, ld_y, M, Cc, addend=None, ld_addend=0, x_off=0, y_off=0):
    v = _rows(x, M, ld_x, Cc, x_off).float()
    if addend is not None:
        v = v + _rows(addend, M, ld_addend, Cc).float()
    _rows(y, M, ld_y, Cc, y_off).copy_(v)


def maxpool3x3s2_fwd(x, out=None, argmax=None, addend=None):
    N, H, W, C = x.shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    pb, pr = (ho - 1) * 2 + 3 - H - pt, (wo - 1) * 2 + 3 - W - pl
    xp = F.pad(x.float().permute(0, 3, 1, 2), (pl, max(pr, 0), pt, max(pb, 0)), value=float("-inf"))
    win = xp.unfold(2, 3, 2).unfold(3, 3, 2)[:, :, :ho, :wo].reshape(N, C, ho, wo, 9)
    val, idx = win.max(dim=-1)          # torch returns the first maximum
    y = val.permute(0, 2, 3, 1)
    if argmax is not None:
        argmax.copy_(idx.permute(0, 2, 3, 1).to(torch.uint8))
    if addend is not None:
        y = y + addend.float()
    if out is None:
        return y.to(x.dtype).contiguous()
    out.copy_(y)
    return out


def maxpool3x3s2_bn_fwd(x, scale, shift, out, ymax, argmax, addend=None):
    z = x.float() * scale + shift
    maxpool3x3s2_fwd(z, out=out, argmax=argmax, addend=addend)
    N, H, W, C = x.shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    xp = F.pad(x.float(), (0, 0, pl, 2, pt, 2))
    am = argmax.long()
    n_i, h_i, w_i, c_i = torch.meshgrid(torch.arange(N), torch.arange(ho), torch.arange(wo), torch.arange(C), indexing="ij")
    ymax.copy_(xp[n_i, h_i * 2 + am // 3, w_i * 2 + am % 3, c_i])
    return out


def maxpool3x3s2_bn_bwd(dy, argmax, x, scale, mean, invstd, red, count, out):
    Cc = x.shape[3]
    g = torch.empty(tuple(x.shape), dtype=torch.float32)
    maxpool3x3s2_bwd(dy, argmax, tuple(x.shape), out=g)
    xh = (x.float() - mean) * invstd
    out.copy_(scale * (g - red[:Cc] / count - xh * red[Cc:2 * Cc] / count))
    return out


def maxpool3x3s2_bwd(dy, argmax, x_shape, addend=None, out=None):
    N, H, W, C = x_shape
    ho, pt = same_pad(H, 3, 2)
    wo, pl = same_pad(W, 3, 2)
    g = torch.zeros((N, H + 4, W + 4, C), dtype=torch.float32)
    am = argmax.long()
    n_i, h_i, w_i, c_i = torch.meshgrid(torch.arange(N), torch.arange(ho), torch.arange(wo), torch.arange(C),
                                        indexing="ij")
    hi = h_i * 2 - pt + am // 3 + 2
    wi = w_i * 2 - pl + am % 3 + 2
    g.index_put_((n_i, hi, wi, c_i), dy.float(), accumulate=True)
    g = g[:, 2:2 + H, 2:2 + W]
    if addend is not None:
        g = g + addend.float()
    out.copy_(g)
    return out


def avgpool_fwd(x, k, out=None):
    y = F.avg_pool2d(x.float().permute(0, 3, 1, 2), k, k).permute(0, 2, 3, 1)
    out.copy_(y)
    return out


def avgpool_bwd(dy, x_shape, k, addend=None, out=None):
    N, H, W, C = x_shape
    g = torch.zeros(x_shape, dtype=torch.float32)
    up = dy.float().repeat_interleave(k, 1).repeat_interleave(k, 2) / (k * k)
    g[:, :up.shape[1], :up.shape[2]] = up
    if addend is not None:
        g = g + addend.float()
    out.copy_(g)
    return out


def _resize(x, fh, fw):
    return F.interpolate(x.permute(0, 3, 1, 2), scale_factor=(fh, fw), mode="bilinear",
                         align_corners=False).permute(0, 2, 3, 1)


def bilinear_fwd(x, fh, fw, out=None, ld_x=None, ld_y=None, C=None, y_off=0, out_dtype=None):
    y = _resize(x.float(), fh, fw)
    if out is None:
        return y.to(out_dtype or x.dtype).contiguous()
    if ld_y is not None and ld_y != y.shape[3]:
        # channel-slice write into a wider (concat) buffer: rows of ld_y elements, slice starts y_off elements in
        out.view(-1, ld_y)[:, y_off:y_off + y.shape[3]].copy_(y.reshape(-1, y.shape[3]))
        return out
    out.copy_(y)
    return out


def bilinear_bwd(dy, x_shape, fh, fw, out=None, addend=None, ld_dy=None, ld_dx=None, dy_off=0, out_dtype=None):
    z = torch.zeros(x_shape, dtype=torch.float32, requires_grad=True)
    if dy.dim() == 2:                                   # [M, C] slice view of a concat gradient (ld_dy = its row pitch)
        dy = dy.reshape(x_shape[0], x_shape[1] * fh, x_shape[2] * fw, x_shape[3])
    _resize(z, fh, fw).backward(dy.float())
    g = z.grad + (addend.float() if addend is not None else 0)
    out.copy_(g)
    return out


def _cb(p, y, pw, nw, eps):
    return -(pw * y * torch.log(p + eps) + nw * (1 - y) * torch.log(1 - p + eps))


def softmax_cbloss_fwd(z, labels, pw, nw, eps, P, Cc, loss_sum, probs=None):
    p = torch.softmax(z.reshape(P, Cc).float(), -1)
    y = F.one_hot(labels.reshape(P).long(), Cc).float()
    loss_sum += _cb(p, y, pw, nw, eps).sum()
    if probs is not None:
        probs.copy_(p.view_as(probs))


def softmax_cbloss_bwd(z, labels, pw, nw, eps, P, Cc, grad_scale, dz):
    zz = z.reshape(P, Cc).float().clone().requires_grad_(True)
    y = F.one_hot(labels.reshape(P).long(), Cc).float()
    (_cb(torch.softmax(zz, -1), y, pw, nw, eps).sum() * grad_scale).backward()
    dz.copy_(zz.grad.view_as(dz))


def upsample_softmax_cbloss_fwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, loss_sum):
    zh = _resize(zl.float(), f, f)
    softmax_cbloss_fwd(zh, labels, pw, nw, eps, N * H * f * W * f, Cc, loss_sum)


def upsample_softmax_cbloss_bwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, grad_scale, dzl):
    zz = zl.float().clone().requires_grad_(True)
    P = N * H * f * W * f
    p = torch.softmax(_resize(zz, f, f).reshape(P, Cc), -1)
    y = F.one_hot(labels.reshape(P).long(), Cc).float()
    (_cb(p, y, pw, nw, eps).sum() * grad_scale).backward()
    dzl += zz.grad


def upsample_softmax_cbloss_fwd_bwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, grad_scale, loss_sum, dzl):
    upsample_softmax_cbloss_fwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, loss_sum)
    upsample_softmax_cbloss_bwd(zl, labels, pw, nw, eps, N, H, W, Cc, f, grad_scale, dzl)


def softmax_argmax(z, P, Cc, probs=None, labels=None):
    zz = z.reshape(P, Cc).float()
    if probs is not None:
        probs.copy_(torch.softmax(zz, -1).view_as(probs))
    if labels is not None:
        labels.copy_(zz.argmax(-1).view_as(labels))


def dropout(x, rate, seed, out, addend=None, seed_offset=None):
    g = torch.Generator().manual_seed(int(seed) + (int(seed_offset.item()) if seed_offset is not None else 0))
    keep = (torch.rand(x.shape, generator=g) >= rate).float() / (1 - rate)
    v = x.float() * keep
    if addend is not None:
        v = v + addend.float()
    out.copy_(v)
    return out


def adam(w, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale=1.0, l2=0.0, w_off=0):
    s = slice(w_off, w_off + n)
    ge = g[s] * grad_scale + 2 * l2 * w[s]
    m[s] = beta1 * m[s] + (1 - beta1) * ge
    v[s] = beta2 * v[s] + (1 - beta2) * ge * ge
    w[s] -= lr_t * m[s] / (v[s].sqrt() + eps)


def sumsq(w, n, out, w_off=0):
    out += (w[w_off:w_off + n] ** 2).sum()


def cast2d(x, ld_x, out, ld_y, M, Cc):
    _rows(out, M, ld_y, Cc).copy_(_rows(x, M, ld_x, Cc).float())
    return out

"""Per-kernel parity: every C-ABI entry point against the oracle (oracle/tf_ops.py, fp64 on CPU) on seeded inputs.

Tolerances: fp32 storage -> rtol 1e-4 (fp32 accumulation order); bf16 storage -> the oracle is evaluated on the
bf16-rounded inputs in fp64 and the result may differ by one bf16 rounding of the output (2^-8 relative) plus fp32
accumulation noise; tensor-core GEMMs: bf16 products are exact in fp32, so only accumulation order differs.
"""
import math

import pytest
import torch

from oracle import tf_ops as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def ops():
    from deeplabv3plus_keras_b200 import ops as _ops
    return _ops


def rnd(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float64) * scale).to(dtype)


def tol(dtype):
    return (2e-2, 2e-2) if dtype == torch.bfloat16 else (2e-4, 2e-4)


def check(name, got, want, rtol, atol):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    bad = err > bound
    if bad.any():
        idx = torch.nonzero(bad)[0].tolist()
        raise AssertionError(
            f"{name}: {int(bad.sum())}/{bad.numel()} mismatches; max abs err {err.max():.4e} "
            f"(max |want| {want.abs().max():.4e}); first bad idx {idx}: got {got[tuple(idx)]:.6e} want {want[tuple(idx)]:.6e}")


DW_CASES = [
    # N, H, W, C, stride, dil, padding
    (2, 17, 19, 64, 1, (1, 1), "same"),
    (1, 33, 33, 96, 1, (18, 15), "same"),
    (1, 33, 33, 32, 1, (6, 21), "same"),
    (2, 32, 32, 256, 1, (12, 12), "same"),
    (1, 64, 64, 64, 1, (36, 36), "same"),
    (2, 30, 31, 728, 1, (1, 1), "same"),
    (1, 65, 65, 144, 2, (1, 1), "mnv2"),
    (1, 64, 48, 32, 2, (1, 1), "mnv2"),
    (1, 9, 9, 8, 1, (6, 3), "same"),
]


def dw_pad(H, W, stride, dil, padding):
    o = ops()
    if padding == "mnv2":
        # ZeroPadding2D(correct_pad) + VALID == SAME arithmetic for a 3x3 kernel
        return o.conv_geometry(H, W, 3, stride, dil, "same")
    return o.conv_geometry(H, W, 3, stride, dil, padding)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", DW_CASES)
@pytest.mark.parametrize("prologue", [False, True])
def test_dwconv3x3_fwd_dgrad_wgrad(case, dtype, prologue):
    o = ops()
    N, H, W, C, stride, dil, padding = case
    x = rnd((N, H, W, C), dtype, 1)
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    sc = (rnd((C,), torch.float32, 3, 0.2) + 1.0) if prologue else None
    sh = rnd((C,), torch.float32, 4, 0.3) if prologue else None
    act = o.ACT_RELU if prologue else o.ACT_NONE
    pad = dw_pad(H, W, stride, dil, padding)
    ho, wo, pt, pl = pad

    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    z = xr
    if prologue:
        z = torch.relu(z * sc.double() + sh.double())
    z.retain_grad()
    y_ref = O.depthwise_conv2d(z, wr.view(3, 3, C, 1), stride, "same", dil)
    assert y_ref.shape[1:3] == (ho, wo)
    gy = rnd(tuple(y_ref.shape), dtype, 5)
    y_ref.backward(gy.double())

    xd, wd = x.to(DEV), w.to(DEV)
    scd = sc.to(DEV) if prologue else None
    shd = sh.to(DEV) if prologue else None
    y = o.dwconv3x3_fwd(xd, wd, stride, dil, in_scale=scd, in_shift=shd, in_act=act, pad=pad)
    rt, at = tol(dtype)
    check("dw fwd", y, y_ref, rt, at * 4)

    # dgrad returns the gradient w.r.t. the activation input z_pre = sc*x+sh (mask applied, no scale factor)
    gyd = gy.to(DEV)
    add = rnd((N, H, W, C), dtype, 6)
    dx = o.dwconv3x3_dgrad(gyd, wd, (N, H, W, C), stride, dil, x_pre=xd if prologue else None, in_scale=scd,
                           in_shift=shd, in_act=act, addend=add.to(DEV), pad=pad)
    dz_ref = z.grad
    if prologue:
        pre = x.double() * sc.double() + sh.double()
        dz_ref = dz_ref * (pre > 0).double()
    check("dw dgrad", dx, dz_ref + add.double(), rt, at * 4)

    dw = torch.zeros((3, 3, C), dtype=torch.float32, device=DEV)
    o.dwconv3x3_wgrad(xd, gyd, dw, stride, dil, in_scale=scd, in_shift=shd, in_act=act, pad=pad)
    check("dw wgrad", dw, wr.grad, 2e-3, 2e-3 * math.sqrt(N * ho * wo))


@pytest.mark.parametrize("case", [c for c in DW_CASES if c[5] == (1, 1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_dwconv3x3_fwd_output_epilogue(case, dtype):
    """Inference DepthwiseConv2D -> BatchNormalization -> ReLU6 in one launch (dlv3p_dwconv3x3_fwd_epi: folded BN + clamp
    in the convolution's epilogue; TMA kernel for the dense stride-1 bf16 case, direct kernel for stride 2 / fp32) against
    the fp64 restatement and against the two-launch form (dwconv3x3_fwd + affine_act)."""
    o = ops()
    N, H, W, C, stride, dil, padding = case
    x = rnd((N, H, W, C), dtype, 1)
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    sc = rnd((C,), torch.float32, 3, 0.3) + 1.0
    sh = rnd((C,), torch.float32, 4, 0.5) + 1.0
    pad = dw_pad(H, W, stride, dil, padding)
    for act, code in (("relu6", o.ACT_RELU6), ("relu", o.ACT_RELU), ("none", o.ACT_NONE)):
        got = o.dwconv3x3_fwd_epi(x.to(DEV), w.to(DEV), sc.to(DEV), sh.to(DEV), code, stride, dil, pad=pad)
        xx = x.double()
        if padding == "mnv2":
            xx = O.zero_pad2d(xx, ((1 - (1 - H % 2), 1), (1 - (1 - W % 2), 1)))
            conv = O.depthwise_conv2d(xx, w.double().view(3, 3, C, 1), stride, "valid", dil)
        else:
            conv = O.depthwise_conv2d(xx, w.double().view(3, 3, C, 1), stride, padding, dil)
        z = conv * sc.double() + sh.double()
        ref = torch.clamp(z, 0, 6) if act == "relu6" else (torch.clamp_min(z, 0) if act == "relu" else z)
        rt, at = tol(dtype)
        check(f"dw fwd epi {act}", got, ref, rt, at * 4)
        two = o.dwconv3x3_fwd(x.to(DEV), w.to(DEV), stride, dil, pad=pad)
        M = two.numel() // C
        o.affine_act(two, M, C, two, sc.to(DEV), sh.to(DEV), code)
        check(f"dw fwd epi vs two launches {act}", got, two, 2e-2 if dtype == torch.bfloat16 else 1e-5, at * 4)


@pytest.mark.parametrize("case", [(2, 17, 19, 64), (2, 30, 31, 728), (1, 32, 32, 736), (3, 40, 70, 128)])
@pytest.mark.parametrize("act", ["relu", "relu6"])
def test_dwconv3x3_dgrad_bnred(case, act):
    """Input gradient of a dense-tap depthwise conv + the BatchNormalization-backward reductions of the producing
    layer in one launch, against the fp64 restatement (masked conv-transpose, sum g, sum g*xhat)."""
    o = ops()
    N, H, W, C = case
    dtype = torch.bfloat16
    code = o.ACT_RELU if act == "relu" else o.ACT_RELU6
    y = rnd((N, H, W, C), dtype, 1, 2.0)                    # raw conv output of the producing layer
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    sc = rnd((C,), torch.float32, 3, 0.2) + 1.0
    sh = rnd((C,), torch.float32, 4, 0.5) + (2.0 if act == "relu6" else 0.0)
    mean = rnd((C,), torch.float32, 7, 0.3)
    invstd = rnd((C,), torch.float32, 8, 0.1).abs() + 0.6
    gy = rnd((N, H, W, C), dtype, 5)
    pad = dw_pad(H, W, 1, (1, 1), "same")

    zr = torch.zeros((N, H, W, C), dtype=torch.float64, requires_grad=True)
    O.depthwise_conv2d(zr, w.double().view(3, 3, C, 1), 1, "same", (1, 1)).backward(gy.double())
    pre = y.double() * sc.double() + sh.double()
    mask = (pre > 0) if act == "relu" else ((pre > 0) & (pre < 6))
    g_ref = zr.grad * mask.double()
    red_ref = torch.cat([g_ref.sum((0, 1, 2)), (g_ref * (y.double() - mean.double()) * invstd.double()).sum((0, 1, 2))])

    red = torch.zeros(2 * C, dtype=torch.float32, device=DEV)
    dx = o.dwconv3x3_dgrad_bnred(gy.to(DEV), w.to(DEV), (N, H, W, C), y.to(DEV), sc.to(DEV), sh.to(DEV), code,
                                 mean.to(DEV), invstd.to(DEV), red, pad=pad)
    rt, at = tol(dtype)
    check("dgrad_bnred dx", dx, g_ref, rt, at * 4)
    check("dgrad_bnred red", red, red_ref, 2e-3, 2e-3 * math.sqrt(N * H * W))
    # accumulates (the caller zeroes once per step): a second call doubles the sums
    o.dwconv3x3_dgrad_bnred(gy.to(DEV), w.to(DEV), (N, H, W, C), y.to(DEV), sc.to(DEV), sh.to(DEV), code,
                            mean.to(DEV), invstd.to(DEV), red, pad=pad)
    check("dgrad_bnred red x2", red, 2 * red_ref, 2e-3, 4e-3 * math.sqrt(N * H * W))
    with pytest.raises(ValueError):
        o.dwconv3x3_dgrad_bnred(gy.float().to(DEV), w.to(DEV), (N, H, W, C), y.float().to(DEV), sc.to(DEV), sh.to(DEV),
                                code, mean.to(DEV), invstd.to(DEV), red, pad=pad)       # fp32: unfused path only


@pytest.mark.parametrize("case", [(2, 17, 19, 64), (2, 30, 31, 728), (1, 32, 32, 736), (3, 40, 70, 128), (1, 9, 33, 48),
                                  (2, 8, 32, 104)])
@pytest.mark.parametrize("mode", ["bnred_relu", "bnred_relu6", "affine_add", "preact", "preact_add", "plain", "plain_add"])
def test_dwconv3x3_bwd_fused(case, mode):
    """dlv3p_dwconv3x3_bwd: input gradient + filter gradient (+ BN-backward reductions) of a stride-1 SAME depthwise
    conv in one launch, against the fp64 restatement on the same bf16 operands, for every operand combination the
    engine emits: BN+ReLU(6) on load with reductions, BN+ReLU on load with a gradient addend, pre-activation ReLU
    with / without addend, no activation.  Shapes straddle the 8 x 32 tile and the 48-channel block."""
    o = ops()
    N, H, W, C = case
    bf = torch.bfloat16
    xs = rnd((N, H, W, C), bf, 1, 2.0)                      # what the forward read before its prologue
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    gy = rnd((N, H, W, C), bf, 5)
    affine = mode.startswith("bnred") or mode == "affine_add"
    act = {"bnred_relu": o.ACT_RELU, "bnred_relu6": o.ACT_RELU6, "affine_add": o.ACT_RELU, "preact": o.ACT_RELU,
           "preact_add": o.ACT_RELU, "plain": o.ACT_NONE, "plain_add": o.ACT_NONE}[mode]
    sc = (rnd((C,), torch.float32, 3, 0.2) + 1.0) if affine else None
    sh = (rnd((C,), torch.float32, 4, 0.5) + (2.0 if act == o.ACT_RELU6 else 0.0)) if affine else None
    add = rnd((N, H, W, C), bf, 6) if mode.endswith("add") else None
    stats = mode.startswith("bnred")
    mean = rnd((C,), torch.float32, 7, 0.3)
    invstd = rnd((C,), torch.float32, 8, 0.1).abs() + 0.6

    pre = xs.double() * sc.double() + sh.double() if affine else xs.double()
    if act == o.ACT_RELU:
        xa, mask = pre.clamp(min=0), (pre > 0).double()
    elif act == o.ACT_RELU6:
        xa, mask = pre.clamp(0, 6), ((pre > 0) & (pre < 6)).double()
    else:
        xa, mask = pre, torch.ones_like(pre)
    xr = xa.clone().requires_grad_(True)
    wr = w.double().view(3, 3, C, 1).clone().requires_grad_(True)
    O.depthwise_conv2d(xr, wr, 1, "same", (1, 1)).backward(gy.double())
    g_ref = xr.grad * mask
    dx_ref = g_ref + (add.double() if add is not None else 0.0)
    dw_ref = wr.grad.view(3, 3, C)
    red_ref = torch.cat([g_ref.sum((0, 1, 2)), (g_ref * (xs.double() - mean.double()) * invstd.double()).sum((0, 1, 2))])

    dw = torch.zeros((3, 3, C), dtype=torch.float32, device=DEV)
    red = torch.zeros(2 * C, dtype=torch.float32, device=DEV) if stats else None
    d = lambda t: None if t is None else t.to(DEV)
    run = lambda: o.dwconv3x3_bwd(gy.to(DEV), xs.to(DEV), w.to(DEV), dw, in_scale=d(sc), in_shift=d(sh), in_act=act,
                                  addend=d(add), bn_mean=d(mean) if stats else None,
                                  bn_invstd=d(invstd) if stats else None, bn_red=red)
    dx = run()
    rt, at = tol(bf)
    check(f"dw bwd dx {mode}", dx, dx_ref, rt, at * 4)
    n_sum = N * H * W
    check(f"dw bwd dw {mode}", dw, dw_ref, 2e-3, 2e-3 * math.sqrt(n_sum) * 2.0)
    if stats:
        check(f"dw bwd red {mode}", red, red_ref, 2e-3, 2e-3 * math.sqrt(n_sum))
    run()                                                    # dw and bn_red accumulate
    check(f"dw bwd dw x2 {mode}", dw, 2 * dw_ref, 2e-3, 4e-3 * math.sqrt(n_sum) * 2.0)
    if stats:
        check(f"dw bwd red x2 {mode}", red, 2 * red_ref, 2e-3, 4e-3 * math.sqrt(n_sum))
    # the fused launch equals the two separate entry points on the same operands
    dw2 = torch.zeros_like(dw)
    pad = dw_pad(H, W, 1, (1, 1), "same")
    dx2 = o.dwconv3x3_dgrad(gy.to(DEV), w.to(DEV), (N, H, W, C), 1, (1, 1), x_pre=xs.to(DEV) if act != o.ACT_NONE else None,
                            in_scale=d(sc), in_shift=d(sh), in_act=act, addend=d(add), pad=pad)
    o.dwconv3x3_wgrad(xs.to(DEV), gy.to(DEV), dw2, 1, (1, 1), in_scale=d(sc), in_shift=d(sh), in_act=act, pad=pad)
    check(f"dw bwd dx vs dgrad {mode}", dx, dx2.double().cpu(), 8e-3, 1e-3)      # one bf16 ulp: different summation order
    check(f"dw bwd dw vs wgrad {mode}", dw, 2 * dw2.double().cpu(), 1e-4, 1e-3 * math.sqrt(n_sum))


@pytest.mark.parametrize("case", [(2, 30, 31, 728), (3, 40, 70, 128), (1, 32, 32, 736)])
@pytest.mark.parametrize("with_add", [False, True])
def test_dwconv3x3_bwd_fused_bn_y(case, with_add):
    """dlv3p_dwconv3x3_bwd with bn_y: besides dx and dw, the BatchNormalization-backward reductions of the layer whose
    raw output is bn_y, taken on the FINAL gradient (addend included) — what the engine asks of the reader of an
    Xception block output."""
    o = ops()
    N, H, W, C = case
    bf = torch.bfloat16
    xs = rnd((N, H, W, C), bf, 1, 2.0)
    by = rnd((N, H, W, C), bf, 9, 1.5) + 0.7
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    gy = rnd((N, H, W, C), bf, 5)
    add = rnd((N, H, W, C), bf, 6) if with_add else None
    mean = rnd((C,), torch.float32, 7, 0.3) + 0.7
    invstd = rnd((C,), torch.float32, 8, 0.1).abs() + 0.6
    xr = xs.double().clamp(min=0).requires_grad_(True)
    wr = w.double().view(3, 3, C, 1).clone().requires_grad_(True)
    O.depthwise_conv2d(xr, wr, 1, "same", (1, 1)).backward(gy.double())
    dx_ref = xr.grad * (xs.double() > 0) + (add.double() if with_add else 0.0)
    dw = torch.zeros((3, 3, C), dtype=torch.float32, device=DEV)
    red = torch.zeros(2 * C, dtype=torch.float32, device=DEV)
    dx = o.dwconv3x3_bwd(gy.to(DEV), xs.to(DEV), w.to(DEV), dw, in_act=o.ACT_RELU, addend=None if add is None else add.to(DEV),
                         bn_mean=mean.to(DEV), bn_invstd=invstd.to(DEV), bn_red=red, bn_y=by.to(DEV))
    rt, at = tol(bf)
    check("dw bwd bn_y dx", dx, dx_ref, rt, at * 4)
    check("dw bwd bn_y dw", dw, wr.grad.view(3, 3, C), 2e-3, 4e-3 * math.sqrt(N * H * W))
    # the reductions are those of the STORED (bf16-rounded) gradient, as dlv3p_bn_bwd_reduce would compute them from dx
    g = dx.double().cpu()
    red_ref = torch.cat([g.sum((0, 1, 2)), (g * (by.double() - mean.double()) * invstd.double()).sum((0, 1, 2))])
    check("dw bwd bn_y red", red, red_ref, 4e-3, 4e-3 * math.sqrt(N * H * W) * 2)


def test_dwconv3x3_bwd_fused_fp32_route():
    """fp32 (parity mode) takes the separate kernels behind the same entry point."""
    o = ops()
    N, H, W, C = 2, 11, 13, 32
    xs = rnd((N, H, W, C), torch.float32, 1, 2.0)
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    gy = rnd((N, H, W, C), torch.float32, 5)
    xr = xs.double().clamp(min=0).requires_grad_(True)
    wr = w.double().view(3, 3, C, 1).clone().requires_grad_(True)
    O.depthwise_conv2d(xr, wr, 1, "same", (1, 1)).backward(gy.double())
    dw = torch.zeros((3, 3, C), dtype=torch.float32, device=DEV)
    dx = o.dwconv3x3_bwd(gy.to(DEV), xs.to(DEV), w.to(DEV), dw, in_act=o.ACT_RELU)
    check("dw bwd fp32 dx", dx, xr.grad * (xs.double() > 0), 1e-5, 1e-5)
    check("dw bwd fp32 dw", dw, wr.grad.view(3, 3, C), 1e-4, 1e-4)


@pytest.mark.parametrize("ratio", [5.0, 40.0])
def test_dwconv3x3_dgrad_bnred_large_mean_over_std(ratio):
    """The fused kernel finishes sum g*xhat from sum g*y and sum g (it has no second pass over y).  With a channel
    mean `ratio` standard deviations away from zero the two terms cancel to 1/ratio of their size: the reductions must
    still hold 1e-2 against fp64 on identical bf16 inputs (review finding: cancellation in the fp32 atomics)."""
    o = ops()
    N, H, W, C = 4, 33, 33, 128
    g = torch.Generator().manual_seed(11)
    std = 0.5
    mean = (torch.rand(C, generator=g) * 2 - 1).sign() * ratio * std
    y = (torch.randn((N, H, W, C), generator=g) * std + mean).to(torch.bfloat16)
    mean_b = y.double().mean((0, 1, 2))
    invstd = 1.0 / torch.sqrt(y.double().var((0, 1, 2), unbiased=False) + 1e-3)
    w = rnd((3, 3, C), torch.float32, 2, 0.3)
    gamma = rnd((C,), torch.float32, 3, 0.2) + 1.0
    sc = (gamma.double() * invstd).float()
    sh = (0.1 - mean_b * sc.double()).float()
    gy = rnd((N, H, W, C), torch.bfloat16, 5)
    pad = dw_pad(H, W, 1, (1, 1), "same")
    zr = torch.zeros((N, H, W, C), dtype=torch.float64, requires_grad=True)
    O.depthwise_conv2d(zr, w.double().view(3, 3, C, 1), 1, "same", (1, 1)).backward(gy.double())
    pre = y.double() * sc.double() + sh.double()
    g_ref = zr.grad * (pre > 0).double()
    s1 = g_ref.sum((0, 1, 2))
    s2 = (g_ref * (y.double() - mean_b) * invstd).sum((0, 1, 2))
    red = torch.zeros(2 * C, dtype=torch.float32, device=DEV)
    o.dwconv3x3_dgrad_bnred(gy.to(DEV), w.to(DEV), (N, H, W, C), y.to(DEV), sc.to(DEV), sh.to(DEV), o.ACT_RELU,
                            mean_b.float().to(DEV), invstd.float().to(DEV), red, pad=pad)
    got = red.double().cpu()
    scale2 = float(s2.abs().mean())
    assert float((got[:C] - s1).abs().max()) <= 1e-2 * float(s1.abs().mean())
    assert float((got[C:] - s2).abs().max()) <= 1e-2 * scale2, (float((got[C:] - s2).abs().max()), scale2)


@pytest.mark.parametrize("case", [(2, 16, 32, 256, 21), (1, 9, 131, 304, 21), (2, 12, 260, 48, 19), (1, 5, 3, 64, 21),
                                  (2, 7, 70, 96, 64), (1, 33, 33, 256, 32), (1, 4, 200, 136, 128)])
def test_conv3x3_same_implicit_gemm(case):
    """Implicit-GEMM 3x3 SAME stride-1 convolution (tap-shifted rank-4 TMA windows; outside the image = zero = the SAME
    padding; no im2col matrix) — the logits layer (21 / 19 classes, fp32 output, Cin = 256 or 304 after boundary
    refinement) and BatchNormalization-carrying shapes (bf16 output + column statistics): forward, input gradient and
    filter gradient against the fp64 convolution on the same bf16 operands.  Widths straddle the 128-pixel M tile and the
    64-pixel reduction block; channel counts straddle the 64-channel k-block (304 = 4.75 blocks) and the 128-row tile."""
    o = ops()
    N, H, W, Cin, Cout = case
    bf = torch.bfloat16
    x = rnd((N, H, W, Cin), bf, 1)
    w = rnd((3, 3, Cin, Cout), torch.float32, 2, 0.1).to(bf)             # HWIO, bf16-representable
    Kp = (9 * Cin + 7) // 8 * 8
    wt = torch.zeros((Cout, Kp), dtype=bf)
    wt[:, :9 * Cin] = w.reshape(9 * Cin, Cout).t()                       # [Cout, 9*Cin] K-major, as the im2col GEMM's B
    kp = (Cout + 63) // 64 * 64
    wd = torch.zeros((Cin, 9, kp), dtype=bf)
    wd[:, :, :Cout] = w.reshape(9, Cin, Cout).permute(1, 0, 2)
    wd = wd.reshape(Cin, 9 * kp).contiguous()

    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    y_ref = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    ld = (Cout + 7) // 8 * 8
    gy = rnd((N, H, W, Cout), bf, 5)
    y_ref.backward(gy.double())
    gy_pad = torch.zeros((N, H, W, ld), dtype=bf)
    gy_pad[..., :Cout] = gy

    # forward, fp32 output (logits) ...
    y32 = torch.full((N, H, W, Cout), float("nan"), dtype=torch.float32, device=DEV)
    o.conv3x3_same_fwd(x.to(DEV), wt.to(DEV), y32, Cout, ldw=Kp)
    check("same fwd f32", y32, y_ref.detach(), 2e-3, 2e-3 * math.sqrt(9 * Cin) * 0.1)
    # ... and bf16 output with column statistics and an epilogue (a convolution followed by BatchNormalization)
    if Cout % 8 == 0:
        yb = torch.full((N, H, W, Cout), float("nan"), dtype=bf, device=DEV)
        stats = torch.zeros(2 * Cout, dtype=torch.float32, device=DEV)
        o.conv3x3_same_fwd(x.to(DEV), wt.to(DEV), yb, Cout, ldw=Kp, col_stats=stats)
        check("same fwd bf16", yb, y_ref.detach(), 1e-2, 1e-2 * math.sqrt(9 * Cin) * 0.1)
        ys = yb.double().cpu().reshape(-1, Cout) if Cout > 32 else y_ref.detach().reshape(-1, Cout)
        check("same fwd stats", stats, torch.cat([ys.sum(0), (ys * ys).sum(0)]), 5e-3, 5e-3 * math.sqrt(N * H * W) * 2)

    dx = torch.full((N, H, W, Cin), float("nan"), dtype=bf, device=DEV)
    o.conv3x3_same_dgrad(gy_pad.to(DEV), ld, wd.to(DEV), (N, H, W, Cin), Cout, dx)
    check("same dgrad", dx, xr.grad, 1e-2, 1e-2 * math.sqrt(9 * Cout) * 0.1)

    dw = torch.zeros((3, 3, Cin, Cout), dtype=torch.float32, device=DEV)
    o.conv3x3_same_wgrad(x.to(DEV), gy_pad.to(DEV), ld, dw, Cout)
    check("same wgrad", dw, wr.grad, 2e-3, 2e-3 * math.sqrt(N * H * W))
    o.conv3x3_same_wgrad(x.to(DEV), gy_pad.to(DEV), ld, dw, Cout)          # accumulates
    check("same wgrad x2", dw, 2 * wr.grad, 2e-3, 4e-3 * math.sqrt(N * H * W))


@pytest.mark.parametrize("case", [(2, 20, 37, 32, 64), (1, 9, 131, 32, 64), (2, 12, 260, 40, 128), (1, 5, 3, 32, 64)])
def test_conv3x3_valid_implicit_gemm(case):
    """Implicit-GEMM 3x3 VALID stride-1 convolution (overlapping-row tensor maps, no im2col matrix): forward with BN
    statistics, input gradient and filter gradient against the fp64 convolution on the same bf16 operands.  Widths
    straddle the 128-pixel M tile / 64-pixel reduction block (ragged row ends are clipped by the rank-3 C map)."""
    o = ops()
    N, H, W, Cin, Cout = case
    bf = torch.bfloat16
    x = rnd((N, H, W, Cin), bf, 1)
    w = rnd((3, 3, Cin, Cout), torch.float32, 2, 0.1).to(bf)             # HWIO, bf16-representable
    wt = w.reshape(9 * Cin, Cout).t().contiguous()                       # [Cout, 9*Cin] K-major, as the im2col GEMM's B
    wd = w.reshape(9, Cin, Cout).permute(1, 0, 2).reshape(Cin, 9 * Cout).contiguous()
    Ho, Wo = H - 2, W - 2

    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    y_ref = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    gy = rnd((N, Ho, Wo, Cout), bf, 5)
    y_ref.backward(gy.double())

    y = torch.full((N, Ho, Wo, Cout), float("nan"), dtype=bf, device=DEV)
    stats = torch.zeros(2 * Cout, dtype=torch.float32, device=DEV)
    o.conv3x3_valid_fwd(x.to(DEV), wt.to(DEV), y, Cout, col_stats=stats)
    check("implicit conv fwd", y, y_ref, 1e-2, 2e-2)
    yq = y.double().cpu().reshape(-1, Cout)                              # statistics are those of the stored tile
    check("implicit conv stats", stats, torch.cat([yq.sum(0), (yq * yq).sum(0)]), 2e-3, 2e-3 * math.sqrt(N * Ho * Wo))
    if o.conv3x3_valid_supported(Cin, Cout):
        dx = torch.full((N, H, W, Cin), float("nan"), dtype=bf, device=DEV)
        o.conv3x3_valid_dgrad(gy.to(DEV), wd.to(DEV), (N, H, W, Cin), Cout, dx)
        check("implicit conv dgrad", dx, xr.grad, 1e-2, 2e-2 * math.sqrt(9 * Cout / 64))
        dw = torch.zeros((3, 3, Cin, Cout), dtype=torch.float32, device=DEV)
        o.conv3x3_valid_wgrad(x.to(DEV), gy.to(DEV), dw, Cout)
        check("implicit conv wgrad", dw, wr.grad, 2e-3, 2e-3 * math.sqrt(N * Ho * Wo))
        o.conv3x3_valid_wgrad(x.to(DEV), gy.to(DEV), dw, Cout)           # accumulates
        check("implicit conv wgrad x2", dw, 2 * wr.grad, 2e-3, 4e-3 * math.sqrt(N * Ho * Wo))
    # fused inference epilogue (folded BN + ReLU)
    sc = (rnd((Cout,), torch.float32, 3, 0.2) + 1.0)
    sh = rnd((Cout,), torch.float32, 4, 0.3)
    o.conv3x3_valid_fwd(x.to(DEV), wt.to(DEV), y, Cout, col_scale=sc.to(DEV), col_shift=sh.to(DEV), act=o.ACT_RELU)
    check("implicit conv fwd+bn+relu", y, torch.relu(y_ref.detach() * sc.double() + sh.double()), 1e-2, 3e-2)


def test_conv3x3_valid_unsupported_shapes():
    o = ops()
    x = torch.zeros((1, 8, 8, 16), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(ValueError):
        o.conv3x3_valid_fwd(x, torch.zeros((32, 144), dtype=torch.bfloat16, device=DEV),
                            torch.zeros((1, 6, 6, 32), dtype=torch.bfloat16, device=DEV), 32)    # Cout < 64
    with pytest.raises(ValueError):
        o.conv3x3_valid_wgrad(x, torch.zeros((1, 6, 6, 64), dtype=torch.bfloat16, device=DEV),
                              torch.zeros((3, 3, 16, 64), device=DEV), 64)                       # 3*Cin <= 64


@pytest.mark.parametrize("case", [(2, 17, 19, 64), (2, 30, 31, 736), (1, 33, 70, 128)])
@pytest.mark.parametrize("updates", [1, 2])
def test_dwconv3x3_bn_fwd_equals_finalize_plus_conv(case, updates):
    """Depthwise forward with the producing layer's training-mode BatchNormalization finished inside the kernel
    (dlv3p_bn_finalize folded in) == dlv3p_bn_finalize followed by dlv3p_dwconv3x3_fwd with in_scale/in_shift: same
    output, same published scale/shift/mean/invstd, same moving statistics."""
    o = ops()
    N, H, W, C = case
    bf = torch.bfloat16
    y = (rnd((N, H, W, C), bf, 1, 1.5) + 0.3).to(bf).to(DEV)
    w = rnd((3, 3, C), torch.float32, 2, 0.3).to(DEV)
    gamma = (rnd((C,), torch.float32, 3, 0.2) + 1.0).to(DEV)
    beta = rnd((C,), torch.float32, 4, 0.3).to(DEV)
    M = N * H * W
    sums = torch.zeros(2 * C, dtype=torch.float32, device=DEV)
    o.bn_stats(y, M, C, sums)

    def fresh():
        return (torch.full((C,), 0.25, device=DEV), torch.full((C,), 2.0, device=DEV),
                *(torch.empty(C, device=DEV) for _ in range(4)))
    mm_a, mv_a, sc_a, sh_a, mu_a, is_a = fresh()
    for u in range(updates):
        o.bn_finalize(sums, gamma, beta, mm_a, mv_a, C, M, 1e-3, 0.99, sc_a, sh_a, mu_a, is_a, True)
    want = o.dwconv3x3_fwd(y, w, 1, (1, 1), in_scale=sc_a, in_shift=sh_a, in_act=o.ACT_RELU)
    mm_b, mv_b, sc_b, sh_b, mu_b, is_b = fresh()
    got = o.dwconv3x3_bn_fwd(y, w, sums, gamma, beta, mm_b, mv_b, M, 1e-3, 0.99, updates, o.ACT_RELU, sc_b, sh_b, mu_b, is_b)
    torch.cuda.synchronize()
    # same formulas in two kernels: equal up to fp32 contraction choices of the compiler
    for name, a, b in (("scale", sc_a, sc_b), ("shift", sh_a, sh_b), ("mean", mu_a, mu_b), ("invstd", is_a, is_b),
                       ("moving_mean", mm_a, mm_b), ("moving_var", mv_a, mv_b)):
        check("bn_fwd " + name, b, a, 2e-6, 2e-6)
    # outputs: bf16 roundings of fp32 values that differ by a few ulp(fp32) can land one bf16 ulp apart
    check("bn_fwd out", got, want, 1e-2, 1e-2)
    assert (got.float() != want.float()).float().mean().item() < 1e-2


@pytest.mark.parametrize("case", [(2, 30, 31, 64), (1, 64, 64, 736), (2, 127, 127, 256), (1, 7, 9, 8)])
@pytest.mark.parametrize("with_add", [False, True])
def test_maxpool_bn_fused_fwd_bwd(case, with_add):
    """MaxPooling2D(3, s2, SAME) fused with the BatchNormalization that feeds it: forward == maxpool(scale*x+shift)
    (+addend) with the raw winner values kept; backward == max-pool backward followed by bn_bwd_apply, in one pass.
    Negative scales exercise the sign-flipped compares."""
    o = ops()
    N, H, W, C = case
    bf = torch.bfloat16
    x = rnd((N, H, W, C), bf, 1, 1.5)
    sc = rnd((C,), torch.float32, 3, 0.7) + 0.3                    # a good share of negative scales
    sh = rnd((C,), torch.float32, 4, 0.3)
    ho, pt = o.same_pad(H, 3, 2)
    wo, pl = o.same_pad(W, 3, 2)
    add = rnd((N, ho, wo, C), bf, 6) if with_add else None
    z = x.double() * sc.double() + sh.double()                      # fp64 BN output (never rounded in the fused path)
    zp = torch.nn.functional.pad(z.permute(0, 3, 1, 2), (pl, 2, pt, 2), value=float("-inf"))
    win = zp.unfold(2, 3, 2).unfold(3, 3, 2)[:, :, :ho, :wo].reshape(N, C, ho, wo, 9)
    val, idx = win.max(dim=-1)
    want = val.permute(0, 2, 3, 1) + (add.double() if with_add else 0.0)

    y = torch.empty((N, ho, wo, C), dtype=bf, device=DEV)
    ymax = torch.empty_like(y)
    am = torch.empty((N, ho, wo, C), dtype=torch.uint8, device=DEV)
    o.maxpool3x3s2_bn_fwd(x.to(DEV), sc.to(DEV), sh.to(DEV), y, ymax, am, addend=add.to(DEV) if with_add else None)
    check("maxpool_bn fwd", y, want, 1e-2, 1e-2)
    # the stored winner is the raw x under the stored argmax, and BN of it is the pooled value
    xp = torch.nn.functional.pad(x.float(), (0, 0, pl, 2, pt, 2))
    n_i, h_i, w_i, c_i = torch.meshgrid(torch.arange(N), torch.arange(ho), torch.arange(wo), torch.arange(C), indexing="ij")
    amc = am.cpu().long()
    assert torch.equal(ymax.float().cpu(), xp[n_i, h_i * 2 + amc // 3, w_i * 2 + amc % 3, c_i])
    check("maxpool_bn ymax->value", ymax.double().cpu() * sc.double() + sh.double(), val.permute(0, 2, 3, 1), 1e-6, 1e-6)

    # backward against the fp64 restatement on the same argmax: the pooled reductions are EXACT sums of the window
    # gradients (the two-step path first rounds the routed gradient to bf16), then pool backward + BN input gradient
    gy = rnd((N, ho, wo, C), bf, 5)
    mean, invstd = rnd((C,), torch.float32, 7, 0.3), rnd((C,), torch.float32, 8, 0.1).abs() + 0.6
    M = N * H * W
    ymc = ymax.double().cpu()
    red_ref = torch.cat([gy.double().sum((0, 1, 2)),
                         (gy.double() * (ymc - mean.double()) * invstd.double()).sum((0, 1, 2))])
    red = torch.zeros(2 * C, dtype=torch.float32, device=DEV)
    o.bn_bwd_reduce(gy.to(DEV), ymax, sc.to(DEV), sh.to(DEV), mean.to(DEV), invstd.to(DEV), o.ACT_NONE, N * ho * wo, C, red)
    check("pooled reductions", red, red_ref, 2e-3, 2e-3 * math.sqrt(N * ho * wo))
    g64 = torch.zeros((N, H + 4, W + 4, C), dtype=torch.float64)
    g64.index_put_((n_i, h_i * 2 - pt + amc // 3 + 2, w_i * 2 - pl + amc % 3 + 2, c_i), gy.double(), accumulate=True)
    g64 = g64[:, 2:2 + H, 2:2 + W]
    xh = (x.double() - mean.double()) * invstd.double()
    r64 = red.double().cpu()
    dy_ref = sc.double() * (g64 - r64[:C] / M - xh * r64[C:] / M)
    dy = torch.empty((N, H, W, C), dtype=bf, device=DEV)
    o.maxpool3x3s2_bn_bwd(gy.to(DEV), am, x.to(DEV), sc.to(DEV), mean.to(DEV), invstd.to(DEV), red, M, dy)
    # a pixel that wins several windows sums their gradients in packed bf16 before the BN map (as dlv3p_maxpool3x3s2_bwd
    # does): when terms of opposite sign cancel, the error is a bf16 ulp of the TERMS, not of the small result
    err = (dy.double().cpu() - dy_ref).abs()
    bound = 2e-2 + 2e-2 * dy_ref.abs()
    assert (err > bound).double().mean().item() < 1e-5, float((err > bound).double().mean())
    assert err.max().item() < 2.0 ** -6 * float(gy.abs().max()) * float(sc.abs().max()) * 4 + 2e-2, float(err.max())
    with pytest.raises(ValueError):
        o.maxpool3x3s2_bn_bwd(gy.float().to(DEV), am, x.float().to(DEV), sc.to(DEV), mean.to(DEV), invstd.to(DEV), red, M,
                              dy.float())                                   # fp32: separate entry points only


GEMM_CASES = [
    # M, N, K
    (128, 32, 64),
    (256, 64, 128),
    (300, 128, 64),
    (1000, 256, 256),
    (517, 728, 728),
    (2048, 21, 2304),
    (4096, 256, 1280),
    (130, 48, 1024),
    (1024, 1024, 728),
    (96, 16, 32),
    (4100, 96, 576),
]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm_bf16_plain(case):
    o = ops()
    M, N, K = case
    a = rnd((M, K), torch.bfloat16, 11)
    b = rnd((N, K), torch.bfloat16, 12, 1.0 / math.sqrt(K))
    ref = a.double() @ b.double().t()
    for cdt in (torch.float32, torch.bfloat16):
        out = torch.full((M, N), float("nan"), dtype=cdt, device=DEV)
        o.gemm_bf16(a.to(DEV), b.to(DEV), M, N, K, out)
        rt, at = (1e-4, 1e-4) if cdt == torch.float32 else (1e-2, 1e-2)
        check(f"gemm {case} {cdt}", out, ref, rt, at)


@pytest.mark.parametrize("case", [(517, 728, 728), (1000, 256, 256), (2048, 21, 288), (300, 40, 64)])
def test_gemm_bf16_epilogue(case):
    o = ops()
    M, N, K = case
    a = rnd((M, K), torch.bfloat16, 21)
    b = rnd((N, K), torch.bfloat16, 22, 1.0 / math.sqrt(K))
    sc = rnd((N,), torch.float32, 23, 0.2) + 1.0
    sh = rnd((N,), torch.float32, 24, 0.5)
    ldc = N + 8
    add = rnd((M, ldc), torch.bfloat16, 25)
    acc = a.double() @ b.double().t()
    ref = torch.relu(acc * sc.double() + sh.double()) + add[:, :N].double()
    out = torch.zeros((M, ldc), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros((2, N), dtype=torch.float32, device=DEV)
    o.gemm_bf16(a.to(DEV), b.to(DEV), M, N, K, out, ldc=ldc, col_scale=sc.to(DEV), col_shift=sh.to(DEV),
                act=o.ACT_RELU, addend=add.to(DEV), ld_addend=ldc, col_stats=stats)
    check("gemm epi", out[:, :N], ref, 1e-2, 2e-2)
    assert float(out[:, N:].abs().max()) == 0.0, "gemm wrote outside its channel slice"
    check("gemm stats sum", stats[0], acc.sum(0), 1e-3, 1e-3 * math.sqrt(M))
    check("gemm stats sumsq", stats[1], (acc * acc).sum(0), 1e-3, 1e-3 * M)


@pytest.mark.parametrize("case", [(517, 728, 728), (1000, 256, 256), (300, 64, 64), (4100, 96, 576), (16384, 1024, 128),
                                  (130, 200, 72)])
def test_gemm_bf16_tma_store_stats(case):
    """bf16 output without addend takes the staged TMA-store epilogue: channel-slice write (ldc > N), fused
    scale/shift/ReLU, and the BatchNormalization statistics of the STORED (bf16-rounded) accumulator."""
    o = ops()
    M, N, K = case
    a = rnd((M, K), torch.bfloat16, 26)
    b = rnd((N, K), torch.bfloat16, 27, 1.0 / math.sqrt(K))
    acc = a.double() @ b.double().t()
    ldc = N + 16
    out = torch.zeros((M, ldc), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros((2, N), dtype=torch.float32, device=DEV)
    o.gemm_bf16(a.to(DEV), b.to(DEV), M, N, K, out, ldc=ldc, col_stats=stats)
    check("gemm tma store", out[:, :N], acc, 1e-2, 1e-2)
    assert float(out[:, N:].abs().max()) == 0.0, "gemm wrote outside its channel slice"
    stored = out[:, :N].double().cpu()
    check("stats sum (stored)", stats[0], stored.sum(0), 1e-4, 1e-4 * math.sqrt(M))
    check("stats sumsq (stored)", stats[1], (stored * stored).sum(0), 1e-4, 1e-4 * M)
    sc = rnd((N,), torch.float32, 28, 0.2) + 1.0
    sh = rnd((N,), torch.float32, 29, 0.5)
    out2 = torch.zeros((M, ldc), dtype=torch.bfloat16, device=DEV)
    o.gemm_bf16(a.to(DEV), b.to(DEV), M, N, K, out2, ldc=ldc, col_scale=sc.to(DEV), col_shift=sh.to(DEV), act=o.ACT_RELU)
    check("gemm tma store epi", out2[:, :N], torch.relu(acc * sc.double() + sh.double()), 1e-2, 2e-2)
    assert float(out2[:, N:].abs().max()) == 0.0


def test_gemm_bf16_strided_a():
    """A is a channel slice of a wider tensor (lda > K), as when a branch reads part of a concat."""
    o = ops()
    M, N, K, lda = 700, 256, 256, 1280
    big = rnd((M, lda), torch.bfloat16, 31)
    b = rnd((N, K), torch.bfloat16, 32, 1.0 / math.sqrt(K))
    off = 512
    ref = big[:, off:off + K].double() @ b.double().t()
    out = torch.empty((M, N), dtype=torch.float32, device=DEV)
    bd = big.to(DEV)
    o.gemm_bf16(bd[:, off:], b.to(DEV), M, N, K, out, lda=lda)
    check("gemm strided A", out, ref, 1e-4, 1e-4)


@pytest.mark.parametrize("case", [(1000, 64, 64), (4096, 128, 256), (5000, 728, 728), (16384, 256, 24),
                                  (3000, 32, 64), (2000, 1024, 256), (900, 304, 24), (8192, 256, 1024), (4100, 304, 256)])
def test_gemm_wgrad_bf16(case):
    o = ops()
    M, K, N = case
    ldy = ((N + 7) // 8) * 8
    x = rnd((M, K), torch.bfloat16, 41)
    dy = rnd((M, ldy), torch.bfloat16, 42)
    ref = x.double().t() @ dy[:, :N].double()
    dw0 = rnd((K, N), torch.float32, 43)
    dw = dw0.to(DEV).clone()
    o.gemm_wgrad_bf16(x.to(DEV), dy.to(DEV), dw, M, K, N, ldy=ldy)
    check(f"wgrad {case}", dw, ref + dw0.double(), 1e-3, 1e-3 * math.sqrt(M))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt(dtype):
    o = ops()
    M, N, K = 333, 77, 130
    a = rnd((M, K), dtype, 51)
    b = rnd((K, N), dtype, 52, 1.0 / math.sqrt(K))
    ref = a.double() @ b.double()
    out = torch.empty((M, N), dtype=torch.float32, device=DEV)
    o.gemm_simt(a.to(DEV), K, 1, b.to(DEV), N, 1, out, N, M, N, K)
    check("simt nn", out, ref, 1e-4, 1e-4)
    # transposed operands: C = A^T-view * B^T-view
    at, bt = a.t().contiguous(), b.t().contiguous()
    out2 = torch.empty((M, N), dtype=torch.float32, device=DEV)
    o.gemm_simt(at.to(DEV), 1, M, bt.to(DEV), 1, K, out2, N, M, N, K)
    check("simt tt", out2, ref, 1e-4, 1e-4)
    out3 = out2.clone()
    o.gemm_simt(a.to(DEV), K, 1, b.to(DEV), N, 1, out3, N, M, N, K, accumulate=True)
    check("simt acc", out3, 2 * ref, 1e-4, 2e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [(2, 13, 11, 3, 2, "valid"), (1, 12, 12, 32, 1, "valid"), (2, 8, 9, 256, 1, "same"),
                                  (1, 9, 9, 24, 2, "same")])
def test_im2col_conv(case, dtype):
    """dense 3x3 conv = im2col + GEMM against tf.nn.conv2d; col2im against autograd."""
    o = ops()
    N, H, W, C, stride, padding = case
    Co = 21
    x = rnd((N, H, W, C), dtype, 61)
    w = rnd((3, 3, C, Co), torch.float32, 62, 0.2)
    ho, wo, pt, pl = o.conv_geometry(H, W, 3, stride, (1, 1), padding)
    xr = x.double().requires_grad_(True)
    y_ref = O.conv2d(xr, w.double(), stride, padding)
    assert tuple(y_ref.shape) == (N, ho, wo, Co)
    ld = ((9 * C + 7) // 8) * 8
    col = o.im2col3x3(x.to(DEV), stride, 1, ho, wo, pt, pl, ld)
    y = torch.empty((N * ho * wo, Co), dtype=torch.float32, device=DEV)
    wk = torch.zeros((ld, Co), dtype=dtype)
    wk[:9 * C] = w.reshape(9 * C, Co).to(dtype)
    o.gemm_simt(col, ld, 1, wk.to(DEV), Co, 1, y, Co, N * ho * wo, Co, ld)
    y_ref2 = O.conv2d(x.double(), wk[:9 * C].double().reshape(3, 3, C, Co), stride, padding)
    rt, at = tol(dtype)
    check("im2col conv", y.view(N, ho, wo, Co), y_ref2, rt, at)
    if C % 8 == 0:
        gcol = rnd((N * ho * wo, ld), dtype, 63)
        gcol[:, 9 * C:] = 0
        # reference: scatter-add of the column gradient == autograd of im2col
        xx = x.double().requires_grad_(True)
        cols = []
        xp = torch.nn.functional.pad(xx.permute(0, 3, 1, 2), (pl, 2 + stride, pt, 2 + stride))
        for i in range(3):
            for j in range(3):
                cols.append(xp[:, :, i:i + stride * ho:stride, j:j + stride * wo:stride].permute(0, 2, 3, 1))
        colr = torch.cat(cols, dim=-1).reshape(N * ho * wo, 9 * C)
        check("im2col", col[:, :9 * C], colr, 0, 0)
        colr.backward(gcol[:, :9 * C].double())
        dx = o.col2im3x3(gcol.to(DEV), (N, H, W, C), stride, 1, ho, wo, pt, pl, ld)
        check("col2im", dx, xx.grad, rt, at * 3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_subsample(dtype):
    o = ops()
    x = rnd((2, 9, 10, 64), dtype, 71)
    y = o.subsample_fwd(x.to(DEV), 2)
    check("subsample", y, x[:, ::2, ::2], 0, 0)
    gy = rnd(tuple(y.shape), dtype, 72)
    add = rnd((2, 9, 10, 64), dtype, 73)
    dx = o.subsample_bwd(gy.to(DEV), (2, 9, 10, 64), 2, addend=add.to(DEV))
    ref = add.double().clone()
    ref[:, ::2, ::2] += gy.double()
    rt, at = tol(dtype)
    check("subsample bwd", dx, ref, rt, at)


def test_weight_prep():
    o = ops()
    K, N = 100, 21
    w = rnd((K, N), torch.float32, 81)
    ldt, ldn = 104, 24
    wt = torch.zeros((N, ldt), dtype=torch.bfloat16, device=DEV)
    wn = torch.zeros((K, ldn), dtype=torch.bfloat16, device=DEV)
    o.weight_prep(w.to(DEV), K, N, wt, ldt, wn, ldn)
    check("wt", wt[:, :K], w.t().to(torch.bfloat16), 0, 0)
    check("wn", wn[:, :N], w.to(torch.bfloat16), 0, 0)
    assert float(wt[:, K:].abs().max()) == 0 and float(wn[:, N:].abs().max()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("MC", [(1000, 64), (4097, 728), (300, 24), (20000, 256)])
def test_batchnorm_train(MC, dtype):
    o = ops()
    M, C = MC
    y = (rnd((M, C), torch.float64, 91) * 1.5 + 0.7).to(dtype)
    gamma = rnd((C,), torch.float32, 92, 0.2) + 1.0
    beta = rnd((C,), torch.float32, 93, 0.3)
    mm = rnd((C,), torch.float32, 94)
    mv = rnd((C,), torch.float32, 95).abs() + 0.5
    eps, mom = 1e-3, 0.99
    yr = y.double().view(1, 1, M, C).requires_grad_(True)
    z_ref, mm_ref, mv_ref = O.batch_norm(yr, gamma.double(), beta.double(), mm.double(), mv.double(), eps, True, mom)
    res = rnd((M, C), dtype, 96)
    out_ref = torch.relu(z_ref.view(M, C)) + res.double()
    gz = rnd((M, C), dtype, 97)
    out_ref.backward(gz.double())

    yd = y.to(DEV)
    sums = torch.zeros((2, C), dtype=torch.float32, device=DEV)
    o.bn_stats(yd, M, C, sums)
    scale, shift, mean, invstd = (torch.empty(C, dtype=torch.float32, device=DEV) for _ in range(4))
    mmd, mvd = mm.to(DEV), mv.to(DEV)
    o.bn_finalize(sums, gamma.to(DEV), beta.to(DEV), mmd, mvd, C, M, eps, mom, scale, shift, mean, invstd)
    check("bn mean", mean, y.double().mean(0), 1e-4, 1e-4)
    check("bn moving mean", mmd, mm_ref, 1e-4, 1e-4)
    check("bn moving var", mvd, mv_ref, 1e-3, 1e-4)
    out = torch.empty((M, C), dtype=dtype, device=DEV)
    o.affine_act(yd, M, C, out, scale, shift, o.ACT_RELU, addend=res.to(DEV))
    rt, at = tol(dtype)
    check("bn apply", out, out_ref, rt, at)

    red = torch.zeros((2, C), dtype=torch.float32, device=DEV)
    gzd = gz.to(DEV)
    o.bn_bwd_reduce(gzd, yd, scale, shift, mean, invstd, o.ACT_RELU, M, C, red)
    dy = torch.empty((M, C), dtype=dtype, device=DEV)
    o.bn_bwd_apply(gzd, yd, scale, shift, mean, invstd, o.ACT_RELU, red, M, C, dy)
    # gradients w.r.t. beta / gamma recomputed from the same bf16-rounded forward
    g_eff = gz.double() * (z_ref.view(M, C) > 0).double()
    xhat = (y.double() - y.double().mean(0)) / torch.sqrt(y.double().var(0, unbiased=False) + eps)
    # a pre-activation within fp32 rounding of zero may land on either side of the ReLU (the batch mean depends on the
    # summation order): such elements may flip, each moving a per-channel sum by |g| — budget exactly that
    amb = (z_ref.view(M, C).abs() < 1e-5).detach()
    slack = (gz.double().abs() * amb).sum(0)
    for name, got, want, extra in (("bn dbeta", red[0], g_eff.sum(0), slack),
                                   ("bn dgamma", red[1], (g_eff * xhat).sum(0), slack * xhat.abs().max())):
        err = (got.double().cpu() - want).abs()
        bound = 2e-3 * want.abs() + 2e-3 * math.sqrt(M) + 1.01 * extra
        assert bool((err <= bound).all()), (name, float(err.max()), float(bound.min()))
    keep = (~amb).double()
    check("bn dx", dy.double().cpu() * keep, yr.grad.view(M, C) * keep, rt, at + float(slack.max()) * 8.0 / M)


def test_bn_fold_inference():
    o = ops()
    C = 48
    gamma, beta = rnd((C,), torch.float32, 1) + 1, rnd((C,), torch.float32, 2)
    mm, mv = rnd((C,), torch.float32, 3), rnd((C,), torch.float32, 4).abs() + 0.1
    scale, shift = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    o.bn_fold(gamma.to(DEV), beta.to(DEV), mm.to(DEV), mv.to(DEV), C, 1e-3, scale, shift)
    x = rnd((1, 1, 5, C), torch.float32, 5)
    ref, _, _ = O.batch_norm(x.double(), gamma.double(), beta.double(), mm.double(), mv.double(), 1e-3, False)
    check("bn fold", x.to(DEV) * scale + shift, ref, 1e-5, 1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("HW", [(254, 254), (127, 127), (64, 64), (9, 12)])
def test_maxpool_add(HW, dtype):
    o = ops()
    H, W = HW
    N, C = 1, 16
    x = rnd((N, H, W, C), dtype, 101)
    xr = x.double().requires_grad_(True)
    y_ref = O.max_pool_3x3_s2_same(xr)
    res = rnd(tuple(y_ref.shape), dtype, 102)
    gy = rnd(tuple(y_ref.shape), dtype, 103)
    y_ref.backward(gy.double())
    am = torch.empty(tuple(y_ref.shape), dtype=torch.uint8, device=DEV)
    y = o.maxpool3x3s2_fwd(x.to(DEV), argmax=am, addend=res.to(DEV))
    rt, at = tol(dtype)
    check("maxpool", y, y_ref + res.double(), rt, at)
    add = rnd((N, H, W, C), dtype, 104)
    dx = o.maxpool3x3s2_bwd(gy.to(DEV), am, (N, H, W, C), addend=add.to(DEV))
    check("maxpool bwd", dx, xr.grad + add.double(), rt, at * 2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_avgpool(dtype):
    o = ops()
    x = rnd((2, 8, 12, 32), dtype, 111)
    xr = x.double().requires_grad_(True)
    y_ref = O.avg_pool_valid(xr, 4)
    gy = rnd(tuple(y_ref.shape), dtype, 112)
    y_ref.backward(gy.double())
    y = o.avgpool_fwd(x.to(DEV), 4)
    rt, at = tol(dtype)
    check("avgpool", y, y_ref, rt, at)
    dx = o.avgpool_bwd(gy.to(DEV), (2, 8, 12, 32), 4)
    check("avgpool bwd", dx, xr.grad, rt, at)


@pytest.mark.parametrize("case", [(2, 5, 7, 21, 16, 16, torch.float32), (1, 9, 9, 48, 8, 8, torch.bfloat16),
                                  (1, 6, 5, 256, 4, 4, torch.bfloat16), (2, 4, 4, 21, 2, 2, torch.float32),
                                  (1, 3, 4, 8, 1, 1, torch.float32), (1, 33, 33, 21, 16, 16, torch.float32)])
def test_bilinear(case):
    o = ops()
    N, H, W, C, fh, fw, dtype = case
    x = rnd((N, H, W, C), dtype, 121)
    xr = x.double().requires_grad_(True)
    y_ref = O.resize_bilinear(xr, fh, fw)
    t_ref = torch.nn.functional.interpolate(x.double().permute(0, 3, 1, 2), scale_factor=(fh, fw), mode="bilinear",
                                            align_corners=False).permute(0, 2, 3, 1)
    check("oracle resize vs torch", y_ref, t_ref, 1e-12, 1e-12)
    gy = rnd(tuple(y_ref.shape), dtype, 122)
    y_ref.backward(gy.double())
    y = o.bilinear_fwd(x.to(DEV), fh, fw)
    rt, at = tol(dtype)
    check("bilinear", y, y_ref, rt, at)
    dx = o.bilinear_bwd(gy.to(DEV), (N, H, W, C), fh, fw)
    check("bilinear bwd", dx, xr.grad, rt, at * fh)


def _labels(shape, C, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, C, shape, generator=g, dtype=torch.int32)


PW = [0.29754999, 0.99106889, 0.99236374, 0.99122957, 0.99350396, 0.99455487, 0.98728424, 0.98090446, 0.96883489,
      0.98753125, 0.99376389, 0.98942612, 0.97222875, 0.99080578, 0.98845309, 0.92606652, 0.99393374, 0.99374322,
      0.98782171, 0.98659656, 0.99233476]
NW = [1.0 - v for v in PW]


def test_softmax_cbloss():
    o = ops()
    P, C = 5000, 21
    z = rnd((P, C), torch.float32, 131, 3.0)
    lab = _labels((P,), C, 132)
    zr = z.double().requires_grad_(True)
    p_ref = O.softmax(zr)
    loss_ref = O.class_balanced_loss(O.one_hot(lab, C), p_ref, PW, NW)
    loss_ref.backward()
    pw, nw = torch.tensor(PW, device=DEV), torch.tensor(NW, device=DEV)
    ls = torch.zeros(1, device=DEV)
    probs = torch.empty((P, C), device=DEV)
    o.softmax_cbloss_fwd(z.to(DEV), lab.to(DEV), pw, nw, 1e-7, P, C, ls, probs)
    check("probs", probs, p_ref, 1e-5, 1e-6)
    check("loss", ls / P, loss_ref.reshape(1), 1e-4, 1e-6)
    dz = torch.empty((P, C), device=DEV)
    o.softmax_cbloss_bwd(z.to(DEV), lab.to(DEV), pw, nw, 1e-7, P, C, 1.0 / P, dz)
    check("dz", dz, zr.grad, 1e-3, 1e-8)
    # argmax / label map
    labels = torch.empty(P, dtype=torch.int32, device=DEV)
    o.softmax_argmax(z.to(DEV), P, C, labels=labels)
    assert torch.equal(labels.cpu().long(), O.argmax_labels(z))
    # dense (Keras-signature) loss
    yt = O.one_hot(lab, C, torch.float32)
    ls2 = torch.zeros(1, device=DEV)
    o.cbloss_dense_fwd(yt.to(DEV), probs, pw, nw, 1e-7, P, C, ls2)
    check("dense loss", ls2 / P, loss_ref.reshape(1), 1e-4, 1e-6)
    pr = p_ref.detach().clone().requires_grad_(True)
    O.class_balanced_loss(O.one_hot(lab, C), pr, PW, NW).backward()
    dyp = torch.empty((P, C), device=DEV)
    o.cbloss_dense_bwd(yt.to(DEV), probs, pw, nw, 1e-7, P, C, 1.0 / P, dyp)
    check("dense dloss", dyp, pr.grad, 2e-3, 1e-9)
    dz2 = torch.empty((P, C), device=DEV)
    o.softmax_bwd(probs, dyp, P, C, dz2)
    check("softmax bwd", dz2, zr.grad, 2e-3, 1e-8)
    cm = torch.zeros((C, C), dtype=torch.float64, device=DEV)
    o.confusion_matrix(lab.to(DEV), labels, P, C, cm)
    check("confusion", cm, O.confusion_matrix(lab, labels.cpu(), C), 0, 0)


def test_fused_upsample_argmax_equals_materialised_path():
    """dlv3p_upsample_argmax (inference tail: no high-resolution logits) gives the label map of bilinear_fwd +
    softmax_argmax bit for bit — including exact ties (first maximum wins) — for int32 and uint8 labels, anisotropic
    factors, and matches the oracle's resize + argmax."""
    o = ops()
    for (N, H, W, C, fh, fw) in ((2, 9, 11, 21, 16, 16), (1, 33, 33, 21, 16, 16), (1, 5, 7, 19, 8, 4), (2, 4, 4, 3, 1, 1)):
        z = rnd((N, H, W, C), torch.float32, 3 + H, 2.0)
        z[..., 1] = z[..., 0]                              # ties everywhere between classes 0 and 1
        zd = z.to(DEV)
        zh = o.bilinear_fwd(zd, fh, fw)
        P = zh.numel() // C
        want = torch.empty(P, dtype=torch.int32, device=DEV)
        o.softmax_argmax(zh.view(P, C), P, C, labels=want)
        for dt in (torch.int32, torch.uint8):
            lab = torch.full((N, H * fh, W * fw), 77, dtype=dt, device=DEV)
            o.upsample_argmax(zd, fh, fw, lab)
            assert torch.equal(lab.view(-1).to(torch.int32), want), (N, H, W, C, fh, fw, dt)
        ref = O.argmax_labels(O.resize_bilinear(z.double(), fh, fw))
        assert (want.view(N, H * fh, W * fw).cpu().long() == ref).float().mean() > 0.999


@pytest.mark.parametrize("case", [(2, 6, 7, 21, 16), (1, 5, 5, 21, 2), (1, 4, 6, 19, 8), (2, 3, 3, 21, 4),
                                  (1, 3, 4, 21, 6), (1, 2, 3, 5, 16), (1, 3, 3, 16, 8), (1, 2, 2, 12, 4)])
def test_fused_upsample_softmax_cbloss(case):
    o = ops()
    N, H, W, C, f = case
    zl = rnd((N, H, W, C), torch.float32, 141, 2.0)
    lab = _labels((N, H * f, W * f), C, 142)
    zr = zl.double().requires_grad_(True)
    p_ref = O.softmax(O.resize_bilinear(zr, f, f))
    loss_ref = O.class_balanced_loss(O.one_hot(lab, C), p_ref, PW[:C], NW[:C])
    loss_ref.backward()
    pw, nw = torch.tensor(PW[:C], device=DEV), torch.tensor(NW[:C], device=DEV)
    P = N * H * f * W * f
    ls = torch.zeros(1, device=DEV)
    o.upsample_softmax_cbloss_fwd(zl.to(DEV), lab.to(DEV), pw, nw, 1e-7, N, H, W, C, f, ls)
    check("fused loss", ls / P, loss_ref.reshape(1), 1e-4, 1e-6)
    dzl = torch.zeros((N, H, W, C), device=DEV)
    o.upsample_softmax_cbloss_bwd(zl.to(DEV), lab.to(DEV), pw, nw, 1e-7, N, H, W, C, f, 1.0 / P, dzl)
    check("fused dzl", dzl, zr.grad, 2e-3, 1e-7)
    # the one-pass training entry point gives the same loss and gradient
    ls2, dzl2 = torch.zeros(1, device=DEV), torch.zeros((N, H, W, C), device=DEV)
    o.upsample_softmax_cbloss_fwd_bwd(zl.to(DEV), lab.to(DEV), pw, nw, 1e-7, N, H, W, C, f, 1.0 / P, ls2, dzl2)
    check("fused loss (fwd_bwd)", ls2 / P, loss_ref.reshape(1), 1e-4, 1e-6)
    check("fused dzl (fwd_bwd)", dzl2, zr.grad, 2e-3, 1e-7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_eltwise_misc(dtype):
    o = ops()
    a, b = rnd((3, 5, 7, 24), dtype, 151), rnd((3, 5, 7, 24), dtype, 152)
    out = torch.empty_like(a, device=DEV)
    rt, at = tol(dtype)
    check("add", o.add(a.to(DEV), b.to(DEV), out), a.double() + b.double(), rt, at)
    dx = torch.empty_like(a, device=DEV)
    o.act_bwd(a.to(DEV), b.to(DEV), o.ACT_RELU6, dx, addend=a.to(DEV))
    m = ((b.double() > 0) & (b.double() < 6)).double()
    check("act bwd", dx, a.double() * m + a.double(), rt, at)
    # concat slice copy
    big = torch.zeros((3 * 5 * 7, 64), dtype=dtype, device=DEV)
    o.copy2d(a.to(DEV), 24, big, 64, 3 * 5 * 7, 24, y_off=16)
    check("copy2d", big[:, 16:40], a.view(-1, 24), 0, 0)
    assert float(big[:, :16].abs().max()) == 0 and float(big[:, 40:].abs().max()) == 0
    # dropout: same mask fwd/bwd, keep-rate statistics
    x = torch.ones((1 << 16,), dtype=dtype, device=DEV)
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    o.dropout(x, 0.5, 1234, y1)
    o.dropout(x, 0.5, 1234, y2)
    assert torch.equal(y1, y2)
    keep = float((y1 > 0).float().mean())
    assert abs(keep - 0.5) < 0.02 and float(y1.max()) == 2.0
    c = torch.empty((3, 5, 7, 24), dtype=torch.float32 if dtype == torch.bfloat16 else torch.bfloat16, device=DEV)
    check("cast", o.cast(a.to(DEV), c), a.double(), 1e-2, 1e-2)


def test_adam_sumsq():
    o = ops()
    n = 10007
    w, g = rnd((n,), torch.float32, 161), rnd((n,), torch.float32, 162)
    m, v = torch.zeros(n), torch.zeros(n)
    lr, b1, b2, eps, l2 = 1e-3, 0.5, 0.99, 1e-7, 4e-5
    wd, md, vd = w.to(DEV), m.to(DEV), v.to(DEV)
    wr, mr, vr = w.double().clone(), m.double(), v.double()
    for t in range(1, 4):
        lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        o.adam(wd, g.to(DEV), md, vd, n, lr_t, b1, b2, eps, 1.0, l2)
        ge = g.double() + 2 * l2 * wr
        mr = b1 * mr + (1 - b1) * ge
        vr = b2 * vr + (1 - b2) * ge * ge
        wr = wr - lr_t * mr / (vr.sqrt() + eps)
    check("adam", wd, wr, 1e-4, 1e-6)
    s = torch.zeros(1, device=DEV)
    o.sumsq(w.to(DEV), n, s)
    check("sumsq", s, (w.double() ** 2).sum().reshape(1), 1e-4, 0)


def test_errors_are_loud():
    o = ops()
    x = torch.zeros((1, 4, 4, 12), device=DEV)
    w = torch.zeros((3, 3, 12), device=DEV)
    with pytest.raises(ValueError):
        o.dwconv3x3_fwd(x, w)          # C not a multiple of 8
    with pytest.raises(ValueError):
        o.dwconv3x3_fwd(torch.zeros((1, 4, 4, 8)), torch.zeros((3, 3, 8)))   # CPU tensor: no CPU path


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("MC", [(1000, 64), (4097, 728), (20000, 256)])
def test_bn_train_apply_matches_two_step(MC, dtype):
    """The one-launch training BatchNormalization forward equals bn_finalize (x updates) + affine_act."""
    o = ops()
    M, C = MC
    y = (rnd((M, C), torch.float64, 191) * 1.5 + 0.7).to(dtype).to(DEV)
    res = rnd((M, C), dtype, 192).to(DEV)
    gamma, beta = (rnd((C,), torch.float32, 193, 0.2) + 1.0).to(DEV), rnd((C,), torch.float32, 194, 0.3).to(DEV)
    mm0, mv0 = rnd((C,), torch.float32, 195), rnd((C,), torch.float32, 196).abs() + 0.5
    sums = torch.zeros((2, C), dtype=torch.float32, device=DEV)
    o.bn_stats(y, M, C, sums)
    for updates, act, use_add in ((1, o.ACT_RELU, True), (2, o.ACT_NONE, False), (0, o.ACT_RELU6, False)):
        ref = [torch.empty(C, dtype=torch.float32, device=DEV) for _ in range(4)]
        mm_r, mv_r = mm0.to(DEV), mv0.to(DEV)
        for u in range(max(updates, 1)):
            o.bn_finalize(sums, gamma, beta, mm_r, mv_r, C, M, 1e-3, 0.9, *ref, update_moving=u < updates)
        out_r = torch.empty((M, C), dtype=dtype, device=DEV)
        o.affine_act(y, M, C, out_r, ref[0], ref[1], act, addend=res if use_add else None)
        got = [torch.empty(C, dtype=torch.float32, device=DEV) for _ in range(4)]
        mm_g, mv_g = mm0.to(DEV), mv0.to(DEV)
        out_g = torch.empty((M, C), dtype=dtype, device=DEV)
        o.bn_train_apply(y, M, C, sums, gamma, beta, mm_g, mv_g, M, 1e-3, 0.9, updates, act, out_g, *got,
                         addend=res if use_add else None)
        for a, b, nm in zip(got + [mm_g, mv_g], ref + [mm_r, mv_r], ("scale", "shift", "mean", "invstd", "mm", "mv")):
            check(f"bn_train_apply {nm}", a, b, 2e-6, 2e-6)
        # the fused kernel finishes the statistics with an fp32 rsqrt (bn_finalize: fp64 1/sqrt): scale/shift agree to
        # ~3e-7 relative, the outputs to that times |y| (bf16: an occasional one-ulp flip of the stored value)
        rt = 1e-5 if dtype == torch.float32 else 1e-2
        check("bn_train_apply out", out_g, out_r, rt, rt)


def test_weight_prep_batch_matches_single():
    o = ops()
    entries, singles = [], []
    for i, (K, N) in enumerate([(728, 728), (27, 32), (2304, 21), (64, 128)]):
        w = rnd((K, N), torch.float32, 200 + i).to(DEV)
        Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
        wt = torch.zeros((N, Kp), dtype=torch.bfloat16, device=DEV)
        wn = torch.zeros((K, Np), dtype=torch.bfloat16, device=DEV) if i % 2 == 0 else None
        wt2, wn2 = torch.zeros_like(wt), (torch.zeros_like(wn) if wn is not None else None)
        o.weight_prep(w, K, N, wt2, Kp, wn2, Np)
        entries.append((w, K, N, wt, Kp, wn, Np))
        singles.append((wt2, wn2))
    table = o.weight_prep_table(entries, DEV)
    o.weight_prep_batch(table, len(entries))
    for (w, K, N, wt, Kp, wn, Np), (wt2, wn2) in zip(entries, singles):
        assert torch.equal(wt, wt2)
        if wn is not None:
            assert torch.equal(wn, wn2)

"""ORACLE (test infrastructure only) — straight-line CPU restatement of the reference's DeepLabV3+ graph.

Follows bodhi/deeplabv3plus_keras/semantic_segmentation.py: base-model taps ss.py:494-525, _make_encoder
ss.py:790-876, _make_decoder ss.py:878-913, _refine_boundary ss.py:915-954, loss ss.py:438-447 (+ Keras L2
regularisers of the Conv2D layers created with kernel_regularizer, ss.py:818,838,847,869,897,935), and the
keras.applications Xception / MobileNetV2 topologies (TF 2.4; source not vendored in the reference — restated from
the published architectures, pinned by their parameter totals).  PARITY UNPINNED against TensorFlow itself: see
oracle/tf_ops.py.  Deliberately written as plain functions over a {"layer/weight": tensor} dict — it shares no
code with the product's layer graph / engine, so a topology bug in one is caught by the other.

Weights are keyed with the names tf.keras would give the layers, including the automatic ones
(conv2d, conv2d_1, ..., separable_conv2d, batch_normalization_k) in creation order.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import tf_ops as T


class Names:
    """tf.keras automatic layer naming: prefix, prefix_1, prefix_2, ..."""

    def __init__(self):
        self.n: Dict[str, int] = {}

    def __call__(self, prefix: str) -> str:
        k = self.n.get(prefix, 0)
        self.n[prefix] = k + 1
        return prefix if k == 0 else f"{prefix}_{k}"


def _bf16(t):
    """Round to bfloat16 and back (differentiable: the cast's gradient is the cast)."""
    return t.to(torch.bfloat16).to(t.dtype)


class Ctx:
    """`emulate_bf16`: round exactly where the product STORES bfloat16 — GEMM operands (1x1 / im2col'd kernels and
    their inputs), every tensor a kernel writes to HBM (depthwise output, GEMM output, fused BN+activation(+add)
    output, pool(+add), resize, concat) — while arithmetic stays in the oracle's precision, as the kernels' fp32
    accumulators do.  Logits are not rounded.  BatchNormalization batch statistics are those of the STORED conv output
    (the GEMM epilogue reduces its bf16 staging tile; only GEMMs with <= 32 output channels still reduce the fp32
    accumulators — _cbn models that rule).
    With it off (default) the oracle is the plain fp64/fp32 restatement."""

    def __init__(self, weights: Dict[str, torch.Tensor], training: bool, bn_momentum_new: Dict[str, torch.Tensor],
                 emulate_bf16: bool = False, momentum_override: Optional[float] = None, probe: Optional["Probe"] = None):
        self.w, self.training, self.new_stats = weights, training, bn_momentum_new
        self.names = Names()
        self.l2_terms: List[torch.Tensor] = []
        self.q = _bf16 if emulate_bf16 else (lambda t: t)
        self.momentum_override = momentum_override
        self.probe = probe

    # ---- named storage points / decision sites (see Probe) ------------------------------------------------
    def store(self, name: str, t, rounded: bool = True):
        """A tensor the product writes to HBM under this name.  Rounded like the product's storage; with a Probe the
        value is recorded and, under teacher forcing, replaced by the product's own stored tensor."""
        if rounded:
            t = self.q(t)
        return t if self.probe is None else self.probe.on_store(name, t)

    def act(self, site: str, z, kind: str):
        """ReLU / ReLU6 at a named decision site (the Keras layer that produced z)."""
        if self.probe is not None:
            return self.probe.on_act(site, z, kind)
        return T.relu(z) if kind == "relu" else T.relu6(z)

    def max_pool(self, site: str, x):
        if self.probe is not None:
            return self.probe.on_pool(site, x)
        return T.max_pool_3x3_s2_same(x)


class Probe:
    """Test instrumentation of the oracle graph (never used by the plain oracle runs).

    * `values` / `pre` / `pool_arg`: every named storage point, every activation's pre-activation tensor and every
      max-pool's winning tap (row-major 0..8 inside the 3x3 window, first maximum wins), as the oracle computed them.
    * teacher forcing (`teacher`, `teacher_grad`): at a storage point the product also has, the oracle's tensor is
      compared with the product's (`fwd[name] = (oracle, product)`) and then REPLACED by it — value of the product,
      gradient path of the oracle — so every operation is checked on identical inputs, with no error amplification
      through the depth of the network.  In backward the gradient arriving at that point is compared with the
      product's gradient buffer (`bwd[name]`) and replaced by it the same way.
    * forced decisions (`masks`, `pool_taps`): the oracle runs freely but takes every ReLU/ReLU6 mask and every
      max-pool winner from the product, so that the two compute the gradient of the SAME piecewise-smooth function;
      what is left is arithmetic error.  The natural decisions stay recorded so that the test can check that the two
      only disagree inside the forward-error band around a tie."""

    def __init__(self, teacher=None, teacher_grad=None, masks=None, pool_taps=None, keep_values=True):
        self.teacher, self.teacher_grad = teacher or {}, teacher_grad or {}
        self.masks, self.pool_taps = masks or {}, pool_taps or {}
        self.keep_values = keep_values
        self.values: Dict[str, torch.Tensor] = {}
        self.pre: Dict[str, torch.Tensor] = {}
        self.pool_arg: Dict[str, torch.Tensor] = {}
        self.pool_margin: Dict[str, torch.Tensor] = {}
        self.fwd: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.bwd: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.order: List[str] = []

    def on_store(self, name, t):
        self.order.append(name)
        if self.keep_values:
            self.values[name] = t.detach()
        tv = self.teacher.get(name)
        if tv is None:
            return t
        tv = tv.to(t.dtype)
        if tuple(tv.shape) != tuple(t.shape):
            raise ValueError(f"teacher tensor {name}: shape {tuple(tv.shape)} vs oracle {tuple(t.shape)}")
        self.fwd[name] = (t.detach(), tv)
        out = t + (tv - t).detach()
        gv = self.teacher_grad.get(name)
        if gv is not None and out.requires_grad:
            def hook(g, name=name, gv=gv):
                self.bwd[name] = (g.detach().clone(), gv)
                return gv.to(g.dtype)
            out.register_hook(hook)
        return out

    def on_act(self, site, z, kind):
        self.pre[site] = z.detach()
        m = self.masks.get(site)
        if m is None:
            return T.relu(z) if kind == "relu" else T.relu6(z)
        if tuple(m.shape) != tuple(z.shape):
            raise ValueError(f"mask {site}: shape {tuple(m.shape)} vs {tuple(z.shape)}")
        if kind == "relu":
            return z * (m == 1).to(z.dtype)
        return torch.where(m == 1, z, torch.where(m == 2, torch.full_like(z, 6.0), torch.zeros_like(z)))

    def on_pool(self, site, x):
        """3x3 / stride 2 / SAME max-pool with recorded (and optionally forced) winners."""
        N, H, W, C = x.shape
        Ho, pt, pb = T.same_pad(H, 3, 2)
        Wo, pl, pr = T.same_pad(W, 3, 2)
        xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float("-inf"))
        win = xp.unfold(2, 3, 2).unfold(3, 3, 2).reshape(N, C, Ho, Wo, 9)        # row-major taps
        best = win.detach().max(dim=-1)
        # first maximum wins (TF / the kernels use a strict >): argmax of the first occurrence
        is_max = win.detach() == best.values.unsqueeze(-1)
        first = torch.argmax(is_max.to(torch.uint8), dim=-1)
        self.pool_arg[site] = first.permute(0, 2, 3, 1)
        top2 = torch.topk(win.detach(), 2, dim=-1).values
        self.pool_margin[site] = (top2[..., 0] - top2[..., 1]).permute(0, 2, 3, 1)
        forced = self.pool_taps.get(site)
        idx = first if forced is None else forced.permute(0, 3, 1, 2).long()
        return torch.gather(win, -1, idx.unsqueeze(-1)).squeeze(-1).permute(0, 2, 3, 1)


def _bn(ctx: Ctx, x, name: str, momentum: float, eps: float = 1e-3, scale: bool = True, stats_from=None):
    """x: the tensor as stored (possibly bf16-rounded); stats_from: the unrounded conv accumulator the product takes
    its batch statistics from (defaults to x)."""
    w = ctx.w
    if ctx.momentum_override is not None:
        momentum = ctx.momentum_override
    gamma = w[f"{name}/gamma"] if scale else None
    beta, mm0, mv0 = w[f"{name}/beta"], w[f"{name}/moving_mean"], w[f"{name}/moving_variance"]
    src = x if stats_from is None else stats_from
    if ctx.training:
        mean = src.mean(dim=(0, 1, 2))
        var = src.var(dim=(0, 1, 2), unbiased=False)
        y = (x - mean) * torch.rsqrt(var + eps)
        y = y * gamma if gamma is not None else y
        y = y + beta
        # moving statistics; a layer applied twice (shared base under boundary refinement) updates them twice
        pm = ctx.new_stats.get(f"{name}/moving_mean", mm0)
        pv = ctx.new_stats.get(f"{name}/moving_variance", mv0)
        _, mm, mv = T.batch_norm(src.detach(), None, None, pm, pv, eps, True, momentum)
        ctx.new_stats[f"{name}/moving_mean"] = mm
        ctx.new_stats[f"{name}/moving_variance"] = mv
        return y
    y, _, _ = T.batch_norm(x, gamma, beta, mm0, mv0, eps, False, momentum)
    return y


def _conv(ctx: Ctx, x, name, stride=1, padding="same", l2: float = 0.0):
    """Returns the UNROUNDED accumulator; callers store it through ctx.q."""
    k = ctx.w[f"{name}/kernel"]
    if l2:
        ctx.l2_terms.append(l2 * (k * k).sum())
    return T.conv2d(x, ctx.q(k), stride, padding)


def _sep(ctx: Ctx, x, name, dilation=(1, 1)):
    d = ctx.store(f"{name}/dw", T.depthwise_conv2d(x, ctx.w[f"{name}/depthwise_kernel"], 1, "same", dilation))
    return T.conv2d(d, ctx.q(ctx.w[f"{name}/pointwise_kernel"]), 1, "same")


def _cbn(ctx: Ctx, acc, bn_name, momentum, act=None, add=None, scale=True, stats_rounded=None, virtual=False):
    """conv accumulator -> stored (rounded) -> BN(+activation)(+residual add) -> stored (rounded): one fused
    macro-op of the product (engine._emit_conv).  `virtual`: in training the product never stores this BN+ReLU output —
    its only reader, a dense-tap stride-1 depthwise stage, applies the map on load in fp32 (engine._BnActValue) — so
    there is no rounding point; in inference the BN is folded into the GEMM epilogue and the output is stored."""
    y = ctx.store(f"{bn_name}/y", acc)
    if stats_rounded is None:
        # the product's rule (gemm_tcgen05.cu): GEMMs with more than 32 output channels take the staged TMA-store
        # epilogue, whose statistics are those of the stored tile; narrower ones reduce the fp32 accumulators
        stats_rounded = acc.shape[-1] > 32
    if stats_rounded:
        src = y
    elif ctx.probe is not None:
        src = y + (acc - y).detach()    # values of the accumulators, gradient through the (hooked) storage point
    else:
        src = acc
    z = _bn(ctx, y, bn_name, momentum, scale=scale, stats_from=src)
    if act is not None:
        z = ctx.act(bn_name, z, "relu6" if act is T.relu6 else "relu")
    if add is not None:
        z = z + add
    # (a virtual tensor is not a rounding point; it is still a named point: the fp32 schedules materialise some of them)
    return ctx.store(f"{bn_name}/out", z, rounded=not (virtual and ctx.training))


# ---- keras.applications.Xception, truncated where the reference taps it (ss.py:517-520) --------------------
def xception_base(ctx: Ctx, img, output_stride: int):
    nm, M = ctx.names, 0.99
    x = _cbn(ctx, _conv(ctx, img, "block1_conv1", 2, "valid"), "block1_conv1_bn", M, T.relu)
    x = _cbn(ctx, _conv(ctx, x, "block1_conv2", 1, "valid"), "block1_conv2_bn", M, T.relu)
    for blk, first_relu in ((2, False), (3, True), (4, True)):
        cname, bname = nm("conv2d"), nm("batch_normalization")
        tap_here = (blk == 4 and output_stride == 8)
        # the strided 1x1 shortcut; pruned from the graph when the network is tapped before the pool (OS8)
        res = None if tap_here else _cbn(ctx, _conv(ctx, x, cname, 2, "same"), bname, M)
        if first_relu:
            x = ctx.act(f"block{blk - 1}_pool", x, "relu")     # block{blk}_sepconv1_act on the previous pool+add output
        x = _cbn(ctx, _sep(ctx, x, f"block{blk}_sepconv1"), f"block{blk}_sepconv1_bn", M, T.relu, virtual=True)
        # (feeds the max-pool directly: in training the product pools scale*y+shift on the fly, no rounding point)
        x = _cbn(ctx, _sep(ctx, x, f"block{blk}_sepconv2"), f"block{blk}_sepconv2_bn", M, virtual=not tap_here)
        if tap_here:
            nm("conv2d"); nm("batch_normalization")       # block13's shortcut layers exist in Keras, pruned here
            return x
        x = ctx.store(f"block{blk}_pool/out", ctx.max_pool(f"block{blk}_pool", x) + res)
    site = "block4_pool"              # name of the layer whose output x is (decision site of the next ReLU)
    for blk in range(5, 13):
        res = x
        for j in (1, 2, 3):
            # sepconv1/2: BN -> the next iteration's ReLU -> dense-tap depthwise stage: virtual in training
            bn = f"block{blk}_sepconv{j}_bn"
            x = _cbn(ctx, _sep(ctx, ctx.act(site, x, "relu"), f"block{blk}_sepconv{j}"), bn, M,
                     add=res if j == 3 else None, virtual=(j < 3))
            site = bn
    nm("conv2d"); nm("batch_normalization")               # block13 shortcut: created by Keras, not on the tapped path
    x = _cbn(ctx, _sep(ctx, ctx.act(site, x, "relu"), "block13_sepconv1"), "block13_sepconv1_bn", M, virtual=True)
    x = _cbn(ctx, _sep(ctx, ctx.act("block13_sepconv1_bn", x, "relu"), "block13_sepconv2"), "block13_sepconv2_bn", M)
    return x


_MNV2 = ((16, 1, 1, 0), (24, 2, 6, 1), (24, 1, 6, 2), (32, 2, 6, 3), (32, 1, 6, 4), (32, 1, 6, 5), (64, 2, 6, 6),
         (64, 1, 6, 7), (64, 1, 6, 8), (64, 1, 6, 9), (96, 1, 6, 10), (96, 1, 6, 11), (96, 1, 6, 12))


def mobilenetv2_base(ctx: Ctx, img, output_stride: int):
    M = 0.999
    x = _cbn(ctx, _conv(ctx, img, "Conv1", 2, "same"), "bn_Conv1", M, T.relu6, virtual=True)   # -> stride-1 depthwise
    last = 5 if output_stride == 8 else 12
    for cout, stride, exp, bid in _MNV2:
        p = f"block_{bid}_" if bid else "expanded_conv_"
        inp, cin = x, x.shape[-1]
        if bid:
            x = _cbn(ctx, _conv(ctx, x, p + "expand"), p + "expand_BN", M, T.relu6, virtual=(stride == 1))
        k = ctx.w[p + "depthwise/depthwise_kernel"]
        if stride == 2:
            h, w = x.shape[1], x.shape[2]
            x = T.zero_pad2d(x, ((1 - (1 - h % 2), 1), (1 - (1 - w % 2), 1)))     # keras correct_pad
            x = T.depthwise_conv2d(x, k, 2, "valid")
        else:
            x = T.depthwise_conv2d(x, k, 1, "same")
        x = _cbn(ctx, x, p + "depthwise_BN", M, T.relu6, stats_rounded=True)   # stats kernel reads the stored tensor
        x = _cbn(ctx, _conv(ctx, x, p + "project"), p + "project_BN", M,
                 add=inp if (cin == cout and stride == 1) else None)
        if bid == last:
            return x
    raise AssertionError


def forward(conf: dict, weights: Dict[str, torch.Tensor], image, training: bool = False,
            dropout_mask: Optional[torch.Tensor] = None, emulate_bf16: bool = False,
            momentum_override: Optional[float] = None, probe: Optional[Probe] = None):
    """Returns dict(logits=[B,h,w,C] low-res, probs=[B,H,W,C], l2=regularisation term, new_stats={...}).
    `dropout_mask` (keep mask / (1-rate), shape of the concat) is required when training with dropout_rate > 0.
    `probe`: test instrumentation (named storage points, teacher forcing, forced decisions), see Probe."""
    arch, hps = conf["nn_arch"], conf["hps"]
    ctx = Ctx(weights, training, {}, emulate_bf16, momentum_override, probe)
    q = ctx.q
    osd = arch["output_stride"]
    base_fn = {"xception": xception_base, "mobilenetv2": mobilenetv2_base}[conf["base_model"]]
    names_after_base = None

    base_cache = []

    def run_base(x):
        nonlocal names_after_base
        if probe is not None and base_cache:
            # instrumented runs evaluate the shared base once (as the product does): the named storage points then
            # exist once and receive the SUM of both call sites' gradients.  (Moving statistics are then updated once;
            # the double update is checked by the plain oracle runs.)
            return base_cache[0]
        saved = ctx.names
        ctx.names = Names()                     # the base's automatic names do not depend on the call site
        y = base_fn(ctx, x, osd)
        names_after_base = ctx.names
        ctx.names = saved
        base_cache.append(y)
        return y

    feats = run_base(image)
    ctx.names = names_after_base                # head layers continue the counters of the Keras application
    nm = ctx.names
    mom, sc, wd = hps["bn_momentum"], hps["bn_scale"], hps["weight_decay"]
    width, mult = arch["reduction_size"], arch["conv_rate_multiplier"]

    def project(x):
        c, b = nm("conv2d"), nm("batch_normalization")
        return _cbn(ctx, _conv(ctx, x, c, 1, "same", wd), b, mom, T.relu, scale=sc)

    branches = []
    for spec in arch["encoder_middle_conf"]:
        src = feats if spec["input"] == -1 else branches[spec["input"]]
        if spec["op"] == "conv" and spec["kernel"] == 1:
            out = project(src)
        elif spec["op"] == "conv":
            s, b = nm("separable_conv2d"), nm("batch_normalization")
            rate = (spec["rate"][0] * mult, spec["rate"][1] * mult)
            out = _cbn(ctx, _sep(ctx, src, s, rate), b, mom, T.relu, scale=sc)
            out = project(out)
        elif spec["op"] == "pyramid_pooling":
            out = project(ctx.store(nm("average_pooling2d") + "/out", T.avg_pool_valid(src, spec["kernel"])))
            out = ctx.store(nm("lambda") + "/out", T.resize_bilinear(out, *spec["target_size_factor"]))
        else:
            raise ValueError("Invalid operation.")
        branches.append(out)
    x = ctx.store(nm("concatenate") + "/out", torch.cat(branches, dim=-1))
    dname = nm("dropout")
    if training and arch["dropout_rate"] > 0:
        if dropout_mask is None:
            raise ValueError("training with dropout needs an explicit mask for parity")
        x = ctx.store(dname + "/out", x * dropout_mask)
    enc = project(x)

    if arch["boundary_refinement"]:
        low = run_base(image)                   # shared weights, second pass (ss.py:930)
        low = project(low)
        f = int(osd / 2)
        low_up = ctx.store(nm("lambda") + "/out", T.resize_bilinear(low, f, f))
        enc_up = ctx.store(nm("lambda") + "/out", T.resize_bilinear(enc, f, f))
        x = ctx.store(nm("concatenate") + "/out", torch.cat([low_up, enc_up], dim=-1))
        up = int(osd / 8 if osd == 16 else osd / 4)
    else:
        x, up = enc, osd
    lname = nm("conv2d")
    logits = ctx.store(lname + "/y", _conv(ctx, x, lname, 1, "same", wd), rounded=False)
    probs = T.softmax(T.resize_bilinear(logits, up, up))
    l2 = sum(ctx.l2_terms) if ctx.l2_terms else torch.zeros((), dtype=image.dtype)
    return dict(logits=logits, probs=probs, l2=l2, new_stats=ctx.new_stats, encoder=enc, features=feats)


def loss_and_grads(conf, weights, image, labels, pos_w, neg_w, eps=1e-7, dropout_mask=None, wrt_logits=False,
                   emulate_bf16=False, probe: Optional[Probe] = None):
    """Training-mode forward + backward through autograd: returns (data_loss, l2, grads dict, forward dict)."""
    ws = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("moving_mean", "moving_variance"))
              else v) for k, v in weights.items()}
    out = forward(conf, ws, image, training=True, dropout_mask=dropout_mask, emulate_bf16=emulate_bf16, probe=probe)
    C = out["probs"].shape[-1]
    y = T.one_hot(labels, C, out["probs"].dtype)
    data = T.class_balanced_loss(y, out["probs"], pos_w, neg_w, eps)
    if wrt_logits:
        out["logits"].retain_grad()
    (data + out["l2"]).backward()
    grads = {k: v.grad for k, v in ws.items() if v.requires_grad and v.grad is not None}
    return data.detach(), out["l2"].detach(), grads, out

"""Teacher-forced / decision-forced parity harness (tests only).

`run(conf, ...)` executes ONE training step of the product (engine.Plan with keep_scratch, every macro-op's dy / dA
kept), reads every stored tensor and gradient buffer the plan exposes by name (Plan.traced) and replays the oracle
graph (oracle/model.py) twice:

1. TEACHER-FORCED: at every storage point the product has, the oracle's tensor — computed from the product's own
   stored inputs — is compared with the product's and replaced by it (forward), and the same for the gradient that
   arrives there (backward).  Every operation of the real engine wiring (depthwise stage, pointwise GEMM + statistics,
   BN + activation + residual, pools, resizes, concat slices, dropout, fused decoder tail, and all their backward
   kernels) is thereby checked on IDENTICAL inputs: no amplification through 40 layers, no decision flips (the masks
   derive from the same stored tensors).  Parameter gradients of the teacher-forced run are local too.
2. DECISION-FORCED: the oracle runs freely from the image but takes every ReLU/ReLU6 mask and max-pool winner from
   the product: both then differentiate the same piecewise-smooth function, so the whole-graph gradient can be held
   to the north-star tolerance, and the decisions on which the two disagree are listed explicitly (they must sit
   inside the forward-error band around a tie).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from oracle import model as OM
from tests import util


def rms_rel(a, b) -> float:
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dev(a, b) -> Dict[str, float]:
    """rms-relative deviation and the 99.99 % quantile of |a-b| relative to max|b|."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    d = (a - b).abs()
    scale = float(b.abs().max().clamp_min(1e-30))
    if d.numel() > 4_000_000:
        d = d[:: d.numel() // 4_000_000 + 1]
    q = float(torch.quantile(d, 0.9999)) if d.numel() > 1 else float(d.max())
    return {"rms": float((a - b).norm() / b.norm().clamp_min(1e-30)), "q9999": q / scale, "max": float(d.max()) / scale,
            "ref_rms": float(b.norm() / max(b.numel(), 1) ** 0.5)}


def run(conf, B=2, pw=None, nw=None, Plan=None, seed=1024, decision_forced=True, plan_kwargs=None, step_counter=0,
        fp32_floor=False, bf16_floor=False):
    from deeplabv3plus_keras_b200 import engine
    from tests.test_ops_gpu import NW, PW
    pw, nw = pw or PW, nw or NW
    Plan = Plan or engine.Plan
    bf16 = conf["hps"]["dtype"] == "bfloat16"
    ss = util.build(conf)
    util.randomize_weights(ss.model, seed=seed)
    plan = Plan(ss.model, B, training=True, keep_scratch=True, **(plan_kwargs or {}))
    x, y = util.synthetic_batch(conf, B, plan.out_shape[1:3], seed=seed)
    plan.set_loss(pw, nw)
    plan.load_batch(x, y)
    if step_counter:
        plan.step_counter.fill_(step_counter)
    plan.step_fwd_bwd()
    plan.regularization()
    if plan.device.type == "cuda":
        torch.cuda.synchronize()

    cpu = lambda t: t.detach().to("cpu", torch.float64)
    teacher = {k: cpu(plan.traced(k, "value")) for k in plan.trace if plan.traced(k, "value") is not None}
    teacher_grad = {k: cpu(plan.traced(k, "grad")) for k in plan.trace if plan.traced(k, "grad") is not None}
    w = util.torch_weights(ss.model)
    xin = torch.from_numpy(x)
    xin = (xin.to(torch.bfloat16) if bf16 else xin).double()
    yt = torch.from_numpy(y)
    drop = None
    if conf["nn_arch"]["dropout_rate"] > 0:
        (name,) = plan.dropout_sites
        drop = cpu(plan.dropout_mask(name))
    got_grads = {k: torch.from_numpy(v).double() for k, v in plan.gradients().items()}
    lam = conf["hps"]["weight_decay"]

    regularised = {f"{l.name}/kernel" for l in ss.model.flat_layers() if getattr(l, "kernel_regularizer", None) is not None}

    def minus_l2(grads):
        out = {}
        for k, g in grads.items():
            g = g.detach().clone()
            if k in regularised:
                g -= 2 * lam * w[k]                 # the product applies the L2 term inside the Adam kernel
            out[k] = g
        return out

    def split_zero(grads):
        """Gradients that are analytically zero (a beta in front of another batch-normalised convolution) come out as
        rounding noise on both sides: compared absolutely against the typical gradient magnitude, not relatively."""
        typical = float(np.median([float(g.abs().max()) for g in grads.values()]))
        live = {k: g for k, g in grads.items() if float(g.abs().max()) > 1e-4 * typical}
        dead = {k: float(got_grads[k].abs().max()) / typical for k in grads if k not in live}
        return live, dead

    res = dict(ss=ss, plan=plan, x=x, y=y, teacher=teacher, teacher_grad=teacher_grad, grads=got_grads)
    masks, taps = plan.decisions()
    masks = {k: v.cpu() for k, v in masks.items()}
    taps = {k: v.cpu() for k, v in taps.items()}
    # ---- 1. teacher-forced (on the product's decisions too: a BN -> ReLU mask is decided from batch statistics that the
    # product accumulates in fp32 atomics and the oracle in fp64, so an element within one ulp of zero can flip even on
    # identical stored inputs — one such element moves a 10^5-term random-sign sum like dbeta by 1/sqrt(N))
    probe = OM.Probe(teacher=teacher, teacher_grad=teacher_grad, masks=masks, pool_taps=taps, keep_values=False)
    d_tf, l2, g_tf, out_tf = OM.loss_and_grads(conf, w, xin, yt, pw, nw, dropout_mask=drop, emulate_bf16=bf16, probe=probe)
    res["unused_teacher"] = sorted(set(teacher) - set(probe.fwd))
    res["fwd"] = {k: dev(o, t) for k, (o, t) in probe.fwd.items()}
    res["bwd"] = {k: dev(t, o) for k, (o, t) in probe.bwd.items()}          # product vs oracle
    res["missing_grad_points"] = sorted(set(probe.fwd) - set(probe.bwd))
    live, res["param_zero"] = split_zero(minus_l2(g_tf))
    res["param_tf"] = {k: dev(got_grads[k], g) for k, g in live.items()}
    res["loss_tf"] = (plan.loss_value(), float(d_tf + l2))
    res["stats_tf"] = out_tf["new_stats"]
    if not decision_forced:
        return res
    # ---- 2. decision-forced, free running
    probe2 = OM.Probe(masks=masks, pool_taps=taps, keep_values=False)
    d_df, l2b, g_df, out_df = OM.loss_and_grads(conf, w, xin, yt, pw, nw, dropout_mask=drop, emulate_bf16=bf16, probe=probe2)
    res["unused_sites"] = sorted((set(masks) - set(probe2.pre)) | (set(taps) - set(probe2.pool_arg)))
    res["unforced_sites"] = sorted((set(probe2.pre) - set(masks)) | (set(probe2.pool_arg) - set(taps)))
    flips = {}
    for site, z in probe2.pre.items():
        m = masks.get(site)
        if m is None:
            continue
        relu6 = plan.act_sites[site][1] == engine.ACT_RELU6
        nat = (z > 0).to(torch.int8) + ((z >= 6).to(torch.int8) if relu6 else 0)
        bad = nat != m
        scale = float(z.abs().mean().clamp_min(1e-30))
        near = torch.minimum(z.abs(), (z - 6).abs()) if relu6 else z.abs()
        flips[site] = dict(count=int(bad.sum()), total=bad.numel(),
                           worst_margin=float((near[bad] / scale).max()) if bool(bad.any()) else 0.0)
    for site, arg in probe2.pool_arg.items():
        t = taps.get(site)
        if t is None:
            continue
        bad = arg != t.long()
        mg = probe2.pool_margin[site]
        scale = float(mg.mean().clamp_min(1e-30))
        flips[site] = dict(count=int(bad.sum()), total=bad.numel(),
                           worst_margin=float((mg[bad] / scale).max()) if bool(bad.any()) else 0.0)
    res["flips"] = flips
    live, dead = split_zero(minus_l2(g_df))
    res["param_df"] = {k: dev(got_grads[k], g) for k, g in live.items()}
    res["param_zero"].update({k: max(v, res["param_zero"].get(k, 0.0)) for k, v in dead.items()})
    res["logits_df"] = dev(cpu(plan.logits.buf[..., :plan.logits.clog]), out_df["logits"].detach())
    res["loss_df"] = (plan.loss_value(), float(d_df + l2b))
    res["out_df"] = out_df
    if bf16_floor:
        # what bf16 STORAGE alone does to this graph: the same bf16-rounding oracle, same forced decisions, with the
        # weights perturbed by 1e-7 (i.e. another fp32 summation order): rounding directions flip and the 2^-9 noise of
        # every stored tensor accumulates through the depth — no implementation can be closer to another than this
        gen = torch.Generator().manual_seed(0)
        w2 = {k: v * (1 + 1e-7 * torch.randn(v.shape, generator=gen, dtype=v.dtype)) for k, v in w.items()}
        probe4 = OM.Probe(masks=masks, pool_taps=taps, keep_values=False)
        _, _, _, out_p = OM.loss_and_grads(conf, w2, xin, yt, pw, nw, dropout_mask=drop, emulate_bf16=True, probe=probe4)
        res["logits_bf16_floor"] = dev(out_p["logits"].detach(), out_df["logits"].detach())
    if fp32_floor:
        # what fp32 ARITHMETIC alone does to this graph: the same oracle, same forced decisions, evaluated in float32
        # instead of float64 — the conditioning of the network (small BatchNormalization populations, cancelling
        # reductions), independent of the product
        probe3 = OM.Probe(masks=masks, pool_taps=taps, keep_values=False)
        w32 = {k: v.float() for k, v in w.items()}
        _, _, g32, out32 = OM.loss_and_grads(conf, w32, xin.float(), yt, pw, nw,
                                             dropout_mask=None if drop is None else drop.float(),
                                             emulate_bf16=bf16, probe=probe3)
        g32 = {k: v.double() for k, v in g32.items()}
        live32, _ = split_zero({k: v - (2 * lam * w[k] if k in regularised else 0) for k, v in g32.items()})
        res["param_floor"] = {k: dev(live32[k], live[k]) for k in live if k in live32}
        res["logits_floor"] = dev(out32["logits"].detach().double(), out_df["logits"].detach())
    return res


def worst(table: Dict[str, Dict[str, float]], key="rms"):
    if not table:
        return ("-", 0.0)
    k = max(table, key=lambda n: table[n][key])
    return k, table[k][key]


def summarize(res) -> str:
    lines = []
    for part in ("fwd", "bwd", "param_tf", "param_df"):
        if part in res:
            t = res[part]
            k, v = worst(t)
            med = float(np.median([e["rms"] for e in t.values()])) if t else 0.0
            lines.append(f"{part}: {len(t)} tensors, median rms-rel {med:.2e}, worst {v:.2e} at {k}, "
                         f"worst q99.99 {worst(t, 'q9999')[1]:.2e} at {worst(t, 'q9999')[0]}")
    if "logits_df" in res:
        lines.append(f"decision-forced whole graph: logits rms-rel {res['logits_df']['rms']:.2e}, loss "
                     f"{res['loss_df'][0]:.6f} vs {res['loss_df'][1]:.6f}")
    if "logits_bf16_floor" in res:
        lines.append(f"bf16-storage floor of the oracle itself (weights perturbed by 1e-7): logits rms-rel "
                     f"{res['logits_bf16_floor']['rms']:.2e}")
    if "param_floor" in res:
        t = res["param_floor"]
        lines.append(f"fp32-arithmetic floor of the oracle itself: median {float(np.median([e['rms'] for e in t.values()])):.2e}, "
                     f"worst {worst(t)[1]:.2e} at {worst(t)[0]}, logits {res['logits_floor']['rms']:.2e}")
    if "flips" in res:
        n = sum(f["count"] for f in res["flips"].values())
        tot = sum(f["total"] for f in res["flips"].values())
        wm = max((f["worst_margin"] for f in res["flips"].values()), default=0.0)
        lines.append(f"decisions: {n} of {tot} differ from the oracle's natural ones, worst relative margin {wm:.2e}")
    return "\n".join(lines)

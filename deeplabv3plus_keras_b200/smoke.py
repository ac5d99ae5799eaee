"""__graft_entry__.smoke(): one small forward+backward of the DeepLabV3+ hot path on cuda:0, checked against the
oracle (CPU restatement).  Imports oracle/ only as the checker."""
from __future__ import annotations

import copy
import warnings

import numpy as np
import torch

ASPP = [
    {"kernel": 3, "rate": [1, 1], "op": "conv", "input": -1},
    {"kernel": 3, "rate": [6, 6], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [12, 12], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [18, 18], "op": "conv", "input": 0},
    {"kernel": 1, "rate": [1, 1], "op": "pyramid_pooling", "input": 0, "target_size_factor": [1, 1]},
]


def small_conf(dtype="bfloat16", image_size=97, base="xception"):
    return {
        "mode": "train", "resource_path": "", "model_loading": False, "base_model": base, "base_weights": None,
        "hps": {"dtype": dtype, "lr": 1e-4, "beta_1": 0.5, "beta_2": 0.99, "decay": 0.0, "epochs": 1, "batch_size": 2,
                "weight_decay": 4e-5, "bn_momentum": 0.9, "bn_scale": True, "reduce_lr_factor": 0.99},
        "nn_arch": {"boundary_refinement": False, "output_stride": 16, "image_size": image_size, "num_classes": 21,
                    "mv2_depth_multiplier": 1, "depth_multiplier": 1, "conv_rate_multiplier": 1,
                    "reduction_size": 256, "dropout_rate": 0.0, "concat_channels": 256,
                    "encoder_middle_conf": copy.deepcopy(ASPP)},
    }


def run(verbose: bool = True) -> dict:
    from oracle import model as OM
    from . import keras
    from .deeplab import SemanticSegmentation, ss_nw, ss_pw
    from .engine import Plan

    torch.cuda.set_device(0)
    conf = small_conf()
    keras.reset_uids()
    keras.set_random_seed(1024)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ss = SemanticSegmentation(conf)
    rng = np.random.default_rng(1024)
    for l in ss.model.flat_layers():          # He-style weights so activations stay O(1)
        vals = []
        for n in l.weight_names():
            w = l._weights[n]
            if n in ("kernel", "pointwise_kernel"):
                v = rng.normal(0, np.sqrt(2.0 / (w.shape[0] * w.shape[1] * w.shape[2])), w.shape)
            elif n == "depthwise_kernel":
                v = rng.normal(0, np.sqrt(2.0 / 9.0), w.shape)
            elif n in ("gamma", "moving_variance"):
                v = 1.0 + 0.1 * np.abs(rng.normal(size=w.shape))
            else:
                v = 0.1 * rng.normal(size=w.shape)
            vals.append(v.astype(np.float32))
        l.set_weights(vals)
    B = 2
    plan = Plan(ss.model, B, training=True)
    Ho, Wo = plan.out_shape[1:3]
    x = rng.uniform(-1, 1, (B, 97, 97, 3)).astype(np.float32)
    y = rng.integers(0, 21, (B, Ho, Wo)).astype(np.int32)
    plan.set_loss(ss_pw, ss_nw)
    plan.load_batch(x, y)
    plan.step_fwd_bwd()
    torch.cuda.synchronize()
    loss = plan.loss_value()

    w = {k: torch.from_numpy(v.copy()) for k, v in ss.model.named_weights().items()}
    xb = torch.from_numpy(x).to(torch.bfloat16).float()      # the product rounds the image to bf16
    data, l2, grads, out = OM.loss_and_grads(conf, w, xb, torch.from_numpy(y), ss_pw, ss_nw)
    ref_logits = out["logits"].detach().numpy()
    got_logits = plan.logits.buf.float().cpu().numpy()
    denom = np.abs(ref_logits).max()
    err = float(np.abs(got_logits - ref_logits).max() / denom)
    rel_loss = abs(loss - float(data)) / max(abs(float(data)), 1e-6)
    gk = "block1_conv1/kernel"
    gref = grads[gk].numpy()
    ggot = plan.gradients()[gk]
    gerr = float(np.abs(ggot - gref).max() / max(np.abs(gref).max(), 1e-12))
    res = dict(loss=loss, oracle_loss=float(data), logits_err=err, loss_rel_err=rel_loss, grad_err_first_layer=gerr,
               launches=plan.launches_fwd + plan.launches_bwd)
    if verbose:
        print("smoke:", res)
    assert np.isfinite(loss), "non-finite loss"
    assert err < 5e-2, f"bf16 logits deviate from the oracle by {err:.3e} of max|logit|"
    assert rel_loss < 2e-2, f"loss {loss} vs oracle {float(data)}"
    assert gerr < 0.15, f"first-layer gradient deviates by {gerr:.3e}"
    return res

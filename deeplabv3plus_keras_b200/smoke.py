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


def _one(dtype: str, verbose: bool) -> dict:
    from oracle import model as OM
    from . import keras
    from .deeplab import SemanticSegmentation, ss_nw, ss_pw
    from .engine import Plan

    conf = small_conf(dtype)
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

    bf16 = dtype == "bfloat16"
    w = {k: torch.from_numpy(v.copy()).double() for k, v in ss.model.named_weights().items()}
    xin = torch.from_numpy(x)
    xin = (xin.to(torch.bfloat16) if bf16 else xin).double()
    data, l2, grads, out = OM.loss_and_grads(conf, w, xin, torch.from_numpy(y), ss_pw, ss_nw, emulate_bf16=bf16)
    ref = out["logits"].detach().numpy()
    got = plan.logits.buf.float().cpu().numpy()
    rms = float(np.sqrt(((got - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))
    rel_loss = abs(loss - float(data)) / max(abs(float(data)), 1e-6)
    gk = "block1_conv1/kernel"
    gref, ggot = grads[gk].numpy(), plan.gradients()[gk]
    gerr = float(np.sqrt(((ggot - gref) ** 2).mean()) / max(np.sqrt((gref ** 2).mean()), 1e-30))
    res = dict(dtype=dtype, loss=loss, oracle_loss=float(data), logits_rms_rel=rms, loss_rel_err=rel_loss,
               grad_rms_rel_first_layer=gerr, launches=plan.launches_fwd + plan.launches_bwd)
    if verbose:
        print("smoke:", res)
    assert np.isfinite(loss), "non-finite loss"
    # bf16: storage noise of a 40-layer random-init net is chaotic (tests/test_model_gpu.py measures the floor)
    assert rms < (0.25 if bf16 else 1e-3), f"{dtype} logits deviate from the oracle by {rms:.3e} rms-relative"
    assert rel_loss < (2e-2 if bf16 else 1e-4), f"loss {loss} vs oracle {float(data)}"
    if not bf16:
        assert gerr < 5e-2, f"first-layer gradient deviates by {gerr:.3e}"
    return res


def run(verbose: bool = True) -> dict:
    """bf16 (tcgen05 tensor-core path) and fp32 (strict parity) training steps against the oracle."""
    torch.cuda.set_device(0)
    return {"bfloat16": _one("bfloat16", verbose), "float32": _one("float32", verbose)}

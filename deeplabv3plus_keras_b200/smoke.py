"""__graft_entry__.smoke(): one small training step (forward + fused loss + hand-written backward) of the DeepLabV3+
hot path on cuda:0 in the benchmarked bf16 tensor-core mode and in fp32, checked against the oracle at the north-star
tolerances (BASELINE.json: fp32 1e-3, bf16 2e-2, gradients at the same tolerance).  Imports oracle/ and the test
harness tests/teacher.py only as the checker.

What is asserted (tests/teacher.py explains the three views):
  * teacher-forced: EVERY stored tensor, gradient buffer and parameter gradient of the engine schedule within the
    tolerance of the oracle evaluated on the product's own stored inputs (each kernel in its real wiring, bf16 too);
  * fp32, decision-forced whole graph: logits, loss and every parameter gradient within 1e-3 of the free-running fp64
    oracle that takes the product's ReLU / max-pool decisions, and those decisions differ from the oracle's own only at
    near-ties;
  * bf16 whole graph: loss within 2e-2.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

ASPP = [
    {"kernel": 3, "rate": [1, 1], "op": "conv", "input": -1},
    {"kernel": 3, "rate": [6, 6], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [12, 12], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [18, 18], "op": "conv", "input": 0},
    {"kernel": 1, "rate": [1, 1], "op": "pyramid_pooling", "input": 0, "target_size_factor": [1, 1]},
]
FP32_TOL, BF16_TOL = 1e-3, 2e-2


def small_conf(dtype="bfloat16", image_size=97, base="xception"):
    return {
        "mode": "train", "resource_path": "", "model_loading": False, "base_model": base, "base_weights": None,
        "hps": {"dtype": dtype, "lr": 1e-4, "beta_1": 0.5, "beta_2": 0.99, "decay": 0.0, "epochs": 1, "batch_size": 2,
                "weight_decay": 4e-5, "bn_momentum": 0.9, "bn_scale": True, "reduce_lr_factor": 0.99},
        "nn_arch": {"boundary_refinement": False, "output_stride": 16, "image_size": image_size, "num_classes": 21,
                    "mv2_depth_multiplier": 1, "depth_multiplier": 1, "conv_rate_multiplier": 1,
                    "reduction_size": 256, "dropout_rate": 0.5, "concat_channels": 256,
                    "encoder_middle_conf": copy.deepcopy(ASPP)},
    }


def _one(dtype: str, verbose: bool) -> dict:
    from tests import teacher

    bf16 = dtype == "bfloat16"
    tol = BF16_TOL if bf16 else FP32_TOL
    res = teacher.run(small_conf(dtype), B=2)
    worst = {part: teacher.worst(res[part]) for part in ("fwd", "bwd", "param_tf", "param_df")}
    flips = sum(f["count"] for f in res["flips"].values())
    sites = sum(f["total"] for f in res["flips"].values())
    out = dict(dtype=dtype, loss=res["loss_tf"][0], oracle_loss_teacher_forced=res["loss_tf"][1],
               oracle_loss_whole_graph=res["loss_df"][1], stored_tensors_checked=len(res["fwd"]),
               gradient_buffers_checked=len(res["bwd"]), parameter_gradients_checked=len(res["param_tf"]),
               worst_forward_rms_rel=worst["fwd"][1], worst_backward_rms_rel=worst["bwd"][1],
               worst_param_grad_rms_rel_teacher_forced=worst["param_tf"][1],
               worst_param_grad_rms_rel_whole_graph=worst["param_df"][1],
               whole_graph_logits_rms_rel=res["logits_df"]["rms"], decisions_differing=flips, decisions=sites,
               launches=res["plan"].launches_fwd + res["plan"].launches_bwd, tolerance=tol)
    if verbose:
        print("smoke:", out)
    assert np.isfinite(out["loss"]), "non-finite loss"
    assert not res["unused_teacher"] and not res["missing_grad_points"], "oracle and product disagree on the storage points"
    for part in ("fwd", "bwd", "param_tf"):
        name, v = worst[part]
        assert v <= tol, f"{dtype} {part}: {name} deviates {v:.3e} (> {tol}) from the oracle on identical inputs"
    a, b = res["loss_tf"]
    assert abs(a - b) <= (2e-3 if bf16 else 1e-5) * max(1.0, abs(b)), f"loss {a} vs teacher-forced oracle {b}"
    a, b = res["loss_df"]
    assert abs(a - b) <= (BF16_TOL if bf16 else 1e-4) * max(1.0, abs(b)), f"loss {a} vs whole-graph oracle {b}"
    if not bf16:
        name, v = worst["param_df"]
        assert v <= FP32_TOL, f"fp32 whole-graph gradient of {name} deviates {v:.3e}"
        assert res["logits_df"]["rms"] <= FP32_TOL
        assert flips <= 1e-4 * sites and all(f["worst_margin"] <= 1e-4 for f in res["flips"].values())
    return out


def run(verbose: bool = True) -> dict:
    """bf16 (tcgen05 tensor-core path, the benchmarked one) and fp32 training steps against the oracle."""
    torch.cuda.set_device(0)
    return {"bfloat16": _one("bfloat16", verbose), "float32": _one("float32", verbose)}

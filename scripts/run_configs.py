"""BASELINE.json configs other than the headline one, at FULL size on one B200: runs them through the public surface
(SemanticSegmentation -> Model.plan / Trainer) and prints one JSON line per config with throughput (CUDA events, >= 3
warm-ups) and the size-independent checks (output geometry of the reference graph, finite values, probabilities that
sum to one, label range).  The headline config (cfg-2) and cfg-3 (data parallel) are bench.py's job.

  cfg-1  MobileNetV2 + default-JSON ASPP (asymmetric rates, chained branches) + decoder, 513x513, batch 1, fp32 inference
  cfg-4  Xception OS8, conv_rate_multiplier 2 (ASPP 2/12/24/36), boundary refinement, 513x513, batch 16, bf16 fwd+bwd
  cfg-5  MobileNetV2 OS16, 1024x2048, 19 classes, batch 8, bf16 inference
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests import util

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()


def timed(fn, steps, warmup=3, stream=None):
    """CUDA events on the stream the kernels are launched on (the Trainer owns its stream)."""
    stream = stream or torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def inference(name, conf, batch, dtype):
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    plan = ss.model.plan(batch, training=False, dtype=dtype)
    x, _ = util.synthetic_batch(conf, batch, plan.out_shape[1:3])
    plan.load_batch(x)
    probs = plan.predict_device()
    torch.cuda.synchronize()
    s = probs.sum(-1)
    labels = plan.segment(x)
    ms = timed(plan.predict_device, args.steps)
    size = conf["nn_arch"]["image_size"]
    print(json.dumps({"config": name, "mode": "inference", "dtype": dtype, "batch": batch, "image_size": size,
                      "out_shape": list(plan.out_shape), "ms_per_batch": ms, "img_per_s": batch / ms * 1e3,
                      "finite": bool(torch.isfinite(probs).all()), "prob_sum_err": float((s - 1).abs().max()),
                      "label_range": [int(labels.min()), int(labels.max())],
                      "launches_fwd": plan.launches_fwd}), flush=True)


def training(name, conf, batch, dtype):
    from deeplabv3plus_keras_b200.trainer import Trainer
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    tr = Trainer(ss.model, batch, dtype=dtype, use_graph=True)
    x, y = util.synthetic_batch(conf, batch, tr.plan.out_shape[1:3])
    xs, ys = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    losses = [tr.train_step_e2e(xs, ys) for _ in range(4)]
    ms = timed(tr.step, args.steps, stream=tr.stream)
    g = tr.plan.params.g[:tr.plan.params.n_train]
    print(json.dumps({"config": name, "mode": "fwd+bwd+adam", "dtype": dtype, "batch": batch,
                      "out_shape": list(tr.plan.out_shape), "ms_per_step": ms, "img_per_s": batch / ms * 1e3,
                      "losses": losses, "finite": bool(np.all(np.isfinite(losses)) and torch.isfinite(g).all()),
                      "grad_nonzero_frac": float((g != 0).float().mean()),
                      "launches_per_step": tr.launches_per_step}), flush=True)


if not args.only or "cfg1" in args.only:
    inference("cfg-1 MobileNetV2 OS16 513^2 default-JSON ASPP", util.make_conf(
        base="mobilenetv2", output_stride=16, image_size=513, aspp=util.DEFAULT_ASPP, dtype="float32"), 1, "float32")
if not args.only or "cfg4" in args.only:
    training("cfg-4 Xception OS8 rate x2 + boundary refinement 513^2", util.make_conf(
        base="xception", output_stride=8, image_size=513, refine=True, rate_mult=2, dtype="bfloat16", dropout=0.5),
        16, "bfloat16")
if not args.only or "cfg5" in args.only:
    conf = util.make_conf(base="mobilenetv2", output_stride=16, image_size=[1024, 2048], num_classes=19,
                          dtype="bfloat16")
    rng = np.random.default_rng(7)
    f = rng.dirichlet(np.ones(19))
    conf["class_weights"] = {"pos": list(1.0 - f), "neg": list(f)}     # ss.py:401-404: pw = 1 - freq, nw = freq
    inference("cfg-5 MobileNetV2 OS16 1024x2048 19 classes", conf, 8, "bfloat16")

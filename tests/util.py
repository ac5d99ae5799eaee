"""Shared helpers for the parity tests: configs in the reference's JSON schema, seeded weights / inputs."""
import copy
import warnings

import numpy as np
import torch

XCEPTION_ASPP = [
    {"kernel": 3, "rate": [1, 1], "op": "conv", "input": -1},
    {"kernel": 3, "rate": [6, 6], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [12, 12], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [18, 18], "op": "conv", "input": 0},
    {"kernel": 1, "rate": [1, 1], "op": "pyramid_pooling", "input": 0, "target_size_factor": [1, 1]},
]
DEFAULT_ASPP = [   # conf.json:39-45 (asymmetric rates, chained branches)
    {"kernel": 3, "rate": [1, 1], "op": "conv", "input": -1},
    {"kernel": 3, "rate": [18, 15], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [6, 3], "op": "conv", "input": 1},
    {"kernel": 3, "rate": [1, 1], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [6, 21], "op": "conv", "input": 0},
]


def global_pool_aspp(feat: int, sub: int = 0):
    """An ASPP exercising what the shipped JSONs leave identity: a `conv` k=1 branch (ss.py:812-820), a TRUE image-pooling
    branch (AveragePooling2D over the whole feat x feat map, 1x1 conv, bilinear x feat: ss.py:841-856) and a second
    pyramid level (pool `sub`, x`sub`) chained behind branch 0."""
    sub = sub or next(d for d in range(2, feat + 1) if feat % d == 0)     # AveragePooling2D is VALID: must divide
    return [
        {"kernel": 1, "rate": [1, 1], "op": "conv", "input": -1},
        {"kernel": 3, "rate": [6, 6], "op": "conv", "input": -1},
        {"kernel": 3, "rate": [12, 12], "op": "conv", "input": 0},
        {"kernel": feat, "rate": [1, 1], "op": "pyramid_pooling", "input": -1, "target_size_factor": [feat, feat]},
        {"kernel": sub, "rate": [1, 1], "op": "pyramid_pooling", "input": 0, "target_size_factor": [sub, sub]},
    ]


def feature_size(base: str, output_stride: int, image_size: int) -> int:
    """Spatial extent of the backbone tap: Xception's two VALID convs give 513 -> 32 at OS16."""
    n = image_size
    if base == "xception":
        n = (n - 3) // 2 + 1          # block1_conv1 3x3 s2 VALID
        n = n - 2                     # block1_conv2 3x3 VALID
        for _ in range({8: 2, 16: 3}[output_stride]):
            n = -(-n // 2)
        return n
    for _ in range({8: 3, 16: 4}[output_stride]):
        n = -(-n // 2)
    return n


def make_conf(base="xception", output_stride=16, image_size=65, refine=False, dtype="float32", aspp=None,
              num_classes=21, dropout=0.0, rate_mult=1, width=256):
    if aspp == "default":
        aspp = DEFAULT_ASPP
    elif aspp == "global_pool":
        aspp = global_pool_aspp(feature_size(base, output_stride, image_size))
    return {
        "mode": "train", "resource_path": "", "model_loading": False, "base_model": base, "base_weights": None,
        "hps": {"dtype": dtype, "lr": 1e-4, "beta_1": 0.5, "beta_2": 0.99, "decay": 0.0, "epochs": 1,
                "batch_size": 1, "weight_decay": 4e-5, "bn_momentum": 0.9, "bn_scale": True, "reduce_lr_factor": 0.99},
        "nn_arch": {"boundary_refinement": refine, "output_stride": output_stride, "image_size": image_size,
                    "num_classes": num_classes, "mv2_depth_multiplier": 1, "depth_multiplier": 1,
                    "conv_rate_multiplier": rate_mult, "reduction_size": width, "dropout_rate": dropout,
                    "concat_channels": width,
                    "encoder_middle_conf": copy.deepcopy(aspp if aspp is not None else XCEPTION_ASPP)},
    }


def build(conf):
    from deeplabv3plus_keras_b200 import keras
    from deeplabv3plus_keras_b200.deeplab import SemanticSegmentation
    keras.reset_uids()
    keras.set_random_seed(1024)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return SemanticSegmentation(conf)


def randomize_weights(model, seed=1024):
    """He-style kernels so activations stay O(1) through ~40 layers; BN gamma~1, beta~0, mean~0, var~1 perturbed."""
    rng = np.random.default_rng(seed)
    for l in model.flat_layers():
        vals = []
        for n in l.weight_names():
            w = l._weights[n]
            if n in ("kernel", "pointwise_kernel"):
                fan_in = w.shape[0] * w.shape[1] * w.shape[2]
                v = rng.normal(0, np.sqrt(2.0 / fan_in), w.shape)
            elif n == "depthwise_kernel":
                v = rng.normal(0, np.sqrt(2.0 / 9.0), w.shape)
            elif n == "gamma":
                v = 1.0 + 0.1 * rng.normal(size=w.shape)
            elif n == "beta":
                v = 0.1 * rng.normal(size=w.shape)
            elif n == "moving_mean":
                v = 0.1 * rng.normal(size=w.shape)
            elif n == "moving_variance":
                v = 1.0 + 0.1 * np.abs(rng.normal(size=w.shape))
            else:
                raise KeyError(n)
            vals.append(v.astype(np.float32))
        l.set_weights(vals)


def synthetic_batch(conf, batch, out_hw, seed=1024):
    """images ~ U(-1,1) (the reference's normalisation range, ss.py:1532); labels: rectangles over background."""
    rng = np.random.default_rng(seed)
    size = conf["nn_arch"]["image_size"]
    h, w = (size, size) if isinstance(size, int) else size
    C = conf["nn_arch"]["num_classes"]
    x = rng.uniform(-1, 1, (batch, h, w, 3)).astype(np.float32)
    y = np.zeros((batch,) + tuple(out_hw), dtype=np.int32)
    for b in range(batch):
        for _ in range(rng.integers(2, 6)):
            c = int(rng.integers(1, C))
            y0, x0 = rng.integers(0, out_hw[0]), rng.integers(0, out_hw[1])
            y1, x1 = y0 + rng.integers(1, out_hw[0] // 2 + 2), x0 + rng.integers(1, out_hw[1] // 2 + 2)
            y[b, y0:y1, x0:x1] = c
    return x, y


def torch_weights(model, dtype=torch.float64):
    return {k: torch.from_numpy(v.copy()).to(dtype) for k, v in model.named_weights().items()}

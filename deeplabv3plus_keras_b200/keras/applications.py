"""Topologies of tf.keras.applications.Xception and MobileNetV2 (TF 2.4) with the Keras layer names the
reference truncates at (ss.py:502-504,518-520).  The Keras-Applications source is NOT part of /root/reference
(un-vendored dependency); the topologies are restated from the published architecture and pinned by the
well-known parameter totals (Xception no-top 20,861,480; MobileNetV2 alpha=1 no-top 2,257,984) in
tests/test_oracle.py.

`weights`: None = random initialisation; a path to an `.npz` written by scripts/convert_keras_weights.py (arrays keyed
"layer/weight" in Keras layouts); "imagenet" = that file looked up as `$DLV3P_PRETRAINED_DIR/<name>_imagenet_notop.npz`
(default `~/.keras/models`).  There is no network here and tf.keras' download cannot run: "imagenet" without a
converted file RAISES — it never falls back to random weights silently (the reference relies on the pretrained
backbone, ss.py:496-499 / 512-515).
"""
from __future__ import annotations

import os

import numpy as np

from . import layers as L
from .base import Input
from .models import Model

PRETRAINED_FILES = {"Xception": "xception_imagenet_notop.npz", "MobileNetV2": "mobilenet_v2_1.0_imagenet_notop.npz"}


def _resolve_weights(weights, who):
    """None, or the path of the converted `.npz` to load after the graph is built."""
    if weights is None:
        return None
    if weights == "imagenet":
        root = os.environ.get("DLV3P_PRETRAINED_DIR") or os.path.join(os.path.expanduser("~"), ".keras", "models")
        path = os.path.join(root, PRETRAINED_FILES[who])
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{who}(weights='imagenet'): {path} not found.  Pretrained weights cannot be downloaded here; convert the "
                f"Keras-Applications file on a machine that has it (python scripts/convert_keras_weights.py --model "
                f"{who.lower()} --out {path}) or pass weights=None (conf['base_weights'] = None) for random initialisation")
        return path
    if isinstance(weights, str) and weights.endswith(".npz"):
        if not os.path.exists(weights):
            raise FileNotFoundError(f"{who}: weights file {weights} not found")
        return weights
    raise ValueError(f"{who}: weights must be None, 'imagenet' or the path of a converted .npz file")


def _load_pretrained(model, path, who):
    """Copy every array of the file into the layer of the same name (layers of the classification top that the
    no-top graph does not have are ignored; a layer of the graph missing from the file, or a shape mismatch, raises)."""
    data = np.load(path)
    have = {}
    for k in data.files:
        have.setdefault(k.rsplit("/", 1)[0], {})[k.rsplit("/", 1)[1]] = data[k]
    for layer in model.flat_layers():
        names = layer.weight_names()
        if not names:
            continue
        if layer.name not in have:
            raise ValueError(f"{who}: {path} holds no weights for layer {layer.name!r}")
        layer.set_weights([have[layer.name][n] for n in names])


def _check_no_top(include_top, who):
    if include_top:
        raise ValueError(f"{who}: include_top=True is outside the DeepLabV3+ hot path")


def Xception(include_top=False, weights="imagenet", input_tensor=None, input_shape=None, pooling=None, classes=1000):
    _check_no_top(include_top, "Xception")
    pretrained = _resolve_weights(weights, "Xception")
    img = input_tensor if input_tensor is not None else Input(shape=input_shape, name="input_1")

    def conv_bn_act(x, filters, name, strides=1):
        x = L.Conv2D(filters, (3, 3), strides=(strides, strides), use_bias=False, name=name)(x)
        x = L.BatchNormalization(name=name + "_bn")(x)
        return L.Activation("relu", name=name + "_act")(x)

    def sep_bn(x, filters, name, pre_act):
        if pre_act:
            x = L.Activation("relu", name=name + "_act")(x)
        x = L.SeparableConv2D(filters, (3, 3), padding="same", use_bias=False, name=name)(x)
        return L.BatchNormalization(name=name + "_bn")(x)

    def strided_shortcut(x, filters):
        r = L.Conv2D(filters, (1, 1), strides=(2, 2), padding="same", use_bias=False)(x)
        return L.BatchNormalization()(r)

    x = conv_bn_act(img, 32, "block1_conv1", strides=2)
    x = conv_bn_act(x, 64, "block1_conv2")

    # entry flow: blocks 2-4, each = strided 1x1 shortcut || two separable convs + maxpool
    for blk, filters in ((2, 128), (3, 256), (4, 728)):
        shortcut = strided_shortcut(x, filters)
        x = sep_bn(x, filters, f"block{blk}_sepconv1", pre_act=(blk != 2))
        x = sep_bn(x, filters, f"block{blk}_sepconv2", pre_act=True)
        x = L.MaxPooling2D((3, 3), strides=(2, 2), padding="same", name=f"block{blk}_pool")(x)
        x = L.add([x, shortcut])

    # middle flow: blocks 5-12, identity shortcut around three pre-activated separable convs
    for blk in range(5, 13):
        shortcut = x
        for j in (1, 2, 3):
            x = sep_bn(x, 728, f"block{blk}_sepconv{j}", pre_act=True)
        x = L.add([x, shortcut])

    # exit flow
    shortcut = strided_shortcut(x, 1024)
    x = sep_bn(x, 728, "block13_sepconv1", pre_act=True)
    x = sep_bn(x, 1024, "block13_sepconv2", pre_act=True)
    x = L.MaxPooling2D((3, 3), strides=(2, 2), padding="same", name="block13_pool")(x)
    x = L.add([x, shortcut])
    for j, filters in ((1, 1536), (2, 2048)):
        x = L.SeparableConv2D(filters, (3, 3), padding="same", use_bias=False, name=f"block14_sepconv{j}")(x)
        x = L.BatchNormalization(name=f"block14_sepconv{j}_bn")(x)
        x = L.Activation("relu", name=f"block14_sepconv{j}_act")(x)
    model = Model(img, x, name="xception")
    if pretrained:
        _load_pretrained(model, pretrained, "Xception")
    return model


def _make_divisible(v, divisor, min_value=None):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def _correct_pad(shape, kernel_size=3):
    h, w = shape[1], shape[2]
    adj = (1 - h % 2, 1 - w % 2)
    c = kernel_size // 2
    return ((c - adj[0], c), (c - adj[1], c))


MNV2_BLOCKS = (  # (filters, stride, expansion, block_id)
    (16, 1, 1, 0), (24, 2, 6, 1), (24, 1, 6, 2), (32, 2, 6, 3), (32, 1, 6, 4), (32, 1, 6, 5), (64, 2, 6, 6),
    (64, 1, 6, 7), (64, 1, 6, 8), (64, 1, 6, 9), (96, 1, 6, 10), (96, 1, 6, 11), (96, 1, 6, 12), (160, 2, 6, 13),
    (160, 1, 6, 14), (160, 1, 6, 15), (320, 1, 6, 16))


def MobileNetV2(input_shape=None, alpha=1.0, include_top=False, weights="imagenet", input_tensor=None, pooling=None,
                classes=1000):
    _check_no_top(include_top, "MobileNetV2")
    pretrained = _resolve_weights(weights, "MobileNetV2")
    img = input_tensor if input_tensor is not None else Input(shape=input_shape, name="input_1")
    bn = dict(epsilon=1e-3, momentum=0.999)

    x = L.Conv2D(_make_divisible(32 * alpha, 8), 3, strides=(2, 2), padding="same", use_bias=False, name="Conv1")(img)
    x = L.BatchNormalization(name="bn_Conv1", **bn)(x)
    x = L.ReLU(6.0, name="Conv1_relu")(x)

    for filters, stride, expansion, bid in MNV2_BLOCKS:
        cin = x.shape[-1]
        cout = _make_divisible(int(filters * alpha), 8)
        prefix = f"block_{bid}_" if bid else "expanded_conv_"
        inp = x
        if bid:
            x = L.Conv2D(expansion * cin, 1, padding="same", use_bias=False, name=prefix + "expand")(x)
            x = L.BatchNormalization(name=prefix + "expand_BN", **bn)(x)
            x = L.ReLU(6.0, name=prefix + "expand_relu")(x)
        if stride == 2:
            x = L.ZeroPadding2D(padding=_correct_pad(x.shape, 3), name=prefix + "pad")(x)
        x = L.DepthwiseConv2D(3, strides=stride, use_bias=False, padding="same" if stride == 1 else "valid",
                              name=prefix + "depthwise")(x)
        x = L.BatchNormalization(name=prefix + "depthwise_BN", **bn)(x)
        x = L.ReLU(6.0, name=prefix + "depthwise_relu")(x)
        x = L.Conv2D(cout, 1, padding="same", use_bias=False, name=prefix + "project")(x)
        x = L.BatchNormalization(name=prefix + "project_BN", **bn)(x)
        if cin == cout and stride == 1:
            x = L.Add(name=prefix + "add")([inp, x])

    last = _make_divisible(1280 * alpha, 8) if alpha > 1.0 else 1280
    x = L.Conv2D(last, 1, use_bias=False, name="Conv_1")(x)
    x = L.BatchNormalization(name="Conv_1_bn", **bn)(x)
    x = L.ReLU(6.0, name="out_relu")(x)
    model = Model(img, x, name=f"mobilenetv2_{alpha:0.2f}_{input_shape[0] if input_shape else 'None'}")
    if pretrained:
        _load_pretrained(model, pretrained, "MobileNetV2")
    return model

"""Layer constructors with the tf.keras signatures the reference uses:
SeparableConv2D (ss.py:823-830), Conv2D (ss.py:814-818,...), BatchNormalization (ss.py:819,...), Activation,
AveragePooling2D (ss.py:842), Lambda (ss.py:852-856), Concatenate (ss.py:863), Dropout (ss.py:864), plus what the
keras.applications backbones need (DepthwiseConv2D, ZeroPadding2D, ReLU, MaxPooling2D, Add).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

from . import base
from .base import KTensor, Layer, Node


# ---- initializers / regularizers -------------------------------------------------------------------
class initializers:
    class Initializer:
        def __call__(self, shape):
            raise NotImplementedError

    class GlorotUniform(Initializer):
        def __call__(self, shape):
            if len(shape) == 4:
                rf = shape[0] * shape[1]
                fan_in, fan_out = shape[2] * rf, shape[3] * rf
            else:
                fan_in = fan_out = int(np.prod(shape))
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return base.rng().uniform(-lim, lim, size=shape)

    class TruncatedNormal(Initializer):
        def __init__(self, mean=0.0, stddev=0.05):
            self.mean, self.stddev = mean, stddev

        def __call__(self, shape):
            r = base.rng()
            x = r.normal(size=shape)
            bad = np.abs(x) > 2.0
            while bad.any():
                x[bad] = r.normal(size=int(bad.sum()))
                bad = np.abs(x) > 2.0
            return self.mean + self.stddev * x

    class Zeros(Initializer):
        def __call__(self, shape):
            return np.zeros(shape)

    class Ones(Initializer):
        def __call__(self, shape):
            return np.ones(shape)


class regularizers:
    class L2:
        def __init__(self, l2=0.01):
            self.l2 = float(l2)

    @staticmethod
    def l2(l=0.01):
        return regularizers.L2(l)


def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        if len(v) != 2:
            raise ValueError(f"expected an int or a pair, got {v}")
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _conv_out(n, k, s, d, padding):
    k_eff = (k - 1) * d + 1
    if padding == "same":
        return -(-n // s)
    return (n - k_eff) // s + 1


# ---- convolutions ---------------------------------------------------------------------------------
class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", dilation_rate=(1, 1), use_bias=True,
                 kernel_initializer=None, kernel_regularizer=None, activation=None, name=None, **kw):
        super().__init__(name=name, **kw)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.padding = padding.lower()
        self.dilation_rate = _pair(dilation_rate)
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer or initializers.GlorotUniform()
        self.kernel_regularizer = kernel_regularizer
        if activation is not None:
            raise ValueError("Conv2D(activation=...) is not on the reference's path; use an Activation layer")
        if self.kernel_size[0] != self.kernel_size[1] or self.kernel_size[0] not in (1, 3):
            raise ValueError(f"Conv2D kernel_size {self.kernel_size}: only 1x1 and 3x3 are on the hot path")
        if self.strides[0] != self.strides[1]:
            raise ValueError("Conv2D: anisotropic strides unsupported")
        if self.padding not in ("same", "valid"):
            raise ValueError(f"Conv2D padding {padding!r}")

    def build(self, shapes):
        cin = shapes[0][-1]
        self.add_weight("kernel", self.kernel_size + (cin, self.filters), self.kernel_initializer)
        if self.use_bias:
            self.add_weight("bias", (self.filters,), initializers.Zeros())

    def compute_output_shape(self, shapes):
        n, h, w, _ = shapes[0]
        k, s, d = self.kernel_size[0], self.strides[0], self.dilation_rate
        return (n, _conv_out(h, k, s, d[0], self.padding), _conv_out(w, k, s, d[1], self.padding), self.filters)


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, strides=(1, 1), padding="valid", depth_multiplier=1, dilation_rate=(1, 1),
                 use_bias=True, depthwise_initializer=None, activation=None, name=None, **kw):
        super().__init__(name=name, **kw)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.padding = padding.lower()
        self.dilation_rate = _pair(dilation_rate)
        self.use_bias = use_bias
        self.depthwise_initializer = depthwise_initializer or initializers.GlorotUniform()
        if depth_multiplier != 1 or self.kernel_size != (3, 3) or activation is not None:
            raise ValueError("DepthwiseConv2D: only 3x3, depth_multiplier=1, no activation is on the hot path")
        if use_bias:
            raise ValueError("DepthwiseConv2D(use_bias=True) is not on the hot path")

    def build(self, shapes):
        self.add_weight("depthwise_kernel", (3, 3, shapes[0][-1], 1), self.depthwise_initializer)

    def compute_output_shape(self, shapes):
        n, h, w, c = shapes[0]
        s, d = self.strides[0], self.dilation_rate
        return (n, _conv_out(h, 3, s, d[0], self.padding), _conv_out(w, 3, s, d[1], self.padding), c)


class SeparableConv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", depth_multiplier=1,
                 dilation_rate=(1, 1), use_bias=True, kernel_initializer=None, depthwise_initializer=None,
                 pointwise_initializer=None, activation=None, name=None, **kw):
        super().__init__(name=name, **kw)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.padding = padding.lower()
        self.dilation_rate = _pair(dilation_rate)
        self.use_bias = use_bias
        # tf.keras SeparableConv2D has no `kernel_initializer`; the reference passes one (ss.py:830) and TF 2.4
        # swallows it in **kwargs without applying it — accepted and ignored here for the same reason.
        self.depthwise_initializer = depthwise_initializer or initializers.GlorotUniform()
        self.pointwise_initializer = pointwise_initializer or initializers.GlorotUniform()
        if depth_multiplier != 1 or self.kernel_size != (3, 3) or activation is not None:
            raise ValueError("SeparableConv2D: only 3x3, depth_multiplier=1, no activation is on the hot path")
        if self.strides != (1, 1) and self.dilation_rate != (1, 1):
            raise ValueError("SeparableConv2D: strides > 1 with dilation_rate > 1 is invalid (as in tf.keras)")
        if use_bias:
            raise ValueError("SeparableConv2D(use_bias=True) is not on the hot path")

    def build(self, shapes):
        cin = shapes[0][-1]
        self.add_weight("depthwise_kernel", (3, 3, cin, 1), self.depthwise_initializer)
        self.add_weight("pointwise_kernel", (1, 1, cin, self.filters), self.pointwise_initializer)

    def compute_output_shape(self, shapes):
        n, h, w, _ = shapes[0]
        s, d = self.strides[0], self.dilation_rate
        return (n, _conv_out(h, 3, s, d[0], self.padding), _conv_out(w, 3, s, d[1], self.padding), self.filters)


class ZeroPadding2D(Layer):
    def __init__(self, padding=((1, 1), (1, 1)), name=None, **kw):
        super().__init__(name=name, **kw)
        if isinstance(padding, int):
            padding = ((padding, padding), (padding, padding))
        self.padding = (tuple(padding[0]), tuple(padding[1]))

    def compute_output_shape(self, shapes):
        n, h, w, c = shapes[0]
        (pt, pb), (pl, pr) = self.padding
        return (n, h + pt + pb, w + pl + pr, c)


# ---- normalisation / activation -------------------------------------------------------------------
class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, name=None, **kw):
        super().__init__(name=name, **kw)
        if axis not in (-1, 3):
            raise ValueError("BatchNormalization: channels_last only")
        self.momentum, self.epsilon, self.center, self.scale = float(momentum), float(epsilon), center, scale

    def build(self, shapes):
        c = shapes[0][-1]
        if self.scale:
            self.add_weight("gamma", (c,), initializers.Ones())
        if self.center:
            self.add_weight("beta", (c,), initializers.Zeros())
        self.add_weight("moving_mean", (c,), initializers.Zeros(), trainable=False)
        self.add_weight("moving_variance", (c,), initializers.Ones(), trainable=False)


class Activation(Layer):
    def __init__(self, activation, name=None, **kw):
        super().__init__(name=name, **kw)
        if activation not in ("relu", "softmax", "linear"):
            raise ValueError(f"Activation({activation!r}): only 'relu' and 'softmax' are on the hot path")
        self.activation = activation


class ReLU(Layer):
    _default_prefix = "re_lu"

    def __init__(self, max_value=None, name=None, **kw):
        super().__init__(name=name, **kw)
        if max_value not in (None, 6, 6.0):
            raise ValueError("ReLU(max_value): only None or 6 supported")
        self.max_value = None if max_value is None else 6.0


class Dropout(Layer):
    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name, **kw)
        self.rate = float(rate)
        if not 0.0 <= self.rate < 1.0:
            raise ValueError(f"Dropout rate {rate}")


# ---- pooling / merge / resize ----------------------------------------------------------------------
class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **kw):
        super().__init__(name=name, **kw)
        self.pool_size, self.strides, self.padding = _pair(pool_size), _pair(strides or pool_size), padding.lower()
        if (self.pool_size, self.strides, self.padding) != ((3, 3), (2, 2), "same"):
            raise ValueError("MaxPooling2D: only (3,3)/strides 2/'same' (Xception) is on the hot path")

    def compute_output_shape(self, shapes):
        n, h, w, c = shapes[0]
        return (n, -(-h // 2), -(-w // 2), c)


class AveragePooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **kw):
        super().__init__(name=name, **kw)
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides) if strides is not None else self.pool_size
        self.padding = padding.lower()
        if self.pool_size[0] != self.pool_size[1] or self.strides != self.pool_size or self.padding != "valid":
            raise ValueError("AveragePooling2D: square pool, strides == pool_size, padding='valid' only (ss.py:842)")

    def compute_output_shape(self, shapes):
        n, h, w, c = shapes[0]
        k = self.pool_size[0]
        return (n, h // k, w // k, c)


class Add(Layer):
    def compute_output_shape(self, shapes):
        if any(s != shapes[0] for s in shapes) or len(shapes) != 2:
            raise ValueError(f"Add: need two tensors of equal shape, got {shapes}")
        return shapes[0]


def add(inputs, name=None):
    return Add(name=name)(inputs)


class Concatenate(Layer):
    def __init__(self, axis=-1, name=None, **kw):
        super().__init__(name=name, **kw)
        if axis not in (-1, 3):
            raise ValueError("Concatenate: channel axis only")

    def compute_output_shape(self, shapes):
        if any(s[:-1] != shapes[0][:-1] for s in shapes):
            raise ValueError(f"Concatenate: spatial shapes differ: {shapes}")
        return shapes[0][:-1] + (sum(s[-1] for s in shapes),)


class ResizeImages(Layer):
    """The op behind K.resize_images(x, hf, wf, 'channels_last', interpolation='bilinear')."""

    def __init__(self, height_factor, width_factor, interpolation="bilinear", name=None, **kw):
        super().__init__(name=name, **kw)
        self.factors = (int(height_factor), int(width_factor))
        if interpolation != "bilinear":
            raise ValueError("resize_images: only 'bilinear' is on the hot path")
        if min(self.factors) < 1:
            raise ValueError(f"resize_images factors {self.factors}")

    def compute_output_shape(self, shapes):
        n, h, w, c = shapes[0]
        return (n, h * self.factors[0], w * self.factors[1], c)


class Lambda(Layer):
    """Lambda(function): the function is traced once on the symbolic tensor; only backend ops that build
    layers (K.resize_images) may appear inside (that is all the reference uses Lambda for)."""

    def __init__(self, function, name=None, **kw):
        super().__init__(name=name, **kw)
        self.function = function

    def __call__(self, inputs):
        out = self.function(inputs)
        if not isinstance(out, KTensor):
            raise TypeError("Lambda function must return a symbolic tensor built from backend ops")
        if out.node.layer is not getattr(inputs, "node", None) and isinstance(out.node.layer, ResizeImages) \
                and out.node.inputs == [inputs]:
            out.node.layer.name = self.name       # tf.keras lists the Lambda ("lambda", "lambda_1", ...) in model.layers
        return out

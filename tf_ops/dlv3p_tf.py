"""TensorFlow side of the custom-op route (`tf.load_op_library`): gradients and Keras layers that keep the
constructor signatures the reference uses (ss.py:795-954) while running libdlv3p kernels.

Usage on a machine with TensorFlow >= 2.4 and a B200 (after `bash tf_ops/build.sh`):

    import tf_ops.dlv3p_tf as dlv3p_tf
    # in bodhi/deeplabv3plus_keras/semantic_segmentation.py replace
    #   from tensorflow.keras.layers import SeparableConv2D        (ss.py:52-57)
    # by
    #   from tf_ops.dlv3p_tf import SeparableConv2D

TensorFlow is not installable in the build image, so this module is exercised only by
tests/test_tf_ops.py, which skips when `import tensorflow` fails.
"""
import os

import tensorflow as tf
from tensorflow.python.framework import ops as _ops

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = tf.load_op_library(os.path.join(_HERE, "libdlv3p_tf_ops.so"))

ACT = {None: 0, "linear": 0, "relu": 1, "relu6": 2}


@_ops.RegisterGradient("Dlv3pDepthwiseConv3x3")
def _dw_grad(op, dy):
    a = {k: op.get_attr(k) for k in ("stride", "dilation_h", "dilation_w", "padding", "pre_activation")}
    x, w = op.inputs
    dx = _lib.dlv3p_depthwise_conv3x3_backprop_input(x=x, filter=w, dy=dy, **a)
    dw = _lib.dlv3p_depthwise_conv3x3_backprop_filter(x=x, dy=dy, **a)
    return dx, dw


@_ops.RegisterGradient("Dlv3pPointwiseConv")
def _pw_grad(op, dy):
    """Valid for the un-fused form (no scale/shift, no activation): dX = dY W, dW^T = (X^T dY)^T."""
    x, w_t, scale, shift = op.inputs
    k = tf.shape(w_t)[1]
    w = tf.transpose(w_t)                                    # [K, Cout] = B operand [N=K, K=Cout] of the dgrad GEMM
    dx = _lib.dlv3p_pointwise_conv(x=dy, w_t=w, scale=tf.zeros([0]), shift=tf.zeros([0]), activation=0)
    dw = _lib.dlv3p_pointwise_conv_backprop_filter(x=x, dy=dy)          # [K, Cout] fp32
    return dx, tf.cast(tf.transpose(dw), w_t.dtype), tf.zeros_like(scale), tf.zeros_like(shift)


class SeparableConv2D(tf.keras.layers.SeparableConv2D):
    """Drop-in for tf.keras.layers.SeparableConv2D(filters, 3, depth_multiplier=1, dilation_rate, padding='same',
    use_bias=False) as constructed at ss.py:823-830: same weights (`depthwise_kernel`, `pointwise_kernel`), the
    arithmetic runs in Dlv3pDepthwiseConv3x3 + Dlv3pPointwiseConv."""

    def call(self, inputs):
        if self.kernel_size != (3, 3) or self.depth_multiplier != 1 or self.use_bias or inputs.dtype != tf.bfloat16:
            return super().call(inputs)                       # outside the hot path: stock TF
        d = _lib.dlv3p_depthwise_conv3x3(x=inputs, filter=tf.cast(self.depthwise_kernel, tf.float32),
                                         stride=self.strides[0], dilation_h=self.dilation_rate[0],
                                         dilation_w=self.dilation_rate[1], padding=self.padding.upper(),
                                         pre_activation=0)
        w_t = tf.cast(tf.transpose(self.pointwise_kernel[0, 0]), tf.bfloat16)
        return _lib.dlv3p_pointwise_conv(x=d, w_t=w_t, scale=tf.zeros([0]), shift=tf.zeros([0]), activation=0)


@tf.custom_gradient
def upsample_softmax_class_balanced_loss(logits, labels, pos_weights, neg_weights, factor, epsilon=1e-7):
    """mean_{b,h,w} class_balanced_loss(onehot(labels), softmax(resize_bilinear(logits, x factor))) — the fused form
    of ss.py:904-909 + 438-447 for integer label maps."""
    loss_sum, dlogits = _lib.dlv3p_upsample_softmax_cb_loss(logits=logits, labels=labels, pos_weights=pos_weights,
                                                           neg_weights=neg_weights, factor=factor, epsilon=epsilon)
    shp = tf.shape(labels)
    loss = loss_sum / tf.cast(shp[0] * shp[1] * shp[2], tf.float32)

    def grad(upstream):
        return upstream * dlogits, None, None, None, None, None
    return loss, grad

"""ORACLE (test infrastructure only — never imported by the product package).

CPU restatement, in torch float64/float32, of the TensorFlow 2.4 raw ops that the reference's Keras layers
execute on the DeepLabV3+ hot path (reference: bodhi/deeplabv3plus_keras/semantic_segmentation.py, cited per
function as ss.py:LINE).  TensorFlow is an un-vendored third-party dependency of the reference (README.md:8,15
"tensorflow==2.4", unpinned in requirements.txt:1-6) and is NOT installable in this image, so:

    PARITY UNPINNED — this oracle follows the published TF 2.4 op semantics (SURVEY.md §8c checklist) but has not
    been executed against TensorFlow itself; the reference ships no tests / golden vectors for this path.

What pins it instead: (1) oracle/np_ref.py, an independent pure-numpy loop restatement of the same formulas,
checked against this file in tests/test_oracle.py; (2) the Keras parameter-count invariants of the two backbones;
(3) scripts/dump_reference_tf.py, which regenerates tests/golden/*.npz from the real reference on a machine that
has TF 2.4.

All tensors are NHWC (TF default), conv kernels HWIO, depthwise kernels [kh,kw,C,1].
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import torch
import torch.nn.functional as F


def same_pad(n: int, k: int, stride: int, dil: int = 1) -> Tuple[int, int, int]:
    """TF 'SAME' geometry: (out, pad_before, pad_after); the extra padding goes to the bottom/right."""
    k_eff = (k - 1) * dil + 1
    out = -(-n // stride)
    total = max((out - 1) * stride + k_eff - n, 0)
    return out, total // 2, total - total // 2


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def _pad_same(x_nchw, kh, kw, stride, dil, value=0.0):
    _, pt, pb = same_pad(x_nchw.shape[2], kh, stride, dil[0])
    _, pl, pr = same_pad(x_nchw.shape[3], kw, stride, dil[1])
    return F.pad(x_nchw, (pl, pr, pt, pb), value=value)


def conv2d(x, w_hwio, stride: int = 1, padding: str = "same", dilation=(1, 1)):
    """tf.nn.conv2d (Keras Conv2D without bias; ss.py:814-818,893-897)."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    xn = _nchw(x)
    if padding == "same":
        xn = _pad_same(xn, kh, kw, stride, dilation)
    elif padding != "valid":
        raise ValueError(padding)
    w = w_hwio.permute(3, 2, 0, 1)
    return _nhwc(F.conv2d(xn, w, stride=stride, dilation=dilation))


def depthwise_conv2d(x, w_hwc1, stride: int = 1, padding: str = "same", dilation=(1, 1)):
    """tf.nn.depthwise_conv2d with depth_multiplier 1 (depthwise half of SeparableConv2D, ss.py:823-830)."""
    kh, kw, C, _ = w_hwc1.shape
    xn = _nchw(x)
    if padding == "same":
        xn = _pad_same(xn, kh, kw, stride, dilation)
    elif padding != "valid":
        raise ValueError(padding)
    w = w_hwc1.permute(2, 3, 0, 1)  # [C,1,kh,kw]
    return _nhwc(F.conv2d(xn, w, stride=stride, dilation=dilation, groups=C))


def zero_pad2d(x, pads):
    """Keras ZeroPadding2D(((top,bottom),(left,right))) — MobileNetV2 correct_pad before stride-2 depthwise."""
    (pt, pb), (pl, pr) = pads
    return _nhwc(F.pad(_nchw(x), (pl, pr, pt, pb)))


def batch_norm(x, gamma, beta, moving_mean, moving_var, eps: float, training: bool, momentum: float = 0.99):
    """FusedBatchNormV3.  Returns (y, new_moving_mean, new_moving_var).  Training normalises with the biased batch
    variance and updates the moving variance with the unbiased one (TF fused convention)."""
    if training:
        mean = x.mean(dim=(0, 1, 2))
        var = x.var(dim=(0, 1, 2), unbiased=False)
        n = x.shape[0] * x.shape[1] * x.shape[2]
        unb = var * (n / max(n - 1, 1))
        new_mm = momentum * moving_mean + (1 - momentum) * mean.detach()
        new_mv = momentum * moving_var + (1 - momentum) * unb.detach()
    else:
        mean, var = moving_mean, moving_var
        new_mm, new_mv = moving_mean, moving_var
    y = (x - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y, new_mm, new_mv


def relu(x):
    return torch.clamp_min(x, 0)


def relu6(x):
    return torch.clamp(x, 0, 6)


def max_pool_3x3_s2_same(x):
    """MaxPooling2D(3, strides=2, padding='same'): padding never wins (−inf)."""
    xn = _pad_same(_nchw(x), 3, 3, 2, (1, 1), value=float("-inf"))
    return _nhwc(F.max_pool2d(xn, 3, 2))


def avg_pool_valid(x, k: int):
    """AveragePooling2D(pool_size=k, padding='valid'), stride = k (ss.py:842)."""
    return _nhwc(F.avg_pool2d(_nchw(x), k, k))


def resize_bilinear(x, fh: int, fw: int):
    """K.resize_images(x, fh, fw, 'channels_last', interpolation='bilinear') (ss.py:852-856,904-908,941-950)
    = tf.image.resize(..., 'bilinear') = ResizeBilinear(align_corners=False, half_pixel_centers=True):
        in = (out + 0.5) * (in_size/out_size) - 0.5; lo = max(floor(in),0); hi = min(ceil(in), in_size-1);
        lerp = in - floor(in)
    written with explicit gathers so it does not depend on torch's interpolate."""
    N, H, W, C = x.shape

    def axis(n_in, f):
        o = torch.arange(n_in * f, dtype=torch.float64)
        src = (o + 0.5) / f - 0.5
        fl = torch.floor(src)
        lo = torch.clamp(fl, min=0).long()
        hi = torch.clamp(torch.ceil(src), max=n_in - 1).long()
        return lo, hi, (src - fl).to(x.dtype)

    y0, y1, ly = axis(H, fh)
    x0, x1, lx = axis(W, fw)
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * lx.view(1, 1, -1, 1)
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * lx.view(1, 1, -1, 1)
    return top + (bot - top) * ly.view(1, -1, 1, 1)


def softmax(x):
    return torch.softmax(x, dim=-1)


def class_balanced_loss(y_true, y_pred, pos_weights: Sequence[float], neg_weights: Sequence[float],
                        epsilon: float = 1e-7):
    """ss.py:438-447, statement for statement (python loop over classes, K.mean over every remaining axis)."""
    loss = 0.0
    for i in range(len(pos_weights)):
        loss = loss + -1.0 * (pos_weights[i] * y_true[..., i] * torch.log(y_pred[..., i] + epsilon)
                              + neg_weights[i] * (1.0 - y_true[..., i]) * torch.log(1.0 - y_pred[..., i] + epsilon))
    return loss.mean()


def one_hot(labels, num_classes: int, dtype=torch.float64):
    """get_one_hot (ss.py:337-362) without the python loop."""
    return F.one_hot(labels.long(), num_classes).to(dtype)


def argmax_labels(y):
    """K.argmax over the class axis (MeanIoUExt, ss.py:310-311; segment(), ss.py:1227): first maximum wins."""
    return torch.argmax(y, dim=-1)


def confusion_matrix(y_true_labels, y_pred_labels, num_classes: int):
    """tf.math.confusion_matrix(..., dtype=float64) as used by MeanIoUExt.update_state (ss.py:326-334)."""
    idx = y_true_labels.reshape(-1).long() * num_classes + y_pred_labels.reshape(-1).long()
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes).double()


def mean_iou(cm):
    """tf.keras.metrics.MeanIoU.result(): mean over classes with a non-zero denominator."""
    tp = torch.diagonal(cm)
    denom = cm.sum(0) + cm.sum(1) - tp
    valid = denom > 0
    iou = torch.where(valid, tp / torch.clamp(denom, min=1e-30), torch.zeros_like(tp))
    return iou.sum() / torch.clamp(valid.sum(), min=1)

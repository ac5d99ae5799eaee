"""DeepLabV3+ model-building surface — the drop-in boundary.

Mirrors `SemanticSegmentation` of the reference (bodhi/deeplabv3plus_keras/semantic_segmentation.py:450-954):
same constructor argument (the JSON `conf` dict, schema conf.json:1-54), same attributes (`base`, `encoder`,
`decoder`, `model`), same builder methods (`_make_encoder`, `_make_decoder`, `_refine_boundary`), `segment()`,
plus `ClassBalancedLoss` / `class_balanced_loss` (ss.py:423-447), `MeanIoUExt` (ss.py:283-334) and the VOC class
weights `ss_pw` / `ss_nw` (ss.py:120-127).  Only the hot path is mirrored: the Xception and MobileNetV2 base
models (the other `base_model` values raise), no dataset Sequences, callbacks, PNG dumps or TFLite export.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import keras
from .keras import (Activation, AveragePooling2D, BatchNormalization, Concatenate, Conv2D, Dropout, Input, Lambda,
                    Model, SeparableConv2D, initializers, regularizers)
from .keras import backend as K
from .keras.applications import MobileNetV2, Xception

BASE_MODEL_MOBILENETV2 = "mobilenetv2"
BASE_MODEL_XCEPTION = "xception"
_OUT_OF_SCOPE_BASES = ("efficientnetb0", "efficientnetb1", "efficientnetb2", "efficientnetb3", "efficientnetb4",
                       "efficientnetb5", "efficientnetb6", "efficientnetb7", "nasnetmobile", "nasnetlarge",
                       "densenet121", "densenet169", "densenet201")

# Pascal-VOC class-balance weights, values of ss.py:120-127 (pw = 1 - pixel frequency, nw = pixel frequency)
ss_pw = [0.29754999, 0.99106889, 0.99236374, 0.99122957, 0.99350396, 0.99455487,
         0.98728424, 0.98090446, 0.96883489, 0.98753125, 0.99376389, 0.98942612,
         0.97222875, 0.99080578, 0.98845309, 0.92606652, 0.99393374, 0.99374322,
         0.98782171, 0.98659656, 0.99233476]
ss_nw = [0.70245001, 0.00893111, 0.00763626, 0.00877043, 0.00649604, 0.00544513,
         0.01271576, 0.01909554, 0.03116511, 0.01246875, 0.00623611, 0.01057388,
         0.02777125, 0.00919422, 0.01154691, 0.07393348, 0.00606626, 0.00625678,
         0.01217829, 0.01340344, 0.00766524]

# backbone tap layers that give the requested output stride (ss.py:501-504, 517-520)
_TAPS = {
    BASE_MODEL_MOBILENETV2: {8: "block_5_add", 16: "block_12_add"},
    BASE_MODEL_XCEPTION: {8: "block4_sepconv2_bn", 16: "block13_sepconv2_bn"},
}


class Adam:
    """optimizers.Adam(lr, beta_1, beta_2, decay) as configured at ss.py:477-480 (Keras epsilon 1e-7;
    `decay` is the legacy inverse-time decay lr / (1 + decay * iterations))."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.0, learning_rate=None):
        self.lr = float(learning_rate if learning_rate is not None else lr)
        self.beta_1, self.beta_2, self.epsilon, self.decay = float(beta_1), float(beta_2), float(epsilon), float(decay)
        self.iterations = 0

    def step_size(self) -> float:
        """lr_t for the NEXT update (iterations is incremented by the caller after the update)."""
        t = self.iterations + 1
        lr = self.lr / (1.0 + self.decay * self.iterations)
        return lr * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)


class ReduceLROnPlateau:
    """keras.callbacks.ReduceLROnPlateau as configured at ss.py:978-982 (monitor='loss', factor=hps.reduce_lr_factor,
    patience=5, min_lr=1e-8; Keras defaults mode='auto' -> min, min_delta=1e-4, cooldown=0): call `on_epoch_end(loss)`
    once per epoch; the new learning rate is written to the optimizer and used by the next Adam launch."""

    def __init__(self, optimizer: "Adam", monitor="loss", factor=0.1, patience=10, min_lr=0.0, min_delta=1e-4,
                 cooldown=0, verbose=0):
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")        # Keras' own check
        self.optimizer, self.monitor, self.factor, self.patience = optimizer, monitor, float(factor), int(patience)
        self.min_lr, self.min_delta, self.cooldown, self.verbose = float(min_lr), float(min_delta), int(cooldown), verbose
        self.best, self.wait, self.cooldown_counter = float("inf"), 0, 0

    def on_epoch_end(self, value: float) -> float:
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if value < self.best - self.min_delta:
            self.best, self.wait = value, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = self.optimizer.lr
                if old > self.min_lr:
                    self.optimizer.lr = max(old * self.factor, self.min_lr)
                    if self.verbose:
                        print(f"ReduceLROnPlateau reducing learning rate to {self.optimizer.lr}.")
                    self.cooldown_counter = self.cooldown
                    self.wait = 0
        return self.optimizer.lr


class ClassBalancedLoss:
    """LossFunctionWrapper around class_balanced_loss (ss.py:423-435); serialisable by name."""

    def __init__(self, pos_weights=1.0, neg_weights=0.0, epsilon=1e-7, reduction="auto", name="class_balanced_loss"):
        self.pos_weights = [float(v) for v in np.atleast_1d(pos_weights)]
        self.neg_weights = [float(v) for v in np.atleast_1d(neg_weights)]
        if len(self.pos_weights) != len(self.neg_weights):
            raise ValueError("pos_weights and neg_weights must have the same length")
        self.epsilon, self.reduction, self.name = float(epsilon), reduction, name

    def get_config(self):
        return dict(pos_weights=self.pos_weights, neg_weights=self.neg_weights, epsilon=self.epsilon,
                    reduction=self.reduction, name=self.name)

    def __call__(self, y_true, y_pred):
        return class_balanced_loss(y_true, y_pred, self.pos_weights, self.neg_weights, self.epsilon)


def class_balanced_loss(y_true, y_pred, pos_weights=1.0, neg_weights=0.0, epsilon=1e-7):
    """mean over (b,h,w) of sum_i -(pw_i*y_i*log(p_i+eps) + nw_i*(1-y_i)*log(1-p_i+eps)) — ss.py:438-447 — on
    CUDA tensors or host arrays (copied to the device); computed by dlv3p_cbloss_dense_fwd.  Returns a float."""
    import torch

    from . import ops
    pw = [float(v) for v in np.atleast_1d(pos_weights)]
    nw = [float(v) for v in np.atleast_1d(neg_weights)]
    yt = torch.as_tensor(np.asarray(y_true) if not torch.is_tensor(y_true) else y_true).to("cuda", torch.float32)
    yp = torch.as_tensor(np.asarray(y_pred) if not torch.is_tensor(y_pred) else y_pred).to("cuda", torch.float32)
    C = yt.shape[-1]
    if len(pw) != C or len(nw) != C or tuple(yt.shape) != tuple(yp.shape):
        raise ValueError(f"class_balanced_loss: {len(pw)} weights for {C} classes / shape mismatch")
    P = yt.numel() // C
    out = torch.zeros(1, device="cuda")
    ops.cbloss_dense_fwd(yt.contiguous(), yp.contiguous(), torch.tensor(pw, device="cuda"),
                         torch.tensor(nw, device="cuda"), float(epsilon), P, C, out)
    return float(out.item()) / P


class MeanIoUExt:
    """MeanIoU over one-hot truth / prediction tensors (ss.py:283-334): argmax both, accumulate (or overwrite) a
    float64 confusion matrix on the device (dlv3p_softmax_argmax + dlv3p_confusion_matrix)."""

    def __init__(self, num_classes, accum_enable=True, name=None, dtype=None):
        self.num_classes, self.accum_enable, self.name = int(num_classes), accum_enable, name or "mean_io_u_ext"
        self.total_cm = None

    def reset_states(self):
        self.total_cm = None

    def update_state(self, y_true, y_pred, sample_weight=None):
        import torch

        from . import ops
        if sample_weight is not None:
            raise ValueError("MeanIoUExt: sample_weight is not supported on the device path")
        C = self.num_classes

        def labels_of(t):
            t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).to("cuda")
            if t.dtype in (torch.int32, torch.int64) and t.shape[-1] != C:
                return t.reshape(-1).to(torch.int32).contiguous()
            t = t.to(torch.float32).contiguous()
            P = t.numel() // C
            lab = torch.empty(P, dtype=torch.int32, device="cuda")
            ops.softmax_argmax(t, P, C, labels=lab)
            return lab

        lt, lp = labels_of(y_true), labels_of(y_pred)
        cm = torch.zeros((C, C), dtype=torch.float64, device="cuda")
        ops.confusion_matrix(lt, lp, lt.numel(), C, cm)
        self.total_cm = cm if (self.total_cm is None or not self.accum_enable) else self.total_cm + cm
        return self.total_cm

    def result(self) -> float:
        if self.total_cm is None:
            return 0.0
        cm = self.total_cm.cpu().numpy()
        tp = np.diag(cm)
        denom = cm.sum(0) + cm.sum(1) - tp
        valid = denom > 0
        return float((tp[valid] / denom[valid]).sum() / max(int(valid.sum()), 1))


class SemanticSegmentation:
    """Keras-style DeepLabV3+ (reference class at ss.py:450)."""

    MODEL_PATH = "semantic_segmentation_deeplabv3plus"

    def __init__(self, conf: Dict):
        arch, hps = conf["nn_arch"], conf["hps"]
        assert arch["output_stride"] in (8, 16)     # ss.py:468
        self.conf, self.hps, self.nn_arch = conf, hps, arch
        self.resource_path = conf.get("resource_path", "")
        self.model_loading = bool(conf.get("model_loading", False))
        opt = Adam(lr=hps["lr"], beta_1=hps["beta_1"], beta_2=hps["beta_2"], decay=hps["decay"])

        if self.model_loading:
            from .utils import load_weights_npz
            self._build(conf)
            load_weights_npz(self.model, f"{self.resource_path}/{self.MODEL_PATH}.npz")
        else:
            self._build(conf)
        n_cls = arch["num_classes"]
        pw, nw = (ss_pw, ss_nw) if n_cls == len(ss_pw) else (conf["class_weights"]["pos"], conf["class_weights"]["neg"])
        self.model.compile(optimizer=opt, loss=ClassBalancedLoss(pw, nw), metrics=[MeanIoUExt(num_classes=n_cls)])
        self.model._init_set_name("deeplabv3plus")

    # -- construction --------------------------------------------------------------------------------------
    def _image_shape(self):
        size = self.nn_arch["image_size"]          # int in the reference; [H, W] accepted as an extension
        h, w = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
        return (h, w, 3)

    def _build(self, conf):
        name = conf["base_model"]
        if name in _OUT_OF_SCOPE_BASES:
            raise NotImplementedError(
                f"base_model {name!r}: only 'xception' and 'mobilenetv2' are on the B200 hot path (SURVEY.md §2 #8)")
        if name not in _TAPS:
            raise ValueError("base model is not valid.")          # ss.py:771
        # the reference lets keras.applications default to weights='imagenet' (ss.py:496-499, 512-515); a resumed model
        # (model_loading) takes every weight from its checkpoint instead
        weights = None if self.model_loading else conf.get("base_weights", "imagenet")
        app = (MobileNetV2 if name == BASE_MODEL_MOBILENETV2 else Xception)(
            input_shape=self._image_shape(), include_top=False, weights=weights)
        tap = app.get_layer(_TAPS[name][self.nn_arch["output_stride"]]).output
        self.base = Model(inputs=app.inputs, outputs=tap)
        self.base.trainable = True
        for layer in self.base.layers:
            layer.trainable = True
        self.base._init_set_name("base")

        self._make_encoder()
        self._make_decoder()
        inputs = self.encoder.inputs
        features = self.encoder(inputs)
        if self.nn_arch["boundary_refinement"]:
            outputs = self.decoder([inputs[0], features])
        else:
            outputs = self.decoder(features)
        self.model = Model(inputs, outputs)

    def _bn(self):
        return BatchNormalization(momentum=self.hps["bn_momentum"], scale=self.hps["bn_scale"])

    def _l2(self):
        return regularizers.l2(self.hps["weight_decay"])

    def _project(self, x, channels, initializer=None):
        """1x1 Conv (no bias, L2) + BN + ReLU — the unit at ss.py:814-820 / 833-840 / 843-849 / 865-871 / 931-937."""
        kw = dict(kernel_initializer=initializer) if initializer is not None else {}
        x = Conv2D(channels, kernel_size=1, padding="same", use_bias=False, kernel_regularizer=self._l2(), **kw)(x)
        x = self._bn()(x)
        return Activation("relu")(x)

    def _make_encoder(self):
        """ASPP driven by nn_arch['encoder_middle_conf'] (ss.py:790-876)."""
        assert hasattr(self, "base")
        arch = self.nn_arch
        image = Input(shape=self._image_shape(), dtype=self.hps["dtype"], name="input_image")
        feats = self.base(image)
        mult = arch["conv_rate_multiplier"]
        width = arch["reduction_size"]

        branches: List = []
        for spec in arch["encoder_middle_conf"]:
            src = feats if spec["input"] == -1 else branches[spec["input"]]
            op = spec["op"]
            if op == "conv" and spec["kernel"] == 1:
                out = self._project(src, width)
            elif op == "conv":
                out = SeparableConv2D(width, spec["kernel"], depth_multiplier=1,
                                      dilation_rate=(spec["rate"][0] * mult, spec["rate"][1] * mult),
                                      padding="same", use_bias=False,
                                      kernel_initializer=initializers.TruncatedNormal())(src)
                out = self._bn()(out)
                out = Activation("relu")(out)
                out = self._project(out, width, initializer=initializers.TruncatedNormal())
            elif op == "pyramid_pooling":
                out = AveragePooling2D(pool_size=spec["kernel"], padding="valid")(src)
                out = self._project(out, width)
                fh, fw = spec["target_size_factor"]
                out = Lambda(lambda t, fh=fh, fw=fw: K.resize_images(t, fh, fw, "channels_last",
                                                                     interpolation="bilinear"))(out)
            else:
                raise ValueError("Invalid operation.")             # ss.py:858
            branches.append(out)

        merged = Concatenate(axis=-1)(branches)
        merged = Dropout(rate=arch["dropout_rate"])(merged)
        out = self._project(merged, arch["concat_channels"])
        self.encoder = Model(image, out)
        self.encoder._init_set_name("encoder")

    def _make_decoder(self):
        """logits conv 3x3 -> bilinear x output_stride (x2 after refinement) -> softmax (ss.py:878-913)."""
        assert hasattr(self, "base") and hasattr(self, "encoder")
        arch = self.nn_arch
        features = Input(shape=K.int_shape(self.encoder.outputs[0])[1:], dtype=self.hps["dtype"])
        refine = bool(arch["boundary_refinement"])
        if refine:
            low_features = Input(shape=K.int_shape(self.encoder.inputs[0])[1:])
            x = self._refine_boundary(low_features, features)
        else:
            x = features
        x = Conv2D(arch["num_classes"], kernel_size=3, padding="same", use_bias=False,
                   kernel_regularizer=self._l2())(x)
        stride = arch["output_stride"]
        factor = (stride // 8 if stride == 16 else stride // 4) if refine else stride    # ss.py:901-902
        x = Lambda(lambda t: K.resize_images(t, factor, factor, "channels_last", interpolation="bilinear"))(x)
        probs = Activation("softmax")(x)
        self.decoder = Model(inputs=[low_features, features] if refine else [features], outputs=probs)
        self.decoder._init_set_name("decoder")

    def _refine_boundary(self, low_features, features):
        """Second pass of the shared base over the image, 1x1->48, both maps x(output_stride/2), concat
        (ss.py:915-954)."""
        low = self._project(self.base(low_features), 48)
        f = int(self.nn_arch["output_stride"] / 2)
        up = lambda t: K.resize_images(t, f, f, "channels_last", interpolation="bilinear")  # noqa: E731
        return Concatenate(axis=-1)([Lambda(up)(low), Lambda(up)(features)])

    # -- inference -----------------------------------------------------------------------------------------
    def segment(self, images):
        """Label maps for a batch of images (ss.py:1207-1227): argmax over the class axis, on the device."""
        images = np.asarray(images)
        plan = self.model.plan(len(images), training=False)
        return plan.segment(images)

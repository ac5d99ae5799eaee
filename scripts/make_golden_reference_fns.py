"""Pins taken from the REFERENCE'S OWN SOURCE (the module itself cannot be imported here: it imports tensorflow /
cupy at the top, ss.py:38-47).  The source of the functions is read from /root/reference at generation time, compiled
and executed in this process — nothing is copied into the repo but the resulting arrays / graph descriptions:

1. `class_balanced_loss` (ss.py:438-447) and the legacy closure `class_imbalance_loss` (ss.py:410-420), executed with
   `K` bound to a numpy shim (K.log = np.log, K.mean = np.mean — the only backend calls they make)
   -> tests/golden/loss_*.npz
2. `MeanIoUExt.update_state` (ss.py:290-334), executed with K.argmax / math_ops.cast / array_ops.reshape /
   confusion_matrix.confusion_matrix bound to numpy shims of those TF ops (tf.math.confusion_matrix: cm[t, p] += w)
   and a stand-in for the tf.keras MeanIoU base class (total_cm variable with assign / assign_add)
   -> tests/golden/miou_*.npz
3. `SemanticSegmentation.__init__ / _make_encoder / _make_decoder / _refine_boundary` (ss.py:459-525, 770-954),
   executed verbatim with `Input, Conv2D, SeparableConv2D, BatchNormalization, Activation, AveragePooling2D, Lambda,
   Concatenate, Dropout, Model, K, regularizers, initializers, optimizers, Xception, MobileNetV2` bound to THIS
   repo's keras mirror: the graph the reference's code builds through the drop-in surface
   -> tests/golden/topology_*.json (tests/topology.py:describe), compared with deeplab.py's graph in
   tests/test_reference_pins.py.
4. `ss_pw` / `ss_nw` (ss.py:120-127) -> tests/golden/loss_voc_weights.npz.
"""
import ast
import functools
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = "/root/reference/bodhi/deeplabv3plus_keras/semantic_segmentation.py"
OUT = os.path.join(ROOT, "tests", "golden")


def reference_tree():
    return ast.parse(open(SRC).read())


def extract(tree, functions=(), classes=(), assigns=(), keep_methods=None):
    """A module holding only the named top-level functions / classes / assignments of the reference source.
    Decorators are dropped (`@tf.autograph.experimental.do_not_convert`, ss.py:437); `keep_methods` prunes a class
    body to the listed methods plus its plain constants."""
    body = []
    for n in tree.body:
        if isinstance(n, ast.FunctionDef) and n.name in functions:
            n.decorator_list = []
            body.append(n)
        elif isinstance(n, ast.ClassDef) and n.name in classes:
            if keep_methods is not None:
                n.body = [b for b in n.body if (isinstance(b, ast.FunctionDef) and b.name in keep_methods)
                          or isinstance(b, (ast.Assign, ast.Expr))]
            body.append(n)
        elif isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id in assigns for t in n.targets):
            body.append(n)
    return ast.Module(body=body, type_ignores=[])


# ------------------------------------------------------------------------------------------------ numpy shims
class _NpK:
    """tensorflow.keras.backend, the three calls the pinned functions make."""
    log = staticmethod(np.log)
    mean = staticmethod(np.mean)

    @staticmethod
    def argmax(x, axis=-1):
        return _T(np.argmax(np.asarray(x), axis=axis))      # first maximum wins, like tf.argmax


class _Shape(tuple):
    @property
    def ndims(self):
        return len(self)


class _T(np.ndarray):
    """numpy array with TensorShape-like `.shape.ndims`."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))


class _Var:
    def __init__(self, v):
        self.v = v

    def assign_add(self, d):
        self.v = self.v + np.asarray(d)
        return self.v

    def assign(self, d):
        self.v = np.asarray(d).copy()
        return self.v


class _MeanIoUBase:
    """Stand-in for tf.keras.metrics.MeanIoU.__init__: num_classes, the metric dtype, a zero total_cm."""

    def __init__(self, num_classes, name=None, dtype=None):
        self.num_classes, self.name, self._dtype = num_classes, name, dtype or np.float32
        self.total_cm = _Var(np.zeros((num_classes, num_classes), dtype=np.float64))


def _confusion_matrix(labels, predictions, num_classes, weights=None, dtype=np.float64):
    """tf.math.confusion_matrix: cm[label, prediction] += weight (1 when weights is None)."""
    t = np.asarray(labels).astype(np.int64).reshape(-1)
    p = np.asarray(predictions).astype(np.int64).reshape(-1)
    w = np.ones(t.shape, dtype=dtype) if weights is None else np.asarray(weights, dtype=dtype).reshape(-1)
    cm = np.zeros((num_classes, num_classes), dtype=dtype)
    np.add.at(cm, (t, p), w)
    return cm


def reference_functions():
    tree = reference_tree()
    ns = {"K": _NpK, "np": np}
    exec(compile(extract(tree, functions=("class_balanced_loss", "class_imbalance_loss"), assigns=("ss_pw", "ss_nw")),
                 SRC, "exec"), ns)
    ns2 = {"K": _NpK, "np": np, "MeanIoU": _MeanIoUBase,
           "math_ops": types.SimpleNamespace(cast=lambda x, dt: _T(np.asarray(x).astype(dt))),
           "array_ops": types.SimpleNamespace(reshape=lambda x, s: _T(np.asarray(x).reshape(s))),
           "confusion_matrix": types.SimpleNamespace(confusion_matrix=_confusion_matrix),
           "dtypes": types.SimpleNamespace(float64=np.float64)}
    exec(compile(extract(reference_tree(), classes=("MeanIoUExt",)), SRC, "exec"), ns2)
    return ns, ns2["MeanIoUExt"]


# ------------------------------------------------------------------------------------------------ 1, 2, 4
def make_loss_and_miou():
    ns, MeanIoUExt = reference_functions()
    cbl, legacy = ns["class_balanced_loss"], ns["class_imbalance_loss"]
    pw, nw = ns["ss_pw"], ns["ss_nw"]
    np.savez_compressed(os.path.join(OUT, "loss_voc_weights.npz"), ss_pw=np.asarray(pw), ss_nw=np.asarray(nw))
    rng = np.random.default_rng(1024)

    def probs(shape, sharp):
        z = rng.normal(size=shape) * sharp
        e = np.exp(z - z.max(-1, keepdims=True))
        return e / e.sum(-1, keepdims=True)

    def onehot(lab, C):
        return np.eye(C, dtype=np.float64)[lab]

    cases = {}
    # VOC weights, 21 classes, fp64 inputs of varied sharpness (sharp = 12: probabilities underflow towards 0 / 1, the
    # epsilon inside both logs decides the value)
    for name, shape, sharp in (("voc_soft", (2, 9, 11, 21), 1.0), ("voc_sharp", (1, 16, 16, 21), 12.0),
                               ("voc_flat", (3, 5, 7, 21), 0.0)):
        p = probs(shape, sharp)
        y = onehot(rng.integers(0, 21, shape[:-1]), 21)
        cases[name] = (y, p, list(pw), list(nw), 1e-7)
    # exact 0 / 1 probabilities (log(eps), log(1 + eps)), another class count / epsilon, soft (non one-hot) truth
    p = onehot(rng.integers(0, 5, (2, 6, 6)), 5)
    y = onehot(rng.integers(0, 5, (2, 6, 6)), 5)
    f = rng.dirichlet(np.ones(5))
    cases["hard_probs_5cls"] = (y, p, list(1 - f), list(f), 1e-7)
    cases["soft_truth_eps1e-3"] = (probs((2, 4, 4, 7), 1.0), probs((2, 4, 4, 7), 2.0), list(rng.uniform(0.2, 1, 7)),
                                   list(rng.uniform(0, 0.3, 7)), 1e-3)
    for name, (y, p, a, b, eps) in cases.items():
        v = float(cbl(y, p, a, b, eps))
        v2 = float(legacy(a, b, eps)(y, p))
        assert abs(v - v2) <= 1e-12 * max(1.0, abs(v)), "the two loss definitions of the reference disagree"
        # float32 evaluation as the reference runs it (hps.dtype float32): the same source on float32 arrays
        v32 = float(cbl(y.astype(np.float32), p.astype(np.float32), [np.float32(t) for t in a],
                        [np.float32(t) for t in b], np.float32(eps)))
        np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), y_true=y, y_pred=p, pos_weights=np.asarray(a),
                            neg_weights=np.asarray(b), epsilon=eps, loss=v, loss_f32=v32)
        print("loss", name, v, v32)

    # MeanIoUExt.update_state: accumulate over three batches / overwrite; ties in the prediction (first max wins)
    for name, C, accum in (("accumulate_21", 21, True), ("overwrite_21", 21, False), ("accumulate_4_ties", 4, True)):
        m = MeanIoUExt(C, accum_enable=accum)
        batches, cms = [], []
        for b in range(3):
            shape = (2, 8, 8, C)
            yt = onehot(rng.integers(0, C, shape[:-1]), C)
            yp = probs(shape, 2.0)
            if "ties" in name:
                yp = np.round(yp * 4) / 4               # many exact ties
            cm = np.asarray(m.update_state(yt, yp))
            batches.append((yt, yp))
            cms.append(cm.copy())
        np.savez_compressed(os.path.join(OUT, f"miou_{name}.npz"), num_classes=C, accum_enable=accum,
                            y_true=np.stack([b[0] for b in batches]), y_pred=np.stack([b[1] for b in batches]),
                            total_cm=np.stack(cms))
        print("miou", name, cms[-1].sum())


# ------------------------------------------------------------------------------------------------ 3
def reference_builder():
    """The reference's SemanticSegmentation class (constructor + the three builder methods) bound to this repo's
    keras mirror."""
    from deeplabv3plus_keras_b200 import deeplab, keras
    from deeplabv3plus_keras_b200.keras import applications
    tree = reference_tree()
    const = [n for n in tree.body if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name)
             and n.targets[0].id.startswith("BASE_MODEL_")]
    mod = extract(tree, classes=("SemanticSegmentation",),
                  keep_methods=("__init__", "_make_encoder", "_make_decoder", "_refine_boundary"))
    mod.body = const + mod.body
    ns = {"os": os, "Input": keras.Input, "Conv2D": keras.Conv2D, "SeparableConv2D": keras.SeparableConv2D,
          "BatchNormalization": keras.BatchNormalization, "Activation": keras.Activation,
          "AveragePooling2D": keras.AveragePooling2D, "Lambda": keras.Lambda, "Concatenate": keras.Concatenate,
          "Dropout": keras.Dropout, "Model": keras.Model, "K": keras.backend, "regularizers": keras.regularizers,
          "initializers": keras.initializers, "optimizers": types.SimpleNamespace(Adam=deeplab.Adam),
          # the reference lets keras.applications default to weights='imagenet' (a download); random init here
          "Xception": functools.partial(applications.Xception, weights=None),
          "MobileNetV2": functools.partial(applications.MobileNetV2, weights=None),
          "ClassBalancedLoss": deeplab.ClassBalancedLoss, "MeanIoUExt": deeplab.MeanIoUExt,
          "ss_pw": deeplab.ss_pw, "ss_nw": deeplab.ss_nw}
    exec(compile(mod, SRC, "exec"), ns)
    return ns["SemanticSegmentation"]


def reference_graph(case_kwargs):
    from deeplabv3plus_keras_b200 import keras
    from tests import topology, util
    conf = util.make_conf(**case_kwargs)
    keras.reset_uids()
    keras.set_random_seed(1024)
    ref = reference_builder()(conf)
    return topology.describe(ref.model), conf


def make_topology():
    from tests import topology
    for name, kw in topology.TOPOLOGY_CASES.items():
        desc, _ = reference_graph(kw)
        with open(os.path.join(OUT, f"topology_{name}.json"), "w") as f:
            json.dump(desc, f, indent=0, sort_keys=True)
        print("topology", name, {k: len(v) for k, v in desc.items()})


if __name__ == "__main__":
    make_loss_and_miou()
    make_topology()

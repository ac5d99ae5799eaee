"""Regenerate tests/golden/*.npz from the REAL reference (tonandr/deeplabv3plus_keras) on a machine that has
TensorFlow 2.4 — the out-of-band pin of the oracle.  Not runnable in the build image (no TensorFlow).

    python scripts/dump_reference_tf.py /path/to/deeplabv3plus_keras  [--out tests/golden_tf]

For each fixture in tests/golden it
  1. builds the reference `SemanticSegmentation(conf)` (bodhi/deeplabv3plus_keras/semantic_segmentation.py:450) with
     the fixture's conf — `keras.applications.{Xception,MobileNetV2}` are wrapped so that `weights=None` (the reference
     hard-codes the ImageNet download, ss.py:496-499 / 512-515) and optional imports the hot path never touches
     (cupy, skimage, matplotlib) are stubbed;
  2. loads the fixture's seeded weights BY LAYER NAME (the product and the oracle use tf.keras' automatic names, so
     `model.get_layer(name).set_weights(...)` works directly, in get_weights() order);
  3. runs the training-mode forward, the reference's own `ClassBalancedLoss` (+ the Keras L2 regularisers) and a
     GradientTape backward on the fixture's inputs / one-hot labels;
  4. writes logits (pre-resize conv output), loss, L2 term and the same sample of gradients to <out>/<name>.npz and
     prints the deviation from the committed fixture.
Agreement within 1e-5 (fp32 TF vs the fp64 oracle) pins oracle/tf_ops.py and oracle/model.py against TensorFlow.
"""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("reference_root")
ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden_tf"))
args = ap.parse_args()

for name in ("cupy", "cupyx", "cupyx.scipy", "cupyx.scipy.ndimage", "skimage", "skimage.io", "matplotlib",
             "matplotlib.pyplot"):
    try:
        __import__(name)
    except Exception:                       # hot path never calls into these
        sys.modules[name] = types.ModuleType(name)

import tensorflow as tf  # noqa: E402

assert tf.__version__.startswith("2."), tf.__version__
for app in ("Xception", "MobileNetV2"):
    orig = getattr(tf.keras.applications, app)

    def wrapped(*a, _orig=orig, **kw):
        kw["weights"] = None                # no ImageNet download; weights are injected below
        return _orig(*a, **kw)
    setattr(tf.keras.applications, app, wrapped)

sys.path.insert(0, args.reference_root)
from bodhi.deeplabv3plus_keras import semantic_segmentation as ref  # noqa: E402

from tests import util  # noqa: E402  (weight generator shared with the fixtures)

os.makedirs(args.out, exist_ok=True)
GOLDEN = os.path.join(ROOT, "tests", "golden")
for fname in sorted(os.listdir(GOLDEN)):
    if not fname.endswith(".npz"):
        continue
    g = np.load(os.path.join(GOLDEN, fname), allow_pickle=True)
    conf = g["conf"].item()
    conf = dict(conf, base_model=conf["base_model"], model_loading=False)
    tf.keras.backend.clear_session()
    tf.random.set_seed(1024)
    ss = ref.SemanticSegmentation(conf)

    # the product's layer graph gives the seeded weights under tf.keras' layer names
    mine = util.build(g["conf"].item())
    util.randomize_weights(mine.model, seed=int(g["weight_seed"]))
    by_layer = {}
    for layer in mine.model.flat_layers():
        if layer._weights:
            by_layer[layer.name] = layer.get_weights()

    def all_layers(m):
        for l in m.layers:
            if isinstance(l, tf.keras.Model):
                yield from all_layers(l)
            else:
                yield l
    seen = set()
    for l in all_layers(ss.model):
        if l.weights and l.name not in seen:
            l.set_weights(by_layer[l.name])
            seen.add(l.name)
    assert seen == set(by_layer), (set(by_layer) - seen, seen - set(by_layer))

    x = tf.constant(g["x"], tf.float32)
    C = conf["nn_arch"]["num_classes"]
    y = tf.one_hot(g["y"], C, dtype=tf.float32)
    loss_fn = ref.ClassBalancedLoss(list(g["pw"]), list(g["nw"]))
    logits_layer = [l for l in all_layers(ss.model) if isinstance(l, tf.keras.layers.Conv2D) and l.filters == C][-1]
    probe = tf.keras.Model(ss.decoder.inputs, logits_layer.output)
    with tf.GradientTape() as tape:
        probs = ss.model(x, training=True)
        data = loss_fn(y, probs)
        l2 = tf.add_n(ss.model.losses) if ss.model.losses else tf.constant(0.0)
        total = data + l2
    grads = tape.gradient(total, ss.model.trainable_variables)
    gmap = {v.name.split(":")[0]: gr.numpy() for v, gr in zip(ss.model.trainable_variables, grads)}
    feats = ss.encoder(x, training=True)
    logits = probe([x, feats] if conf["nn_arch"]["boundary_refinement"] else feats, training=True).numpy()
    out = {"logits": logits, "loss": float(data), "l2": float(l2)}
    for k in g["grad_keys"]:
        k = str(k)
        out["grad/" + k] = gmap[k]
    np.savez_compressed(os.path.join(args.out, fname), **out)
    dev = np.abs(logits - g["logits"]).max() / np.abs(g["logits"]).max()
    print(f"{fname}: logits rel dev {dev:.3e}  loss {float(data):.8f} vs {float(g['loss']):.8f}  "
          f"l2 {float(l2):.3e} vs {float(g['l2']):.3e}")
    for k in g["grad_keys"]:
        k = str(k)
        a, b = gmap[k], g["grad/" + k]
        print(f"   grad {k}: rms-rel {np.sqrt(((a - b) ** 2).mean()) / max(np.sqrt((b ** 2).mean()), 1e-30):.3e}")

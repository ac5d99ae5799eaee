#!/usr/bin/env python
"""Convert Keras weights to the `.npz` interchange format of this package (arrays keyed "layer_name/weight_name" in
Keras `get_weights()` layouts: conv HWIO, depthwise [3,3,C,1], pointwise [1,1,Cin,Cout], BatchNormalization gamma /
beta / moving_mean / moving_variance) — SURVEY.md §8f rank 4.

Run it on a machine that has the source weights (this build container has neither TensorFlow nor h5py nor a network):

  # ImageNet-pretrained backbones the reference starts from (ss.py:496-499, 512-515)
  python scripts/convert_keras_weights.py --model xception    --out ~/.keras/models/xception_imagenet_notop.npz
  python scripts/convert_keras_weights.py --model mobilenetv2 --out ~/.keras/models/mobilenet_v2_1.0_imagenet_notop.npz
  # a Keras-Applications .h5 file already on disk (needs h5py only)
  python scripts/convert_keras_weights.py --h5 xception_weights_tf_dim_ordering_tf_kernels_notop.h5 --out x.npz
  # a model checkpointed by the reference itself (SavedModel directory or .h5, ss.py:983-986): backbone + head
  python scripts/convert_keras_weights.py --saved-model resource/semantic_segmentation_deeplabv3plus --out ckpt.npz

Then `Xception(weights='imagenet')` (i.e. the reference's default call) finds the file under $DLV3P_PRETRAINED_DIR or
~/.keras/models, and `utils.load_weights_npz(model, 'ckpt.npz')` / conf['model_loading'] resumes a full model.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def short(weight_name: str) -> str:
    """'block1_conv1/kernel:0' -> 'kernel'; 'bn/moving_mean:0' -> 'moving_mean'."""
    return weight_name.split("/")[-1].split(":")[0]


def arrays_from_layers(layers) -> dict:
    """{'layer/weight': array} from objects with `.name`, `.weights` (each with `.name`) and `.get_weights()` —
    tf.keras layers, or any stand-in with the same attributes.  Nested models are expanded."""
    out = {}
    for layer in layers:
        if hasattr(layer, "layers"):
            out.update(arrays_from_layers(layer.layers))
            continue
        values = layer.get_weights()
        for w, v in zip(layer.weights, values):
            out[f"{layer.name}/{short(w.name)}"] = np.asarray(v, dtype=np.float32)
    return out


def arrays_from_h5(path: str) -> dict:
    """Keras `save_weights` HDF5 layout: root attr `layer_names`, per-layer group attr `weight_names`."""
    import h5py
    out = {}
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        for lname in g.attrs["layer_names"]:
            lname = lname.decode() if isinstance(lname, bytes) else lname
            for wname in g[lname].attrs["weight_names"]:
                wname = wname.decode() if isinstance(wname, bytes) else wname
                out[f"{lname}/{short(wname)}"] = np.asarray(g[lname][wname], dtype=np.float32)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--model", choices=["xception", "mobilenetv2"], help="tf.keras.applications model, ImageNet, no top")
    src.add_argument("--h5", help="Keras .h5 weights file")
    src.add_argument("--saved-model", help="SavedModel directory / .h5 written by the reference's ModelCheckpoint")
    ap.add_argument("--out", required=True)
    args = ap.parse_args(argv)
    if args.h5:
        arrays = arrays_from_h5(args.h5)
    else:
        import tensorflow as tf
        if args.model:
            ctor = {"xception": tf.keras.applications.Xception, "mobilenetv2": tf.keras.applications.MobileNetV2}[args.model]
            model = ctor(include_top=False, weights="imagenet", input_shape=(513, 513, 3))
        else:
            model = tf.keras.models.load_model(args.saved_model, compile=False)
        arrays = arrays_from_layers(model.layers)
    np.savez_compressed(args.out, **arrays)
    print(f"wrote {len(arrays)} arrays ({sum(a.size for a in arrays.values())} parameters) to {args.out}")


if __name__ == "__main__":
    sys.exit(main())

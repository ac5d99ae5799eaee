"""Weight interchange: `.npz` files keyed "layer_name/weight_name" with arrays in Keras
`get_weights()` layouts (depthwise [3,3,C,1], pointwise [1,1,Cin,Cout], conv HWIO, BatchNormalization gamma / beta /
moving_mean / moving_variance).  Layers carry the names tf.keras gives them, so a file written on a TensorFlow
machine with `np.savez(path, **{f"{l.name}/{w.name.split('/')[-1].split(':')[0]}": w.numpy() ...})` loads here and
vice versa.  The reference checkpoints a SavedModel directory (`ModelCheckpoint`, ss.py:983-986) and resumes with
`load_model` when `model_loading` is true (ss.py:482-485); here `model_loading` reads
`<resource_path>/semantic_segmentation_deeplabv3plus.npz`."""
from __future__ import annotations

from typing import Dict

import numpy as np


def save_weights_npz(model, path: str, include_optimizer: bool = True) -> None:
    """Layer weights (read back from the device if an optimizer step ran since) and, like the reference's SavedModel
    checkpoints (ss.py:983-986), the optimizer state under 'optimizer/': Adam moments, iteration count, learning rate
    and the dropout stream position."""
    arrays = {k: np.asarray(v) for k, v in model.named_weights().items()}
    if include_optimizer:
        arrays.update(model.optimizer_state())
    np.savez_compressed(path, **arrays)


def load_weights_npz(model, path: str, strict: bool = True) -> Dict[str, str]:
    """Copies arrays into the model's layers by name; returns {name: "loaded" | "missing"}.  With `strict`, a missing
    or unexpected name or a shape mismatch raises ValueError (Keras `load_weights` behaviour)."""
    data = np.load(path if path.endswith(".npz") else path + ".npz")
    named = model.named_weights()
    report = {}
    opt_keys = [k for k in data.files if k.startswith("optimizer/")]
    extra = set(data.files) - set(named) - set(opt_keys)
    if strict and extra:
        raise ValueError(f"{path}: unexpected weights {sorted(extra)[:5]} ...")
    for k, w in named.items():
        if k not in data.files:
            if strict:
                raise ValueError(f"{path}: weight {k!r} is missing")
            report[k] = "missing"
            continue
        v = data[k]
        if tuple(v.shape) != tuple(w.shape):
            raise ValueError(f"{path}: {k!r} has shape {tuple(v.shape)}, the layer expects {tuple(w.shape)}")
        w[...] = v.astype(w.dtype)
        report[k] = "loaded"
    model._invalidate()                                      # device copies re-upload before their next use
    if opt_keys:
        model.set_optimizer_state({k: data[k] for k in opt_keys})
    return report

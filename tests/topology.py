"""Canonical description of a functional model graph (layer types, names, constructor arguments, connectivity,
output shapes, weight shapes) — used to prove that deeplab.py builds the SAME graph as the reference's own
`_make_encoder` / `_make_decoder` / `_refine_boundary` source executed against this repo's keras mirror
(scripts/make_golden_reference_fns.py writes tests/golden/topology_*.json from the reference source)."""
from __future__ import annotations

from typing import Dict, List

from deeplabv3plus_keras_b200.keras.base import InputLayer
from deeplabv3plus_keras_b200.keras.models import Model

_SKIP = {"built", "function", "layers", "nodes", "inputs", "outputs", "optimizer", "loss", "metrics"}


def _plain(v):
    if isinstance(v, (bool, int, float, str)) or v is None:
        return v
    if isinstance(v, (tuple, list)):
        return [_plain(x) for x in v]
    if hasattr(v, "l2"):
        return {"l2": float(v.l2)}
    return type(v).__name__                       # initializer objects: the class is what matters


def layer_config(layer) -> Dict:
    return {k: _plain(v) for k, v in sorted(vars(layer).items()) if not k.startswith("_") and k not in _SKIP}


def describe(model: Model) -> Dict[str, List[Dict]]:
    """{model name: [one row per node, topological order]} for the model and every nested model."""
    out: Dict[str, List[Dict]] = {}

    def rec(m: Model):
        if m.name in out:
            return
        rows = []
        out[m.name] = rows
        for node in m.nodes:
            lay = node.layer
            row = {"layer": lay.name, "type": type(lay).__name__,
                   "inputs": [t.node.layer.name for t in node.inputs],
                   "output_shape": [list(t.shape) for t in node.outputs],
                   "dtype": node.outputs[0].dtype}
            if isinstance(lay, Model):
                rec(lay)
            else:
                row["config"] = layer_config(lay)
                row["weights"] = {n: list(lay._weights[n].shape) for n in lay.weight_names()}
                row["trainable_weights"] = [n for n in lay.weight_names() if lay._trainable[n]]
                if isinstance(lay, InputLayer):
                    row["config"]["shape"] = list(lay.shape)
            rows.append(row)
        rows.append({"model_inputs": [t.node.layer.name for t in m.inputs],
                     "model_outputs": [t.node.layer.name for t in m.outputs]})
    rec(model)
    return out


TOPOLOGY_CASES = {
    "xception_os16": dict(base="xception", output_stride=16, image_size=513),
    "xception_os8_br": dict(base="xception", output_stride=8, image_size=513, refine=True, rate_mult=2),
    "mobilenetv2_os16_default_aspp": dict(base="mobilenetv2", output_stride=16, image_size=513, aspp="default"),
    "mobilenetv2_os8_br": dict(base="mobilenetv2", output_stride=8, image_size=224, refine=True, aspp="default"),
    "xception_os16_global_pool": dict(base="xception", output_stride=16, image_size=513, aspp="global_pool"),
}

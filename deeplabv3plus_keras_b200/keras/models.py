"""tf.keras.Model subset: functional construction, nesting (a Model is a Layer that can be called on new
symbolic tensors — ss.py:802,930,778-780), get_layer / layers / weights, compile / predict / train_on_batch.
Execution is delegated to engine.Plan (built lazily per batch size), never to a CPU path."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .base import InputLayer, KTensor, Layer, Node


def _toposort(outputs: Sequence[KTensor]) -> List[Node]:
    order: List[Node] = []
    seen = set()

    def visit(t: KTensor):
        node = t.node
        if id(node) in seen:
            return
        seen.add(id(node))
        for i in node.inputs:
            visit(i)
        order.append(node)

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        for t in outputs:
            visit(t)
    finally:
        sys.setrecursionlimit(old)
    return order


class Model(Layer):
    _default_prefix = "model"

    def __init__(self, inputs, outputs, name=None):
        super().__init__(name=name)
        self.inputs: List[KTensor] = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs: List[KTensor] = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self.nodes: List[Node] = _toposort(self.outputs)
        input_ids = {id(t) for t in self.inputs}
        for n in self.nodes:
            if isinstance(n.layer, InputLayer) and id(n.outputs[0]) not in input_ids:
                raise ValueError(f"graph disconnected: Input '{n.layer.name}' is not among the model inputs")
        self.layers: List[Layer] = []
        for n in self.nodes:
            if n.layer not in self.layers:
                self.layers.append(n.layer)
        self.built = True
        self.optimizer = None
        self.loss = None
        self.metrics = []
        self._plans: Dict = {}

    # -- Layer protocol: a nested model is one node of the outer graph ------------------------------------
    def compute_output_shape(self, shapes):
        return self.outputs[0].shape

    def output_dtype(self, inputs):
        return self.outputs[0].dtype

    def __call__(self, inputs):
        ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        if len(ins) != len(self.inputs):
            raise ValueError(f"model {self.name} expects {len(self.inputs)} inputs, got {len(ins)}")
        for a, b in zip(ins, self.inputs):
            if tuple(a.shape[1:]) != tuple(b.shape[1:]):
                raise ValueError(f"model {self.name}: input shape {a.shape} incompatible with {b.shape}")
        node = Node(self, ins)
        node.outputs = [KTensor(o.shape, o.dtype, node, i, f"{self.name}/out{len(self._nodes)}_{i}")
                        for i, o in enumerate(self.outputs)]
        self._nodes.append(node)
        return node.outputs[0] if len(node.outputs) == 1 else node.outputs

    # -- introspection -------------------------------------------------------------------------------------
    def get_layer(self, name=None, index=None) -> Layer:
        if index is not None:
            return self.layers[index]
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def flat_layers(self) -> List[Layer]:
        """All weight-bearing leaf layers, nested models expanded, each once, in graph order."""
        out: List[Layer] = []

        def rec(m: "Model"):
            for l in m.layers:
                if isinstance(l, Model):
                    rec(l)
                elif l not in out:
                    out.append(l)
        rec(self)
        return out

    def get_weights(self):
        return [w for l in self.flat_layers() for w in l.get_weights()]

    def set_weights(self, values):
        values = list(values)
        i = 0
        for l in self.flat_layers():
            n = len(l.weight_names())
            l.set_weights(values[i:i + n])
            i += n
        if i != len(values):
            raise ValueError(f"set_weights: {len(values)} arrays given, model has {i}")
        self._invalidate()

    def named_weights(self) -> Dict[str, np.ndarray]:
        """{'layer/weight': array} — the exchange format of utils.save_weights_npz."""
        return {f"{l.name}/{n}": l._weights[n] for l in self.flat_layers() for n in l.weight_names()}

    def count_params(self) -> int:
        return int(sum(l.count_params() for l in self.flat_layers()))

    def _invalidate(self):
        for p in self._plans.values():
            p.upload_weights()

    # -- execution -----------------------------------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, metrics=None):
        self.optimizer, self.loss, self.metrics = optimizer, loss, list(metrics or [])

    def plan(self, batch_size: int, training: bool = False, **kw):
        from ..engine import Plan
        key = (int(batch_size), bool(training), tuple(sorted(kw.items())))
        if key not in self._plans:
            self._plans[key] = Plan(self, int(batch_size), training=training, **kw)
        return self._plans[key]

    def predict(self, x, batch_size: Optional[int] = None):
        """Model.predict (ss.py:1084,1172,1225): host array in, softmax probabilities [B,H,W,C] out (host)."""
        x = np.asarray(x)
        bs = int(batch_size or min(len(x), 32))
        outs = []
        for i in range(0, len(x), bs):
            chunk = x[i:i + bs]
            p = self.plan(len(chunk), training=False)
            outs.append(p.predict(chunk))
        return np.concatenate(outs, axis=0)

    def train_on_batch(self, x, y):
        """One optimizer step on host arrays; y is one-hot [B,H,W,C] (the reference's Sequence output, ss.py:1602)
        or an integer label map [B,H,W]."""
        p = self.plan(len(x), training=True)
        return p.train_on_batch(np.asarray(x), np.asarray(y))

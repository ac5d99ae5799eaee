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
        self._stores_by_device: Dict = {}        # device string -> engine.ParamStore shared by all plans
        self._pending_opt_state: Optional[Dict] = None

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

    def _sync_host(self):
        for s in self._stores_by_device.values():
            s.sync_host()
        for l in self.flat_layers():
            l._sync_host()

    def get_weights(self):
        self._sync_host()
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
        """{'layer/weight': array} — the exchange format of utils.save_weights_npz.  The arrays are the layers' host
        master copies (brought up to date with the device first); after writing into them in place call
        `_invalidate()`."""
        self._sync_host()
        return {f"{l.name}/{n}": l._weights[n] for l in self.flat_layers() for n in l.weight_names()}

    def count_params(self) -> int:
        return int(sum(l.count_params() for l in self.flat_layers()))

    def _invalidate(self):
        """The host arrays were modified: every device store re-uploads before its next use."""
        stores = list(self._stores_by_device.values())
        for l in self.flat_layers():
            stores += [s for s in l._stores if s not in stores]
        for s in stores:
            s.dev_stale = True

    # -- optimizer state (checkpoint / resume, ss.py:482-485: load_model restores the optimizer) ------------
    def optimizer_state(self) -> Dict[str, np.ndarray]:
        """{'optimizer/iterations', 'optimizer/lr', 'optimizer/step_counter', 'optimizer/m/<layer>/<weight>',
        'optimizer/v/...'}: what Adam needs to continue (moments, the iteration count that drives the bias correction
        and the inverse-time decay) plus the position of the dropout stream."""
        out: Dict[str, np.ndarray] = {}
        if self.optimizer is not None:
            out["optimizer/iterations"] = np.asarray(self.optimizer.iterations, dtype=np.int64)
            out["optimizer/lr"] = np.asarray(self.optimizer.lr, dtype=np.float64)
        for store in self._stores_by_device.values():
            out["optimizer/step_counter"] = store.step_counter.cpu().numpy().astype(np.int64).reshape(())
            for l, n in store.trainable_items():
                out[f"optimizer/m/{l.name}/{n}"] = store.logical(l, n, arena="m").cpu().numpy().copy()
                out[f"optimizer/v/{l.name}/{n}"] = store.logical(l, n, arena="v").cpu().numpy().copy()
            break
        return out

    def set_optimizer_state(self, state: Dict[str, np.ndarray]) -> None:
        """Inverse of optimizer_state(); parts that have no home yet (no optimizer compiled, no device store
        created) are kept and applied by compile() / the first plan."""
        self._pending_opt_state = dict(state)
        self._apply_pending_optimizer_state()

    def _apply_pending_optimizer_state(self):
        st = self._pending_opt_state
        if not st:
            return
        import torch
        if self.optimizer is not None:
            if "optimizer/iterations" in st:
                self.optimizer.iterations = int(st.pop("optimizer/iterations"))
            if "optimizer/lr" in st:
                self.optimizer.lr = float(st.pop("optimizer/lr"))
        if self._stores_by_device:
            moments = [k for k in st if k.startswith(("optimizer/m/", "optimizer/v/"))]
            for store in self._stores_by_device.values():
                if "optimizer/step_counter" in st:
                    store.step_counter.fill_(int(st["optimizer/step_counter"]))
                by_name = {f"{l.name}/{n}": (l, n) for l, n in store.trainable_items()}
                for k in moments:
                    arena, name = k.split("/", 2)[1], k.split("/", 2)[2]
                    if name not in by_name:
                        raise ValueError(f"optimizer state for unknown weight {name!r}")
                    l, n = by_name[name]
                    dst = store.logical(l, n, arena=arena)
                    v = np.asarray(st[k], dtype=np.float32)
                    if tuple(v.shape) != tuple(dst.shape):
                        raise ValueError(f"{k}: shape {tuple(v.shape)}, expected {tuple(dst.shape)}")
                    dst.copy_(torch.from_numpy(v))
            st.pop("optimizer/step_counter", None)
            for k in moments:
                st.pop(k)
        if not st:
            self._pending_opt_state = None

    # -- execution -----------------------------------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, metrics=None):
        self.optimizer, self.loss, self.metrics = optimizer, loss, list(metrics or [])
        self._apply_pending_optimizer_state()

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

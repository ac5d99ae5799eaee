"""Symbolic layer graph: the minimum of the tf.keras functional API that the reference's model-building code
touches (ss.py:52-57,79 imports; call sites ss.py:795-954).  Layers here only record topology, shapes and fp32
master weights; arithmetic happens when an engine.Plan lowers the graph to libdlv3p kernel launches.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_UIDS: Dict[str, int] = {}
_RNG = np.random.default_rng(1024)   # the reference seeds everything with 1024 (ss.py:1797-1802)


def set_random_seed(seed: int) -> None:
    global _RNG
    _RNG = np.random.default_rng(seed)


def rng() -> np.random.Generator:
    return _RNG


def reset_uids() -> None:
    """K.clear_session(): restart the automatic layer-name counters (conv2d, conv2d_1, ...)."""
    _UIDS.clear()


def unique_name(prefix: str) -> str:
    n = _UIDS.get(prefix, 0)
    _UIDS[prefix] = n + 1
    return prefix if n == 0 else f"{prefix}_{n}"


class KTensor:
    """Symbolic tensor: static shape (batch axis None), dtype string, producing node."""

    def __init__(self, shape: Tuple, dtype: str, node: Optional["Node"], index: int = 0, name: str = ""):
        self.shape = tuple(shape)
        self.dtype = dtype
        self.node = node
        self.index = index
        self.name = name

    def __repr__(self):
        return f"<KTensor {self.name} shape={self.shape} dtype={self.dtype}>"


class Node:
    def __init__(self, layer: "Layer", inputs: List[KTensor]):
        self.layer = layer
        self.inputs = inputs
        self.outputs: List[KTensor] = []


class Layer:
    _default_prefix: Optional[str] = None

    def __init__(self, name: Optional[str] = None, trainable: bool = True, dtype: Optional[str] = None, **kwargs):
        if kwargs:
            raise TypeError(f"{type(self).__name__}: unsupported keyword arguments {sorted(kwargs)}")
        prefix = self._default_prefix or _snake(type(self).__name__)
        self.name = name if name is not None else unique_name(prefix)
        self.trainable = trainable
        self.dtype = dtype
        self.built = False
        self._weights: "OrderedDict[str, np.ndarray]" = OrderedDict()
        self._trainable: Dict[str, bool] = {}
        self._nodes: List[Node] = []
        self._stores: list = []        # engine.ParamStore objects holding this layer's weights on a device

    # -- weights -------------------------------------------------------------------------------------
    def add_weight(self, name: str, shape, initializer, trainable: bool = True) -> np.ndarray:
        w = np.asarray(initializer(tuple(shape)), dtype=np.float32)
        self._weights[name] = w
        self._trainable[name] = trainable
        return w

    def _sync_host(self) -> None:
        """Device -> host, when an optimizer step has run since the host arrays were last current."""
        for s in self._stores:
            s.sync_host()

    @property
    def weights(self) -> List[np.ndarray]:
        self._sync_host()
        return list(self._weights.values())

    def get_weights(self) -> List[np.ndarray]:
        """Keras order: trainable weights first (in creation order), then non-trainable."""
        self._sync_host()
        tr = [w for n, w in self._weights.items() if self._trainable[n]]
        nt = [w for n, w in self._weights.items() if not self._trainable[n]]
        return [w.copy() for w in tr + nt]

    def weight_names(self) -> List[str]:
        return [n for n in self._weights if self._trainable[n]] + [n for n in self._weights if not self._trainable[n]]

    def set_weights(self, values: Sequence[np.ndarray]) -> None:
        names = self.weight_names()
        if len(values) != len(names):
            raise ValueError(f"layer {self.name}: expected {len(names)} weight arrays, got {len(values)}")
        self._sync_host()              # the other layers' trained values must survive the re-upload
        for n, v in zip(names, values):
            v = np.asarray(v, dtype=np.float32)
            if v.shape != self._weights[n].shape:
                raise ValueError(f"layer {self.name}/{n}: shape {v.shape} != {self._weights[n].shape}")
            self._weights[n][...] = v
        for s in self._stores:
            s.dev_stale = True         # uploaded lazily by the next plan that runs (Plan.ensure_current)

    def count_params(self) -> int:
        return int(sum(w.size for w in self._weights.values()))

    # -- graph ---------------------------------------------------------------------------------------
    def build(self, input_shapes: List[Tuple]) -> None:
        pass

    def compute_output_shape(self, input_shapes: List[Tuple]) -> Tuple:
        return input_shapes[0]

    def output_dtype(self, inputs: List[KTensor]) -> str:
        return inputs[0].dtype

    def __call__(self, inputs):
        ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        for t in ins:
            if not isinstance(t, KTensor):
                raise TypeError(f"layer {self.name} called on {type(t).__name__}; expected symbolic tensors")
        shapes = [t.shape for t in ins]
        if not self.built:
            self.build(shapes)
            self.built = True
        out_shape = self.compute_output_shape(shapes)
        node = Node(self, ins)
        node.outputs = [KTensor(out_shape, self.output_dtype(ins), node, 0, f"{self.name}/out{len(self._nodes)}")]
        self._nodes.append(node)
        return node.outputs[0]

    @property
    def output(self) -> KTensor:
        if not self._nodes:
            raise AttributeError(f"layer {self.name} has never been called")
        return self._nodes[0].outputs[0]

    @property
    def input(self) -> KTensor:
        if not self._nodes:
            raise AttributeError(f"layer {self.name} has never been called")
        return self._nodes[0].inputs[0]

    def _init_set_name(self, name: str) -> None:    # used by the reference (ss.py:509,876,913)
        self.name = name


def _snake(name: str) -> str:
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and (not name[i - 1].isupper() or (i + 1 < len(name) and name[i + 1].islower())):
            out.append("_")
        out.append(ch.lower())
    s = "".join(out)
    return s.replace("conv2_d", "conv2d").replace("pooling2_d", "pooling2d").replace("padding2_d", "padding2d")


class InputLayer(Layer):
    _default_prefix = "input"

    def __init__(self, shape, dtype="float32", name=None):
        super().__init__(name=name, dtype=dtype)
        self.shape = (None,) + tuple(shape)
        node = Node(self, [])
        node.outputs = [KTensor(self.shape, dtype or "float32", node, 0, self.name)]
        self._nodes.append(node)
        self.built = True


def Input(shape, dtype=None, name=None) -> KTensor:
    """tf.keras.Input (ss.py:795-799,883,887)."""
    return InputLayer(shape, dtype or "float32", name).output

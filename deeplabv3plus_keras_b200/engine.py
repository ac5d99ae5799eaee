"""Execution engine: lowers a keras.Model layer graph to a static schedule of libdlv3p kernel launches.

No tracing compiler: the graph is flattened (nested models expanded, repeated calls of a shared sub-model on the
same tensor de-duplicated), neighbouring layers are fused into macro-ops by pattern (Conv/SeparableConv/Depthwise
-> BatchNormalization -> ReLU/ReLU6 -> Add; MaxPooling -> Add; pre-activation ReLU -> depthwise prologue), every
activation / gradient / workspace buffer is allocated once at plan time, and forward + explicit backward are lists
of closures over raw device pointers, so that a whole training step can be captured in a CUDA graph and replayed.
The backward schedule is written by hand per macro-op (there is no autograd anywhere on the product path).
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6
from .keras import layers as L
from .keras.base import InputLayer, KTensor, Layer
from .keras.models import Model

_TORCH_DT = {"float32": torch.float32, "bfloat16": torch.bfloat16}


def _ceil8(n: int) -> int:
    return (n + 7) // 8 * 8


def phys_channels(c: int) -> int:
    """Channel count as laid out in HBM.  Xception's 728-channel tensors have a 1456-byte pixel pitch: every other
    pixel row starts mid-sector, a 128-byte TMA box row then costs five 32-byte sectors over two L2 lines instead of four
    over one, and the operand feed of the 728-wide GEMMs drops from 11.5 to 8.2 TB/s (profiles/r1_gemm_feed.md).  Such
    tensors are stored with the channel axis padded to a multiple of 16 (728 -> 736); the pad channels of every
    parameter are zero, so the pad activations, their gradients and the pad parameter gradients are exact zeros and
    Adam leaves them at zero — the logical model is unchanged.  Small / already aligned counts are left alone."""
    return c if (c < 64 or c % 16 == 0) else (c + 15) // 16 * 16


_CHANNEL_AXES = {"depthwise_kernel": (2,), "pointwise_kernel": (2, 3), "kernel": (2, 3), "bias": (0,), "gamma": (0,),
                 "beta": (0,), "moving_mean": (0,), "moving_variance": (0,)}


def _phys_shape(name: str, shape, phys=phys_channels) -> Tuple[int, ...]:
    axes = _CHANNEL_AXES.get(name, ())
    return tuple(phys(d) if i in axes else d for i, d in enumerate(shape))


class _MacroScope(SimpleNamespace):
    """Names set up by Plan._emit_conv for its forward / backward halves; a name a branch did not define reads as None."""

    def __getattr__(self, name):
        return None


class Value:
    """A materialised NHWC activation, its gradient buffer and pending (zero-copy) gradient contributions."""

    # plan-time: set by the fused depthwise-backward launch that wrote this tensor's gradient LAST, so that the producer
    # (a Conv -> BN (+ residual) macro-op) can ask that launch for its BatchNormalization reductions as well
    fold_hook: Optional[dict] = None

    def __init__(self, shape, dtype, buf=None, name=""):
        self.shape = tuple(shape)
        self.dtype = dtype
        self.buf: Optional[torch.Tensor] = buf
        self.name = name
        self.needs_grad = False
        self.grad: Optional[torch.Tensor] = None
        self.grad_written = False          # plan-time state of the backward schedule
        self.pending: List[torch.Tensor] = []
        self.pre_act = ACT_NONE            # consumers must apply this activation on load (virtual pre-activation)
        self.clog = self.shape[3]          # logical (Keras) channel count; shape[3] is the physical one (phys_channels)

    @property
    def M(self):
        return self.shape[0] * self.shape[1] * self.shape[2]

    @property
    def C(self):
        return self.shape[3]


class FlatNode:
    def __init__(self, layer: Layer, inputs: List[int], output: int, shape, dtype):
        self.layer, self.inputs, self.output, self.shape, self.dtype = layer, inputs, output, shape, dtype
        self.calls = 1
        self.absorbed = False


def flatten(model: Model) -> Tuple[List[FlatNode], List[int], List[int]]:
    """Expand nested models; identical (layer, inputs) applications are merged (the reference calls the shared
    `base` twice on the same image when boundary refinement is on, ss.py:802 and :930)."""
    nodes: List[FlatNode] = []
    cse: Dict[Tuple, int] = {}
    counter = [0]

    def new_id():
        counter[0] += 1
        return counter[0]

    def run(m: Model, in_ids: List[int]) -> List[int]:
        env: Dict[int, int] = {id(t): i for t, i in zip(m.inputs, in_ids)}
        for n in m.nodes:
            out_t = n.outputs
            if isinstance(n.layer, InputLayer):
                if id(out_t[0]) not in env:
                    raise ValueError(f"unbound input {n.layer.name}")
                continue
            ins = [env[id(t)] for t in n.inputs]
            if isinstance(n.layer, Model):
                outs = run(n.layer, ins)
                for t, o in zip(out_t, outs):
                    env[id(t)] = o
                continue
            key = (id(n.layer), tuple(ins))
            if key in cse:
                fn = nodes[cse[key]]
                fn.calls += 1
                env[id(out_t[0])] = fn.output
                continue
            oid = new_id()
            cse[key] = len(nodes)
            nodes.append(FlatNode(n.layer, ins, oid, out_t[0].shape, out_t[0].dtype))
            env[id(out_t[0])] = oid
        return [env[id(t)] for t in m.outputs]

    in_ids = [new_id() for _ in model.inputs]
    out_ids = run(model, in_ids)
    return nodes, in_ids, out_ids


class ParamStore:
    """Flat fp32 arenas: trainable weights (L2-regularised kernels first), their gradients and Adam moments, and
    the non-trainable BatchNormalization moving statistics."""

    def __init__(self, layers_in_order: List[Layer], device, phys=phys_channels, no_pad=frozenset()):
        self.device = device
        self.phys = phys
        self.no_pad = frozenset(no_pad)   # channel counts kept at their logical pitch (a property of the graph)
        # ONE store per (model, device), shared by every Plan of the model (training / inference, any batch size):
        # `version` counts device-side weight changes so that each plan knows when its derived operands (bf16 copies,
        # folded BN) are stale; `host_stale` = the device holds newer weights than the layers' host arrays (after an
        # optimizer step); `dev_stale` = the host arrays are newer (after set_weights / load).
        self.version = 0
        self.host_stale = False
        self.dev_stale = False
        self.entries: Dict[Tuple[int, str], Tuple[str, int, Tuple]] = {}
        reg, plain, frozen = [], [], []
        for l in layers_in_order:
            names = list(l.weight_names())
            if "beta" in names and "gamma" in names:
                # BatchNormalization: keep [dbeta | dgamma] adjacent in the gradient arena so that the backward
                # reduction kernel (red[0..C) = sum g, red[C..2C) = sum g*xhat) accumulates straight into it
                # (the lists are reversed below, hence gamma before beta here)
                names = [n for n in names if n not in ("beta", "gamma")] + ["gamma", "beta"]
            for n in names:
                w = l._weights[n]
                if not l._trainable[n]:
                    frozen.append((l, n, w))
                elif n == "kernel" and getattr(l, "kernel_regularizer", None) is not None:
                    reg.append((l, n, w))
                else:
                    plain.append((l, n, w))
        self.l2 = 0.0
        for l, _, _ in reg:
            lam = l.kernel_regularizer.l2
            if self.l2 and abs(lam - self.l2) > 0:
                raise ValueError("a single L2 coefficient per model is supported (hps.weight_decay)")
            self.l2 = lam

        def lay(items, kind, start=0):
            off = start
            for l, n, w in items:
                pshape = _phys_shape(n, w.shape, self.phys)       # physical (channel-padded) layout, see phys_channels
                self.entries[(id(l), n)] = (kind, off, pshape)
                off += _ceil8(int(np.prod(pshape)))    # keep every parameter 32-byte aligned
            return off

        # REVERSE forward order: backward finishes the head first and block1 last, so the gradients of a growing PREFIX
        # of each region are final while backward is still running — the data-parallel exchange all-reduces prefix
        # increments behind backward instead of the whole arena after it (trainer.py).
        reg.reverse()
        plain.reverse()
        self.n_reg = lay(reg, "w")
        self.n_train = lay(plain, "w", self.n_reg)
        self.order_w = [(id(l), n) for l, n, _ in reg + plain]          # arena order of the trainable entries
        self.n_frozen = lay(frozen, "f")
        self._items = reg + plain + frozen
        self.w = torch.zeros(max(self.n_train, 8), dtype=torch.float32, device=device)
        self.g = torch.zeros_like(self.w)
        self.m = torch.zeros_like(self.w)
        self.v = torch.zeros_like(self.w)
        self.f = torch.zeros(max(self.n_frozen, 8), dtype=torch.float32, device=device)
        self.num_params = int(sum(w.size for _, n, w in reg + plain))
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=device)    # dropout stream position
        for l in layers_in_order:
            l._stores.append(self)

    def view(self, layer: Layer, name: str, grad=False, arena: Optional[str] = None) -> torch.Tensor:
        """Contiguous view in the PHYSICAL (channel-padded) shape — what the kernels see.  `arena`: "w" weights,
        "g" gradients, "m" / "v" Adam moments (trainable parameters only)."""
        kind, off, shape = self.entries[(id(layer), name)]
        n = int(np.prod(shape))
        if kind == "w":
            buf = {"w": self.w, "g": self.g, "m": self.m, "v": self.v}[arena or ("g" if grad else "w")]
        else:
            buf = self.f
        return buf[off:off + n].view(shape)

    def logical(self, layer: Layer, name: str, grad=False, arena: Optional[str] = None) -> torch.Tensor:
        """The Keras-shaped part of a parameter (pad channels sliced away)."""
        v = self.view(layer, name, grad, arena)
        return v[tuple(slice(0, d) for d in layer._weights[name].shape)]

    def trainable_items(self):
        """(layer, weight name) of every trainable parameter, arena order."""
        return [(l, n) for l, n, _ in self._items if self.entries[(id(l), n)][0] == "w"]

    def mark_updated(self):
        """The device weights were changed by an optimizer step."""
        self.version += 1
        self.host_stale = True

    def sync_host(self):
        """Bring the layers' host arrays up to date with the device (get_weights / save after training)."""
        if self.host_stale:
            self.download()

    def sync_device(self):
        if self.dev_stale:
            self.upload()

    def has(self, layer: Layer, name: str) -> bool:
        return (id(layer), name) in self.entries

    def upload(self):
        host_w = np.zeros(self.w.numel(), dtype=np.float32)
        host_f = np.zeros(self.f.numel(), dtype=np.float32)
        for l, n, w in self._items:
            kind, off, pshape = self.entries[(id(l), n)]
            dst = (host_w if kind == "w" else host_f)[off:off + int(np.prod(pshape))].reshape(pshape)
            dst[tuple(slice(0, d) for d in w.shape)] = w        # pad channels stay zero
        self.w.copy_(torch.from_numpy(host_w))
        self.f.copy_(torch.from_numpy(host_f))
        self.version += 1
        self.host_stale = self.dev_stale = False

    def download(self):
        host_w, host_f = self.w.cpu().numpy(), self.f.cpu().numpy()
        for l, n, w in self._items:
            kind, off, pshape = self.entries[(id(l), n)]
            src = (host_w if kind == "w" else host_f)[off:off + int(np.prod(pshape))].reshape(pshape)
            w[...] = src[tuple(slice(0, d) for d in w.shape)]
        self.host_stale = False


class Plan:
    """A model lowered for one (batch size, training flag, dtype).  See module docstring."""

    def __init__(self, model: Model, batch_size: int, training: bool = False, dtype: Optional[str] = None,
                 device: Optional[str] = None, dropout_seed: int = 1024, fused_tail: bool = True,
                 fuse_bn_dw: bool = True, implicit_conv: bool = True, concat_in_place: bool = True,
                 fuse_bn_pool: bool = True, keep_scratch: bool = False):
        fake = getattr(ops, "FAKE", False)       # tests/fake_ops.py test double (host-logic tests without a GPU)
        if not torch.cuda.is_available() and not fake:
            raise RuntimeError("engine.Plan needs a CUDA device: there is no CPU execution path")
        self.model, self.N, self.training = model, batch_size, training
        self.device = torch.device("cpu") if fake else torch.device(device or f"cuda:{torch.cuda.current_device()}")
        act_dtype = dtype or model.inputs[0].dtype
        if act_dtype not in _TORCH_DT:
            raise ValueError(f"hps.dtype must be 'float32' or 'bfloat16', got {act_dtype!r}")
        self.dt = _TORCH_DT[act_dtype]
        self.bf16 = self.dt == torch.bfloat16
        self.fused_tail = fused_tail
        self.fuse_bn_dw = fuse_bn_dw        # Conv->BN->ReLU->depthwise: BN+ReLU applied on load (False: A/B, materialise it)
        self.fuse_bn_pool = fuse_bn_pool         # Conv->BN->MaxPooling2D: BN applied inside the pool, fused backward
        self.concat_in_place = concat_in_place   # Conv->BN(->ReLU) read only by a Concatenate writes its slice directly
        self.implicit_conv = implicit_conv  # dense 3x3 VALID stride-1 convs as implicit GEMMs (False: A/B, im2col + GEMM)
        self.dropout_seed = dropout_seed
        # introspection for the parity tests (tests/teacher.py): named storage points (value / gradient getters),
        # activation decision sites and max-pool winners.  `keep_scratch` gives every macro-op its own backward
        # scratch (dy / dA) instead of two alternating slots, so those gradients can be read after the step.
        self.keep_scratch = keep_scratch
        self.trace: Dict[str, dict] = {}
        self.act_sites: Dict[str, tuple] = {}
        self.pool_sites: Dict[str, torch.Tensor] = {}
        self._tname: Dict[int, str] = {}
        self.fwd: List[Callable[[], None]] = []
        self.bwd: List[Tuple[Callable[[], None], bool, Optional[int]]] = []   # (launch, side-stream ok, scratch slot)
        self.side_stream = None            # set by the Trainer: filter-gradient kernels overlap the main backward chain
        self._bwd_macro = 0
        self._grad_done: List[Tuple[int, List[Tuple[int, str]]]] = []   # (bwd index, parameters final from there on)
        self.prep: List[Callable[[], None]] = []        # bf16 weight copies / BN folding, after each weight update
        self._wprep: List[Tuple] = []                   # GEMM weights re-quantised by ONE batched launch
        self._wprep_table = None
        self._bwd_thunks: List[Callable[[], None]] = []
        self.launches_fwd = self.launches_bwd = 0
        self._scratch: Dict[str, torch.Tensor] = {}
        self._stat_slices: List[Tuple[int, int]] = []
        self._stats_total = 0

        self.nodes, in_ids, out_ids = flatten(model)
        if len(in_ids) != 1 or len(out_ids) != 1:
            raise ValueError("Plan supports single-input single-output models (the DeepLabV3+ graph)")
        # channel counts that reach a Concatenate keep their logical pitch: a padded operand would interleave pad
        # channels into the concatenated axis and the consumer's kernel rows would no longer be a plain zero-extension
        shape_of = {n.output: n.shape for n in self.nodes}
        no_pad = {shape_of[i][-1] for n in self.nodes if isinstance(n.layer, L.Concatenate) for i in n.inputs
                  if i in shape_of}
        no_pad |= {n.shape[-1] for n in self.nodes if isinstance(n.layer, L.Concatenate)}
        self.phys = lambda c: c if c in no_pad else phys_channels(c)
        # ONE parameter store per (model, device): weights, gradients, Adam moments and BN moving statistics are shared
        # by every plan of the model, so predict / get_weights / a plan for another batch size see the trained state
        store = model._stores_by_device.get(str(self.device))
        if store is None:
            weight_layers = [l for l in model.flat_layers() if l._weights]
            store = ParamStore(weight_layers, self.device, self.phys, no_pad)
            store.upload()
            model._stores_by_device[str(self.device)] = store
            model._apply_pending_optimizer_state()
        elif store.no_pad != frozenset(no_pad):
            raise RuntimeError("internal: channel-pitch rule differs between plans of one model")
        self.params = store
        self.params.sync_device()
        self.step_counter = store.step_counter
        self._prep_version = -1

        self.values: Dict[int, Value] = {}
        H, W, Cin = model.inputs[0].shape[1:]
        self.x_in = Value((self.N, H, W, Cin), self.dt, self._alloc((self.N, H, W, Cin), self.dt), "image")
        self.values[in_ids[0]] = self.x_in
        self._lower(out_ids[0])
        self.stats = torch.zeros(max(self._stats_total, 8), dtype=torch.float32, device=self.device)
        self.ensure_current()
        self.graph = None
        self.finalize()

    # ------------------------------------------------------------------------------------------------ utils
    def _alloc(self, shape, dtype, zero=False):
        return (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)

    def scratch(self, key: str, numel: int, dtype) -> torch.Tensor:
        """Shared workspace for temporaries that only live inside one macro-op's backward."""
        nbytes = numel * torch.tensor([], dtype=dtype).element_size()
        cur = self._scratch.get(key)
        if cur is None or cur.numel() < nbytes:
            self._scratch[key] = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            # earlier closures captured views of the old buffer; force re-resolution by always going through _sv
        return self._scratch[key]

    def _sv(self, key: str, shape, dtype) -> torch.Tensor:
        numel = int(np.prod(shape))
        buf = self._scratch[key]
        esz = torch.tensor([], dtype=dtype).element_size()
        return buf[:numel * esz].view(dtype).view(shape)

    def _reserve(self, key: str, shape, dtype):
        self.scratch(key, int(np.prod(shape)), dtype)
        return lambda: self._sv(key, shape, dtype)

    def _stat_slot(self, C: int) -> Callable[[], torch.Tensor]:
        off = self._stats_total
        self._stats_total += 2 * C
        return lambda: self.stats[off:off + 2 * C]

    def _grad_of(self, v: Value) -> torch.Tensor:
        if v.grad is None:
            gd = torch.float32 if v.dtype == torch.float32 else self.dt
            v.grad = self._alloc(v.shape, gd)
        return v.grad

    def _grad_target(self, v: Value) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(buffer to write, addend) for a real gradient writer, given what the schedule has emitted so far."""
        g = self._grad_of(v)
        v.fold_hook = None                                   # a new writer: any earlier one is no longer the last
        if v.grad_written:
            return g, g
        v.grad_written = True
        if v.pending:
            return g, v.pending.pop()
        return g, None

    def _final_grad(self, v: Value) -> Optional[torch.Tensor]:
        """The complete gradient of `v` at the point its producer's backward runs (emits adds if needed)."""
        if not v.grad_written:
            if not v.pending:
                return None
            if len(v.pending) == 1:
                return v.pending[0]                      # pure pass-through: alias, no copy
            g = self._grad_of(v)
            a, b = v.pending.pop(), v.pending.pop()
            self.bwd_seq(lambda a=a, b=b, g=g: ops.add(a, b, g))
            v.grad_written = True
        g = self._grad_of(v)
        while v.pending:
            p = v.pending.pop()
            v.fold_hook = None                               # the gradient is completed by this add, not by the hook's launch
            self.bwd_seq(lambda p=p, g=g: ops.add(g, p, g))
        return g

    def _trace_value(self, name: str, getter, clog: int, shape=None):
        self.trace.setdefault(name, {})["value"] = (getter if callable(getter) else (lambda t=getter: t), clog, shape)

    def _trace_grad(self, name: str, getter, clog: int, shape=None):
        self.trace.setdefault(name, {})["grad"] = (getter if callable(getter) else (lambda t=getter: t), clog, shape)

    def traced(self, name: str, what: str = "value") -> Optional[torch.Tensor]:
        """The stored tensor (or its gradient) of a named storage point as a contiguous NHWC tensor with the LOGICAL
        channel count, or None when the plan does not materialise it."""
        ent = self.trace.get(name, {}).get(what)
        if ent is None:
            return None
        getter, clog, shape = ent
        t = getter()
        if t is None:
            return None
        if shape is not None:
            if t.dim() == 2 and t.shape[1] != shape[3]:
                t = t[:, :shape[3]]
            t = t.reshape(shape) if t.is_contiguous() else t.contiguous().view(shape)
        return t[..., :clog].contiguous()

    def dropout_mask(self, name: str) -> torch.Tensor:
        """keep / (1 - rate) multiplier the Dropout layer `name` applies at the CURRENT step counter (the counter-based
        generator is a pure function of seed, counter and element index): what a parity run injects into the oracle."""
        out, rate, seed = self.dropout_sites[name]
        ones = torch.ones_like(out.buf)
        m = torch.empty_like(out.buf)
        ops.dropout(ones, rate, seed, m, seed_offset=self.step_counter)
        return m[..., :out.clog].float()

    def decisions(self):
        """(masks, pool_taps): per activation site an int8 tensor (0 = clamped below, 1 = passed, 2 = clamped above,
        ReLU6 only) computed the way the backward kernels decide it — from the raw conv output and the BN scale/shift
        where the activation follows a BatchNormalization, from the stored input otherwise — and per max-pool site the
        uint8 winning tap (row-major inside the 3x3 window)."""
        masks = {}
        for site, ent in self.act_sites.items():
            if ent[0] == "bn":
                _, act, y, scale, shift, clog, shape = ent
                # the kernels decide on fmaf(y, scale, shift) (ONE rounding): y*scale is exact in fp64 and the sum rounds
                # once, so this reproduces their sign / clamp decision bit for bit.  A separate fp32 multiply and add rounds
                # twice and flips the decision of an element within an ulp of the threshold — about one parity run in
                # eight had such an element (the BN statistics differ in the last bit from run to run: fp32 atomics)
                z = (y.double().view(-1, y.shape[-1]) * scale.double() + shift.double()).float().view(shape)[..., :clog]
            else:
                _, act, v = ent
                z = v.buf.float()[..., :v.clog]
            m = (z > 0).to(torch.int8)
            if act == ACT_RELU6:
                m = m + (z >= 6).to(torch.int8)
            masks[site] = m
        taps = {k: (am[..., :c] if c else am) for k, (am, c) in self.pool_sites.items()}
        return masks, taps

    # ---------------------------------------------------------------------------------------------- lowering
    def _lower(self, out_id: int):
        nodes = self.nodes
        consumers: Dict[int, List[FlatNode]] = {}
        for n in nodes:
            for i in n.inputs:
                consumers.setdefault(i, []).append(n)
        producer = {n.output: n for n in nodes}
        shape_of_node = {n.output: n.shape for n in nodes}

        def sole(tid) -> Optional[FlatNode]:
            c = consumers.get(tid, [])
            return c[0] if len(c) == 1 and tid != out_id else None

        def act_code(layer) -> Optional[int]:
            if isinstance(layer, L.Activation) and layer.activation == "relu":
                return ACT_RELU
            if isinstance(layer, L.ReLU):
                return ACT_RELU6 if layer.max_value else ACT_RELU
            return None

        conv_like = (L.Conv2D, L.SeparableConv2D, L.DepthwiseConv2D)
        # ---- pattern pass: build macro-ops -------------------------------------------------------------
        macros: Dict[int, dict] = {}            # keyed by index of the node where the macro is emitted
        index = {id(n): i for i, n in enumerate(nodes)}
        for n in nodes:
            if n.absorbed:
                continue
            if isinstance(n.layer, L.ZeroPadding2D):
                c = sole(n.output)
                if c is None or not isinstance(c.layer, (L.DepthwiseConv2D, L.Conv2D)) or c.layer.padding != "valid":
                    raise NotImplementedError("ZeroPadding2D is only supported directly before a VALID convolution")
                n.absorbed = True
                c.explicit_pad = n.layer.padding
                c.inputs = list(n.inputs)
                consumers[n.inputs[0]] = [c if x is n else x for x in consumers[n.inputs[0]]]
                continue
            if isinstance(n.layer, conv_like):
                m = dict(kind="conv", conv=n, bn=None, act=ACT_NONE, other=None, out=n.output, tail=n)
                nxt = sole(n.output)
                if nxt is not None and isinstance(nxt.layer, L.BatchNormalization):
                    m["bn"], m["out"], m["tail"] = nxt, nxt.output, nxt
                    nxt.absorbed = True
                    nxt = sole(nxt.output)
                    if nxt is not None and act_code(nxt.layer) is not None:
                        m["act"], m["out"], m["tail"] = act_code(nxt.layer), nxt.output, nxt
                        nxt.absorbed = True
                        nxt = sole(nxt.output)
                        # Conv -> BN -> ReLU whose only reader is a dense-tap depthwise stage: the BN+ReLU output is
                        # never written, the reader applies it on load (see _BnActValue)
                        if (nxt is not None and isinstance(nxt.layer, (L.SeparableConv2D, L.DepthwiseConv2D))
                                and nxt.layer.strides[0] == 1 and tuple(nxt.layer.dilation_rate) == (1, 1)
                                and nxt.layer.padding == "same"):
                            m["virt"] = True
                    elif nxt is not None and isinstance(nxt.layer, L.MaxPooling2D):
                        m["pool_virt"] = True        # Conv -> BN -> MaxPooling2D: BN applied inside the pool (see _BnPoolValue)
                    elif nxt is not None and isinstance(nxt.layer, L.Add) and not nxt.absorbed:
                        other = [i for i in nxt.inputs if i != m["out"]]
                        if len(other) == 1:
                            m["other"], m["out"], m["tail"] = other[0], nxt.output, nxt
                            nxt.absorbed = True
                n.absorbed = True
                macros[index[id(m["tail"])]] = m
            elif isinstance(n.layer, L.MaxPooling2D):
                m = dict(kind="maxpool", node=n, other=None, out=n.output, tail=n)
                nxt = sole(n.output)
                if nxt is not None and isinstance(nxt.layer, L.Add) and not nxt.absorbed:
                    other = [i for i in nxt.inputs if i != n.output]
                    if len(other) == 1:
                        m["other"], m["out"], m["tail"] = other[0], nxt.output, nxt
                        nxt.absorbed = True
                n.absorbed = True
                macros[index[id(m["tail"])]] = m

        # ---- Concatenate inputs written in place: a training-mode Conv -> BN (-> ReLU) macro-op whose only reader is
        # a Concatenate writes its output straight into its channel slice of the concat buffer (bn_train_apply's ld_out)
        # and reads its gradient from the same slice of the concat gradient (ld_dz): no slice copies either way
        self._concat_slot: Dict[int, Tuple[int, int, int]] = {}       # tensor id -> (concat tensor id, offset, Ct)
        self._concat_buf: Dict[int, torch.Tensor] = {}
        if self.training and self.concat_in_place:
            for n in nodes:
                if isinstance(n.layer, L.Concatenate) and n.output != out_id:
                    Ct, off = sum(shape_of_node[i][-1] for i in n.inputs), 0
                    for i in n.inputs:
                        src, reader = i, n
                        # look through identity layers (x1 resize, 1x1 average pool: conf.json:51)
                        while sole(src) is reader and src in producer and (
                                (isinstance(producer[src].layer, L.ResizeImages) and tuple(producer[src].layer.factors) == (1, 1))
                                or (isinstance(producer[src].layer, L.AveragePooling2D) and producer[src].layer.pool_size[0] == 1)):
                            reader, src = producer[src], producer[src].inputs[0]
                        if sole(src) is reader:
                            self._concat_slot[src] = (n.output, off, Ct)
                        off += shape_of_node[i][-1]

        # ---- emission in topological order ---------------------------------------------------------------
        for i, n in enumerate(nodes):
            if i in macros:
                m = macros[i]
                if m["kind"] == "conv":
                    self._emit_conv(m, out_id)
                else:
                    self._emit_maxpool(m)
                continue
            if n.absorbed:
                continue
            lay = n.layer
            code = act_code(lay)
            if code is not None:
                self._emit_act(n, code, consumers)
            elif isinstance(lay, L.Activation) and lay.activation == "softmax":
                if n.output != out_id:
                    raise NotImplementedError("softmax is only supported as the model output")
                self._emit_softmax_tail(n, producer)
            elif isinstance(lay, L.Activation) and lay.activation == "linear":
                self.values[n.output] = self.values[n.inputs[0]]
            elif isinstance(lay, L.Add):
                self._emit_add(n)
            elif isinstance(lay, L.AveragePooling2D):
                self._emit_avgpool(n)
            elif isinstance(lay, L.ResizeImages):
                self._emit_resize(n, out_id, consumers)
            elif isinstance(lay, L.Concatenate):
                self._emit_concat(n)
            elif isinstance(lay, L.Dropout):
                self._emit_dropout(n)
            elif isinstance(lay, L.BatchNormalization):
                raise NotImplementedError(f"BatchNormalization '{lay.name}' not preceded by a convolution")
            else:
                raise NotImplementedError(f"layer type {type(lay).__name__} is not on the hot path")
        self.out_value = self.values.get(out_id)

    def _input_of(self, tid: int, allow_pre_act: bool) -> Value:
        v = self.values[tid]
        if v.pre_act != ACT_NONE and not allow_pre_act:
            raise NotImplementedError("internal: virtual activation reached a consumer that cannot fuse it")
        return v

    # ---- conv macro-op ---------------------------------------------------------------------------------
    def _emit_conv(self, m: dict, out_id: int):
        """Conv / SeparableConv / Depthwise (+ BatchNormalization)(+ ReLU/ReLU6)(+ residual Add) macro-op: geometry, output
        value, parameters and BN state here; launches in _emit_conv_forward / _emit_conv_backward (they share the
        namespace `cx` of this set-up)."""
        P, N = self.params, self.N
        node: FlatNode = m["conv"]
        lay = node.layer
        x = self._input_of(node.inputs[0], allow_pre_act=isinstance(lay, (L.SeparableConv2D, L.DepthwiseConv2D)))
        _, H, W, Cin = x.shape
        training = self.training
        bn_node, act, other_id = m["bn"], m["act"], m["other"]
        is_sep, is_dw = isinstance(lay, L.SeparableConv2D), isinstance(lay, L.DepthwiseConv2D)
        stride = lay.strides[0]
        dil = lay.dilation_rate
        k = 3 if (is_sep or is_dw) else lay.kernel_size[0]
        # geometry (TF SAME / VALID, or the explicit ZeroPadding2D folded in front of a VALID conv)
        if getattr(node, "explicit_pad", None) is not None:
            (pt, pb), (pl, pr) = node.explicit_pad
            Ho = (H + pt + pb - ((k - 1) * dil[0] + 1)) // stride + 1
            Wo = (W + pl + pr - ((k - 1) * dil[1] + 1)) // stride + 1
        else:
            Ho, Wo, pt, pl = ops.conv_geometry(H, W, k, stride, dil, lay.padding)
        Mo = N * Ho * Wo
        Cout = Cin if is_dw else self.phys(lay.filters)
        Cout_log = x.clog if is_dw else lay.filters
        out_shape = (N, Ho, Wo, Cout)
        assert tuple(m["conv"].shape[1:]) == (Ho, Wo, Cout_log), (lay.name, m["conv"].shape, out_shape)

        # output dtype: logits (conv without BN feeding the softmax tail) stay fp32
        is_logits = bn_node is None and not is_dw
        y_dtype = torch.float32 if is_logits else self.dt
        ld_out = Cout
        virt = (bool(m.get("virt")) and self.fuse_bn_dw and training and bn_node is not None and act != ACT_NONE
                and other_id is None and m["out"] != out_id and Cout % 8 == 0)
        gemm_like = not is_dw
        slot_c = self._concat_slot.get(m["out"])
        in_place = (slot_c is not None and not virt and training and bn_node is not None and other_id is None
                    and Cout % 8 == 0 and Cout == Cout_log and y_dtype == self.dt)
        pool_virt = (bool(m.get("pool_virt")) and self.fuse_bn_pool and (self.bf16 or FORCE_BNRED) and training
                     and bn_node is not None and act == ACT_NONE and other_id is None and m["out"] != out_id
                     and Cout % 8 == 0 and gemm_like)
        if virt:
            out = _BnActValue(out_shape, y_dtype, lay.name, act)       # buffers / BN operands attached below
        elif pool_virt:
            out = _BnPoolValue(out_shape, y_dtype, lay.name)
        elif in_place:
            cid, c_off, Ct = slot_c
            if cid not in self._concat_buf:
                self._concat_buf[cid] = self._alloc((N, Ho, Wo, Ct), self.dt)
            ld_out = Ct
            out = Value(out_shape, y_dtype, self._concat_buf[cid].view(N * Ho * Wo, Ct)[:, c_off:c_off + Cout], lay.name)
            out.concat_slice = True
        else:
            out = Value(out_shape, y_dtype, self._alloc((N, Ho, Wo, ld_out), y_dtype), lay.name)
        out.clog = Cout_log
        self.values[m["out"]] = out
        tname = bn_node.layer.name if bn_node is not None else lay.name
        self._tname[m["out"]] = tname
        other = self.values[other_id] if other_id is not None else None
        needs_in_grad = x.needs_grad
        out.needs_grad = training

        # ---- parameters
        dw_w = P.view(lay, "depthwise_kernel").view(3, 3, Cin) if (is_sep or is_dw) else None
        dw_g = P.view(lay, "depthwise_kernel", grad=True).view(3, 3, Cin) if (is_sep or is_dw) else None
        gemm = not is_dw
        implicit = implicit_same = False
        if gemm:
            wname = "pointwise_kernel" if is_sep else "kernel"
            Kdim = Cin if (is_sep or k == 1) else 9 * Cin
            Kp = _ceil8(Kdim)
            Np = _ceil8(Cout)
            w32 = P.view(lay, wname).view(Kdim, Cout)
            g32 = P.view(lay, wname, grad=True).view(Kdim, Cout)
            # dense 3x3 VALID stride-1 conv (Xception block1_conv2): implicit GEMM straight from the NHWC tensor
            implicit = (self.implicit_conv and (self.bf16 or FORCE_IMPLICIT) and k == 3 and not is_sep and stride == 1
                        and tuple(dil) == (1, 1) and lay.padding == "valid"
                        and getattr(node, "explicit_pad", None) is None and not is_logits and other_id is None
                        and ops.conv3x3_valid_supported(Cin, Cout))
            # dense 3x3 SAME stride-1 conv (the logits convolution, 304 channels at 256^2 after boundary refinement):
            # implicit GEMMs over tap-shifted TMA windows, no [pixels, 9*Cin] column matrix
            implicit_same = (self.implicit_conv and (self.bf16 or FORCE_IMPLICIT) and k == 3 and not is_sep
                             and stride == 1 and tuple(dil) == (1, 1) and lay.padding == "same"
                             and getattr(node, "explicit_pad", None) is None and other_id is None
                             and (bn_node is None or Cout % 8 == 0) and ops.conv3x3_same_supported(Cin, Cout))
            if implicit_same:
                wdt = torch.bfloat16 if self.bf16 else torch.float32          # (fp32 only under FORCE_IMPLICIT, CPU tests)
                kp_tap = (Cout + 63) // 64 * 64
                wt = self._alloc((Cout, Kp), wdt, zero=True)
                wd = self._alloc((Cin, 9 * kp_tap), wdt, zero=True) if training else None
                self._wprep.append((w32, Kdim, Cout, wt, Kp, None, Np))

                def prep_same():
                    if wd is not None:       # wd[c, tap*kp + o] = W[tap, c, o]
                        wd.view(Cin, 9, kp_tap)[:, :, :Cout].copy_(w32.view(9, Cin, Cout).permute(1, 0, 2))
                self.prep.append(prep_same)
            elif implicit:
                wdt = torch.bfloat16 if self.bf16 else torch.float32          # (fp32 only under FORCE_IMPLICIT, CPU tests)
                wt = self._alloc((Cout, Kp), wdt, zero=True)                  # forward B operand = the im2col GEMM's
                wd = self._alloc((Cin, 9 * Cout), wdt) if training else None  # input-gradient B operand
                self._wprep.append((w32, Kdim, Cout, wt, Kp, None, Np))

                def prep_implicit():
                    if wd is not None:
                        wd.view(Cin, 9, Cout).copy_(w32.view(9, Cin, Cout).permute(1, 0, 2))
                self.prep.append(prep_implicit)
            elif self.bf16:
                wt = self._alloc((Cout, Kp), torch.bfloat16, zero=True)       # [N,K] K-major: forward B operand
                wn = self._alloc((Kdim, Np), torch.bfloat16, zero=True) if training else None   # dgrad B operand
                self._wprep.append((w32, Kdim, Cout, wt, Kp, wn, Np))
        # ---- BatchNormalization state
        if bn_node is not None:
            bn = bn_node.layer
            C = Cout
            gamma = P.view(bn, "gamma") if P.has(bn, "gamma") else None
            beta = P.view(bn, "beta") if P.has(bn, "beta") else None
            mm, mv = P.view(bn, "moving_mean"), P.view(bn, "moving_variance")
            scale, shift = self._alloc((C,), torch.float32), self._alloc((C,), torch.float32)
            if training:
                mean, invstd = self._alloc((C,), torch.float32), self._alloc((C,), torch.float32)
                stat = self._stat_slot(C)
                # BN parameter gradients: [dbeta | dgamma] must be contiguous for the reduce kernel.  When both
                # exist and C % 8 == 0 they are adjacent in the gradient arena (ParamStore) and the kernel writes
                # there directly; otherwise it reduces into a scratch slot that is then added to the arena.
                red_direct = False
                if gamma is not None and beta is not None:
                    gb_, gg_ = P.view(bn, "beta", grad=True), P.view(bn, "gamma", grad=True)
                    if gb_.data_ptr() + 4 * C == gg_.data_ptr():
                        red_direct = True
                        k0 = P.entries[(id(bn), "beta")][1]
                        red_slot = lambda k0=k0, C=C: P.g[k0:k0 + 2 * C]
                if not red_direct:
                    red_slot = self._stat_slot(C)
                y = self._alloc((N, Ho, Wo, Cout), self.dt)          # raw conv output, saved for backward
                if virt:
                    out.attach(y, scale, shift, mean, invstd, red_slot)
                if pool_virt:
                    out.attach(y, scale, shift, mean, invstd, red_slot, Mo)
            else:
                self.prep.append(lambda: ops.bn_fold(gamma, beta, mm, mv, C, bn.epsilon, scale, shift))
                y = None
        else:
            y = out.buf

        in_act = x.pre_act
        in_sc, in_sh = getattr(x, "pre_scale", None), getattr(x, "pre_shift", None)
        pad4 = (Ho, Wo, pt, pl)
        xb = x.buf
        launches_f = 0
        # introspection (parity tests): storage points and the activation decision site of this macro-op
        if bn_node is not None and training:
            self._trace_value(f"{tname}/y", y, Cout_log, out_shape)
            if act != ACT_NONE:
                self.act_sites[tname] = ("bn", act, y, scale, shift, Cout_log, out_shape)
        elif bn_node is None:
            self._trace_value(f"{tname}/y", out.buf, Cout_log, out_shape)
        if bn_node is not None and not virt and not pool_virt:
            self._trace_value(f"{tname}/out", lambda: out.buf, Cout_log, out_shape)

        cx = _MacroScope(**{k: v for k, v in locals().items() if k not in ("self", "cx")})
        self._emit_conv_forward(cx)
        if training:
            self._emit_conv_backward(cx)

    def _emit_conv_forward(self, cx):
        """Forward launches of a conv macro-op (A operand: depthwise stage / subsample / im2col / implicit; GEMM with
        epilogue; training BN statistics + apply, or the BN finished inside the reader)."""
        Cin, Cout, Ho, Kdim, Kp, Mo, N, Wo = cx.Cin, cx.Cout, cx.Ho, cx.Kdim, cx.Kp, cx.Mo, cx.N, cx.Wo
        act, beta, bn, bn_node = cx.act, cx.beta, cx.bn, cx.bn_node
        dil, dw_w, gamma, gemm = cx.dil, cx.dw_w, cx.gamma, cx.gemm
        implicit, implicit_same, in_act, in_sc = cx.implicit, cx.implicit_same, cx.in_act, cx.in_sc
        in_sh, invstd, is_dw, is_sep = cx.in_sh, cx.invstd, cx.is_dw, cx.is_sep
        k, launches_f, lay, ld_out = cx.k, cx.launches_f, cx.lay, cx.ld_out
        mean, mm, mv, other = cx.mean, cx.mm, cx.mv, cx.other
        out, pad4, pl, pool_virt = cx.out, cx.pad4, cx.pl, cx.pool_virt
        pt, scale, shift, stat = cx.pt, cx.scale, cx.shift, cx.stat
        stride, training, virt, w32, wt, x, xb, y = cx.stride, cx.training, cx.virt, cx.w32, cx.wt, cx.x, cx.xb, cx.y
        # ---- forward ------------------------------------------------------------------------------------
        # A operand of the GEMM
        if is_sep or is_dw:
            d = self._alloc((N, Ho, Wo, Cin), self.dt) if is_sep else None
            dw_out = d if is_sep else (y if (bn_node is not None and training) else None)
            if is_dw and dw_out is None:
                # inference depthwise+BN: conv into `out`, then fold BN in place
                dw_out = out.buf
            fold = getattr(x, "bn_fold", None)
            # inference DepthwiseConv2D -> BN (-> ReLU/ReLU6) (MobileNetV2): the folded BN runs in the convolution's epilogue
            dw_epi = (is_dw and bn_node is not None and not training and other is None and in_act == ACT_NONE
                      and in_sc is None and fold is None)
            cx.dw_epi = dw_epi
            if dw_epi:
                self.fwd.append(lambda: ops.dwconv3x3_fwd_epi(xb, dw_w, scale, shift, act, stride, dil, out=out.buf,
                                                              pad=pad4))
            elif fold is not None:
                self.fwd.append(lambda: ops.dwconv3x3_bn_fwd(
                    xb, dw_w, fold["sums"](), fold["gamma"], fold["beta"], fold["mm"], fold["mv"], fold["count"],
                    fold["eps"], fold["momentum"], fold["updates"], in_act, in_sc, in_sh, x.bn_mean, x.bn_invstd,
                    out=dw_out, pad=pad4))
            else:
                self.fwd.append(lambda: ops.dwconv3x3_fwd(xb, dw_w, stride, dil, in_scale=in_sc, in_shift=in_sh,
                                                          in_act=in_act, out=dw_out, pad=pad4))
            launches_f += 1
            A, lda = d, Cin
            if is_sep:
                self._trace_value(f"{lay.name}/dw", d, x.clog, (N, Ho, Wo, Cin))
        elif k == 1 and stride == 1:
            A, lda = xb, Cin
        elif k == 1:
            xs = self._alloc((N, Ho, Wo, Cin), self.dt)
            self.fwd.append(lambda: ops.subsample_fwd(xb, stride, out=xs))
            launches_f += 1
            A, lda = xs, Cin
        elif implicit or implicit_same:
            A, lda = None, Kp
        else:
            col = self._alloc((Mo, Kp), self.dt)
            self.fwd.append(lambda: ops.im2col3x3(xb, stride, dil[0], Ho, Wo, pt, pl, Kp, out=col))
            launches_f += 1
            A, lda = col, Kp

        fuse_epi = (bn_node is not None and not training)      # inference: BN folded into the GEMM epilogue
        addend_f = other.buf if other is not None else None
        if gemm:
            Kg = lda if (k == 3 and not is_sep) else Kdim
            if implicit_same:
                tgt = out.buf if (fuse_epi or bn_node is None) else y
                stats_fn = stat if (bn_node is not None and training) else None
                self.fwd.append(lambda: ops.conv3x3_same_fwd(
                    xb, wt, tgt, Cout, ldw=Kp, col_scale=scale if fuse_epi else None, col_shift=shift if fuse_epi else None,
                    act=act if fuse_epi else ACT_NONE, col_stats=stats_fn() if stats_fn else None))
            elif implicit:
                tgt = out.buf if (fuse_epi or bn_node is None) else y
                stats_fn = stat if (bn_node is not None and training) else None
                self.fwd.append(lambda: ops.conv3x3_valid_fwd(
                    xb, wt, tgt, Cout, ldw=Kp, col_scale=scale if fuse_epi else None, col_shift=shift if fuse_epi else None,
                    act=act if fuse_epi else ACT_NONE, col_stats=stats_fn() if stats_fn else None))
            elif self.bf16:
                tgt = out.buf if (fuse_epi or bn_node is None) else y
                stats_fn = stat if (bn_node is not None and training) else None
                self.fwd.append(lambda: ops.gemm_bf16(
                    A, wt, Mo, Cout, Kg, tgt, lda=lda, ldb=Kp, ldc=Cout,
                    col_scale=scale if fuse_epi else None, col_shift=shift if fuse_epi else None,
                    act=act if fuse_epi else ACT_NONE, addend=addend_f if fuse_epi else None, ld_addend=Cout,
                    col_stats=stats_fn() if stats_fn else None))
            else:
                tgt = out.buf if (fuse_epi or bn_node is None) else y
                self.fwd.append(lambda: ops.gemm_simt(
                    A, lda, 1, w32, Cout, 1, tgt, Cout, Mo, Cout, Kdim,
                    col_scale=scale if fuse_epi else None, col_shift=shift if fuse_epi else None,
                    act=act if fuse_epi else ACT_NONE, addend=addend_f if fuse_epi else None, ld_addend=Cout))
            launches_f += 1
        if bn_node is not None and training:
            if not (gemm and (self.bf16 or implicit or implicit_same)):
                self.fwd.append(lambda: ops.bn_stats(y, Mo, Cout, stat()))
                launches_f += 1
            upd = bn_node.calls
            if virt and (self.bf16 or FORCE_BNRED):
                # statistics -> scale/shift (+ moving statistics) happen inside the reader's forward kernel
                # (dlv3p_dwconv3x3_bn_fwd): nothing to launch here
                out.bn_fold = dict(sums=stat, gamma=gamma, beta=beta, mm=mm, mv=mv, count=Mo, eps=bn.epsilon,
                                   momentum=bn.momentum, updates=upd)
            elif virt or pool_virt:
                # statistics -> scale/shift (+ moving statistics); the BN (+ReLU) map itself runs inside the reader
                for r in range(upd):
                    self.fwd.append(lambda r=r: ops.bn_finalize(stat(), gamma, beta, mm, mv, Cout, Mo, bn.epsilon,
                                                                bn.momentum, scale, shift, mean, invstd, True))
                    launches_f += 1
            elif Cout % 8 == 0:
                # statistics -> scale/shift, moving-statistics update and BN+activation(+add) in one launch
                self.fwd.append(lambda: ops.bn_train_apply(y, Mo, Cout, stat(), gamma, beta, mm, mv, Mo, bn.epsilon,
                                                           bn.momentum, upd, act, out.buf, scale, shift, mean, invstd,
                                                           addend=addend_f, ld_out=ld_out))
                launches_f += 1
            else:
                for r in range(upd):
                    self.fwd.append(lambda r=r: ops.bn_finalize(stat(), gamma, beta, mm, mv, Cout, Mo, bn.epsilon,
                                                                bn.momentum, scale, shift, mean, invstd, True))
                    launches_f += 1
                self.fwd.append(lambda: ops.affine_act(y, Mo, Cout, out.buf, scale, shift, act, addend=addend_f))
                launches_f += 1
        elif bn_node is not None and is_dw and not cx.dw_epi:
            self.fwd.append(lambda: ops.affine_act(out.buf, Mo, Cout, out.buf, scale, shift, act, addend=addend_f))
            launches_f += 1
        self.launches_fwd += launches_f

        cx.A = A
        cx.lda = lda

    def _emit_conv_backward(self, cx):
        """Backward schedule of a conv macro-op (deferred: finalize() runs the thunks in reverse topological order):
        BN backward -> dy; filter gradient (side stream); input gradient GEMM / implicit conv; depthwise backward."""
        A, Cin, Cout, Cout_log, Ho, Kdim, Kp, Mo = cx.A, cx.Cin, cx.Cout, cx.Cout_log, cx.Ho, cx.Kdim, cx.Kp, cx.Mo
        N, Np, P, Wo, act, beta, bn, bn_node = cx.N, cx.Np, cx.P, cx.Wo, cx.act, cx.beta, cx.bn, cx.bn_node
        dil, dw_g, dw_w, g32 = cx.dil, cx.dw_g, cx.dw_w, cx.g32
        gamma, gemm, implicit, implicit_same = cx.gamma, cx.gemm, cx.implicit, cx.implicit_same
        in_act, invstd, is_sep, k = cx.in_act, cx.invstd, cx.is_sep, cx.k
        lay, lda, mean, needs_in_grad = cx.lay, cx.lda, cx.mean, cx.needs_in_grad
        other, out, out_shape, pad4 = cx.other, cx.out, cx.out_shape, cx.pad4
        pl, pool_virt, pt, red_direct = cx.pl, cx.pool_virt, cx.pt, cx.red_direct
        red_slot, scale, shift, stride = cx.red_slot, cx.scale, cx.shift, cx.stride
        tname, virt, w32, wd = cx.tname, cx.virt, cx.w32, cx.wd
        wn, x, xb, y, y_dtype = cx.wn, cx.x, cx.xb, cx.y, cx.y_dtype
        # ---- backward (emitted in forward order; the list is reversed at the end, so write steps in REVERSE) --
        def sched():
            g = out.pool_bwd["g"] if pool_virt else self._final_grad(out)
            if g is None:
                raise RuntimeError(f"no gradient reaches {lay.name}")
            if other is not None:
                other.pending.append(g)
            slot = self._bwd_macro & 1                     # scratch copy used by this macro-op (see bwd_seq)
            skey = self._bwd_macro if self.keep_scratch else slot
            self._bwd_macro += 1
            self.bwd_seq(None, slot=slot)
            if not virt and not pool_virt:
                self._trace_grad(f"{tname}/out" if bn_node is not None else f"{tname}/y", g, Cout_log, out_shape)
            # dy: gradient w.r.t. the raw conv output
            if bn_node is not None:
                dy_get = self._reserve(f"dy{skey}", (Mo, Cout), self.dt)
                if self.keep_scratch:
                    self._trace_grad(f"{tname}/y", dy_get, Cout_log, out_shape)
                red = red_slot
                # virtual BN+ReLU output: g arrives as the gradient w.r.t. the BN output, ReLU mask already applied
                # by the reader's input-gradient kernel (which may also have produced the two reductions)
                act_b = ACT_NONE if virt else act
                if pool_virt:
                    # the BN output fed a MaxPooling2D directly: reductions over the POOLED tensors (gradient x raw
                    # winner values), then pool backward + BN input gradient in one pass straight into dy
                    pb = out.pool_bwd
                    Mp = g.shape[0] * g.shape[1] * g.shape[2]
                    self.bwd_seq(lambda: ops.bn_bwd_reduce(g, pb["ymax"], scale, shift, mean, invstd, ACT_NONE, Mp, Cout,
                                                           red()))
                    self.bwd_seq(lambda: ops.maxpool3x3s2_bn_bwd(g, pb["argmax"], y, scale, mean, invstd, red(), Mo,
                                                                 dy_get().view(N, Ho, Wo, Cout)))
                ld_g = g.stride(0) if g.dim() == 2 else Cout          # 2-D: a channel slice of a concat gradient
                hook = out.fold_hook
                folded = (hook is not None and not virt and not pool_virt and act_b == ACT_NONE and g is out.grad
                          and ld_g == Cout and y.dtype == g.dtype)
                if folded:
                    # the launch that wrote g last (a fused depthwise backward of the reader) reduces it against y as well
                    hook.update(bn_y=y, bn_mean=mean, bn_invstd=invstd, bn_red=red)
                if not pool_virt and not folded and not (virt and out.red_done):
                    self.bwd_seq(lambda: ops.bn_bwd_reduce(g, y, scale, shift, mean, invstd, act_b, Mo, Cout, red(),
                                                           ld_dz=ld_g))
                if not pool_virt:
                    self.bwd_seq(lambda: ops.bn_bwd_apply(g, y, scale, shift, mean, invstd, act_b, red(), Mo, Cout,
                                                          dy_get(), ld_dz=ld_g))
                # parameter gradients live in the stats arena; copy into the grad arena
                if beta is not None and not red_direct:
                    gb = P.view(bn, "beta", grad=True)
                    self.bwd_seq(lambda: gb.add_(red()[:Cout]))
                if gamma is not None and not red_direct:
                    gg = P.view(bn, "gamma", grad=True)
                    self.bwd_seq(lambda: gg.add_(red()[Cout:]))
            else:
                if y_dtype == torch.float32 and self.bf16:
                    # logits gradient arrives fp32 [M,Cout]; tensor-core operands need bf16 with ld % 8 == 0
                    dyb = self._alloc((Mo, Np), torch.bfloat16, zero=True)
                    self.bwd_seq(lambda: ops.cast2d(g, Cout, dyb, Np, Mo, Cout))
                    dy_get = lambda: dyb
                else:
                    dy_get = lambda: g
            ld_dy = Np if (bn_node is None and y_dtype == torch.float32 and self.bf16) else Cout

            if gemm and implicit_same:
                self.bwd_seq(lambda: ops.conv3x3_same_wgrad(xb, dy_get(), ld_dy, g32, Cout), side=True, slot=slot)
                if needs_in_grad:
                    tgt, addend2 = self._grad_target(x)
                    if addend2 is None:
                        self.bwd_seq(lambda: ops.conv3x3_same_dgrad(dy_get(), ld_dy, wd, x.shape, Cout, tgt))
                    else:
                        tmp_get = self._reserve(f"dA{skey}", x.shape, self.dt)
                        self.bwd_seq(lambda: ops.conv3x3_same_dgrad(dy_get(), ld_dy, wd, x.shape, Cout, tmp_get()))
                        self.bwd_seq(lambda: ops.add(tmp_get(), addend2, tgt))
            elif gemm and implicit:
                self.bwd_seq(lambda: ops.conv3x3_valid_wgrad(xb, dy_get().view(N, Ho, Wo, Cout), g32, Cout), side=True,
                             slot=slot)
                if needs_in_grad:
                    tgt, addend2 = self._grad_target(x)
                    if addend2 is None:
                        self.bwd_seq(lambda: ops.conv3x3_valid_dgrad(dy_get().view(N, Ho, Wo, Cout), wd, x.shape, Cout, tgt))
                    else:
                        tmp_get = self._reserve(f"dA{skey}", x.shape, self.dt)
                        self.bwd_seq(lambda: ops.conv3x3_valid_dgrad(dy_get().view(N, Ho, Wo, Cout), wd, x.shape, Cout,
                                                                     tmp_get()))
                        self.bwd_seq(lambda: ops.add(tmp_get(), addend2, tgt))
            elif gemm:
                # filter gradient
                if self.bf16:
                    self.bwd_seq(lambda: ops.gemm_wgrad_bf16(A, dy_get(), g32, Mo, Kdim, Cout, ldx=lda, ldy=ld_dy,
                                                             ldw=Cout), side=True, slot=slot)
                else:
                    self.bwd_seq(lambda: ops.gemm_simt(A, 1, lda, dy_get(), ld_dy, 1, g32, Cout, Kdim, Cout, Mo,
                                                       accumulate=True), side=True, slot=slot)
                need_dA = needs_in_grad
                if need_dA:
                    direct = (k == 1 and stride == 1 and not is_sep)
                    if direct:
                        tgt, addend = self._grad_target(x)
                        dA_get = lambda: tgt
                    else:
                        shape_dA = (Mo, lda)
                        dA_get = self._reserve(f"dA{skey}", shape_dA, self.dt)
                        addend = None
                        if is_sep and self.keep_scratch:
                            self._trace_grad(f"{lay.name}/dw", dA_get, x.clog, (N, Ho, Wo, Cin))
                    if self.bf16:
                        # dA[M,K] = dy[M,N] * W[K,N]^T : B operand = wn (rows = K, contraction over Np, zero padded)
                        self.bwd_seq(lambda: ops.gemm_bf16(dy_get(), wn, Mo, Kdim, ld_dy, dA_get(), lda=ld_dy,
                                                           ldb=Np, ldc=lda, addend=addend, ld_addend=lda))
                    else:
                        self.bwd_seq(lambda: ops.gemm_simt(dy_get(), ld_dy, 1, w32, 1, Cout, dA_get(), lda, Mo, Kdim,
                                                           Cout, addend=addend, ld_addend=lda))
                    if is_sep:
                        self._dw_backward(x, dw_w, dw_g, dA_get, stride, dil, pad4, in_act, slot=slot)
                    elif k == 1 and stride != 1:
                        tgt, addend2 = self._grad_target(x)
                        self.bwd_seq(lambda: ops.subsample_bwd(dA_get().view(N, Ho, Wo, Cin), x.shape, stride,
                                                               addend=addend2, out=tgt))
                    elif k == 3:
                        tgt, addend2 = self._grad_target(x)
                        self.bwd_seq(lambda: ops.col2im3x3(dA_get(), x.shape, stride, dil[0], Ho, Wo, pt, pl, Kp,
                                                           addend=addend2, out=tgt))
                elif is_sep:
                    # depthwise filter gradient still needs d(dw out) even if the input itself needs no gradient
                    dA_get = self._reserve(f"dA{skey}", (Mo, lda), self.dt)
                    if self.keep_scratch:
                        self._trace_grad(f"{lay.name}/dw", dA_get, x.clog, (N, Ho, Wo, Cin))
                    if self.bf16:
                        self.bwd_seq(lambda: ops.gemm_bf16(dy_get(), wn, Mo, Kdim, ld_dy, dA_get(), lda=ld_dy, ldb=Np,
                                                           ldc=lda))
                    else:
                        self.bwd_seq(lambda: ops.gemm_simt(dy_get(), ld_dy, 1, w32, 1, Cout, dA_get(), lda, Mo, Kdim,
                                                           Cout))
                    self._dw_backward(x, dw_w, dw_g, dA_get, stride, dil, pad4, in_act, need_dx=False, slot=slot)
            else:
                self._dw_backward(x, dw_w, dw_g, dy_get, stride, dil, pad4, in_act, need_dx=needs_in_grad, slot=slot)
            # every parameter gradient of this macro-op is final once the launches appended so far have run
            done = [(id(lay), n) for n in lay.weight_names() if lay._trainable[n]]
            if bn_node is not None:
                done += [(id(bn_node.layer), n) for n in bn_node.layer.weight_names() if bn_node.layer._trainable[n]]
            self._grad_done.append((len(self.bwd), done))

        self._defer_backward(sched)

    def _dw_backward(self, x: Value, dw_w, dw_g, dd_get, stride, dil, pad4, in_act, need_dx=True, slot=None):
        N = self.N
        Ho, Wo = pad4[0], pad4[1]
        Cin = x.C
        xb = x.buf
        in_sc, in_sh = getattr(x, "pre_scale", None), getattr(x, "pre_shift", None)
        if (need_dx and (self.bf16 or FORCE_BNRED) and FUSE_DW_BWD and stride == 1 and tuple(dil) == (1, 1)
                and (Ho, Wo, pad4[2], pad4[3]) == (x.shape[1], x.shape[2], 1, 1) and Cin % 4 == 0
                and (in_act != ACT_NONE or in_sc is None)):
            # input gradient + filter gradient (+ the BN-backward reductions of the layer that produced x) in ONE pass
            # over d(dw out): both need the same 3x3 window of it around every input pixel
            tgt, addend = self._grad_target(x)
            red = {}
            if isinstance(x, _BnActValue) and addend is None:
                # x = relu(BN(y)) never written: the masked gradient and the reductions of that BN come out together
                x.red_done = True
                red = dict(bn_mean=x.bn_mean, bn_invstd=x.bn_invstd, bn_red=x.bn_red)
            elif in_act == ACT_RELU and in_sc is None and FOLD_BLOCK_RED:
                # x = a materialised BN output (+ residual) read through a pre-activation (the next block's first
                # SeparableConv2D): if this launch turns out to write the FINAL gradient of x, the producer's backward
                # (_emit_conv_backward) fills `red` with its raw output y / mean / invstd / slot and skips its own
                # dlv3p_bn_bwd_reduce pass over the two tensors
                x.fold_hook = red

            def launch():
                slot = red.get("bn_red")
                ops.dwconv3x3_bwd(dd_get().view(N, Ho, Wo, Cin), xb, dw_w, dw_g, in_scale=in_sc, in_shift=in_sh,
                                  in_act=in_act, addend=addend, bn_mean=red.get("bn_mean"),
                                  bn_invstd=red.get("bn_invstd"), bn_red=slot() if slot is not None else None,
                                  bn_y=red.get("bn_y"), out=tgt)
            self.bwd_seq(launch)
            return
        self.bwd_seq(lambda: ops.dwconv3x3_wgrad(xb, dd_get().view(N, Ho, Wo, Cin), dw_g, stride, dil, in_scale=in_sc,
                                                 in_shift=in_sh, in_act=in_act, pad=pad4), side=True, slot=slot)
        if need_dx:
            tgt, addend = self._grad_target(x)
            if (isinstance(x, _BnActValue) and addend is None and stride == 1 and tuple(dil) == (1, 1)
                    and (self.bf16 or FORCE_BNRED)):
                # input gradient + the BN-backward reductions of the layer that produced x, in one launch
                x.red_done = True
                self.bwd_seq(lambda: ops.dwconv3x3_dgrad_bnred(dd_get().view(N, Ho, Wo, Cin), dw_w, x.shape, xb, in_sc,
                                                               in_sh, in_act, x.bn_mean, x.bn_invstd, x.bn_red(),
                                                               out=tgt, pad=pad4))
            else:
                self.bwd_seq(lambda: ops.dwconv3x3_dgrad(dd_get().view(N, Ho, Wo, Cin), dw_w, x.shape, stride, dil,
                                                         x_pre=xb if in_act != ACT_NONE else None, in_scale=in_sc,
                                                         in_shift=in_sh, in_act=in_act, addend=addend, out=tgt,
                                                         pad=pad4))

    # The backward schedule is generated in REVERSE topological order (so that _final_grad sees every consumer's
    # contribution): forward emission records one scheduling thunk per macro-op, finalize() runs them backwards
    # and each thunk appends its kernel launches (in execution order) through bwd_seq.
    def _defer_backward(self, sched: Callable[[], None]):
        self._bwd_thunks.append(sched)

    def bwd_seq(self, fn: Callable[[], None], side: bool = False, slot: Optional[int] = None):
        """Append one backward launch.  `side`: a filter-gradient kernel nothing downstream in backward depends on —
        it may run on the side stream, concurrently with the input-gradient chain, filling the SMs the ~20-30 us
        kernels of the middle flow leave idle in their ramp and tail.  `slot`: the scratch slot (0/1) the launch READS
        (side) or whose macro-op is about to WRITE (marker with fn=None): the two scratch copies alternate between
        consecutive macro-ops, so the main stream only has to wait for the side kernels of the macro before last."""
        self.bwd.append((fn, side, slot))
        if fn is not None:
            self.launches_bwd += 1

    def final_prefixes(self, cut: int) -> Tuple[int, int]:
        """Arena offsets (a, b) such that g[0:a] (L2-regularised region) and g[n_reg:b] are FINAL once the backward
        launches [0, cut) have run — what the data-parallel exchange may all-reduce behind backward."""
        P = self.params
        done = set()
        for idx, entries in self._grad_done:
            if idx <= cut:
                done.update(entries)
        a, b = 0, P.n_reg
        for key in P.order_w:
            _, off, shape = P.entries[key]
            size = _ceil8(int(np.prod(shape)))
            if off < P.n_reg:
                if key in done and off == a:
                    a = off + size
            else:
                if key in done and off == b:
                    b = off + size
        return a, b

    def run_bwd_range(self, a: int, b: int, join: bool = True, pending=None):
        """Launch backward[a:b].  `join=False` leaves the side stream un-joined at the end and returns its open events
        (pass them back as `pending` to the call that continues the range): the data-parallel step captures all its
        segments in ONE graph and only marks the cuts with external events (Trainer)."""
        side = self.side_stream
        if side is None:
            for fn, _, _ in self.bwd[a:b]:
                if fn is not None:
                    fn()
            return
        main = torch.cuda.current_stream()
        pending = {} if pending is None else pending
        for fn, on_side, slot in self.bwd[a:b]:
            if fn is None:                                   # macro-op boundary: it is about to overwrite scratch `slot`
                ev = pending.pop(slot, None)
                if ev is not None:
                    main.wait_event(ev)
                continue
            if on_side:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    fn()
                    done = torch.cuda.Event()
                    done.record(side)
                pending[slot if slot is not None else -1] = done
            else:
                fn()
        if not join:
            return pending
        for ev in pending.values():                          # join: the gradient arena is complete after this range
            main.wait_event(ev)
        return {}

    # ---- other ops ---------------------------------------------------------------------------------------
    def _emit_maxpool(self, m: dict):
        n = m["node"]
        x = self.values[n.inputs[0]]
        fused_bn = isinstance(x, _BnPoolValue)
        if not fused_bn:
            x = self._input_of(n.inputs[0], False)
        N, H, W, C = x.shape
        Ho, Wo = -(-H // 2), -(-W // 2)
        out = Value((N, Ho, Wo, C), self.dt, self._alloc((N, Ho, Wo, C), self.dt), n.layer.name)
        out.clog = x.clog
        other = self.values[m["other"]] if m["other"] is not None else None
        am = self._alloc((N, Ho, Wo, C), torch.uint8) if self.training else None
        self.values[m["out"]] = out
        self._tname[m["out"]] = n.layer.name
        self._trace_value(f"{n.layer.name}/out", out.buf, out.clog)
        if am is not None:
            self.pool_sites[n.layer.name] = (am, out.clog)
        out.needs_grad = self.training
        addend = other.buf if other is not None else None
        if fused_bn:
            ymax = self._alloc((N, Ho, Wo, C), self.dt)        # raw conv output of every window's winner
            self.fwd.append(lambda: ops.maxpool3x3s2_bn_fwd(x.buf, x.bn_scale, x.bn_shift, out.buf, ymax, am,
                                                            addend=addend))
        else:
            self.fwd.append(lambda: ops.maxpool3x3s2_fwd(x.buf, out=out.buf, argmax=am, addend=addend))
        self.launches_fwd += 1
        if not self.training:
            return

        def sched():
            g = self._final_grad(out)
            self._trace_grad(f"{n.layer.name}/out", g, out.clog)
            if other is not None:
                other.pending.append(g)
            if fused_bn:
                x.pool_bwd = dict(g=g, argmax=am, ymax=ymax)   # consumed by the producing conv's backward
                return
            if x.needs_grad:
                tgt, add2 = self._grad_target(x)
                self.bwd_seq(lambda: ops.maxpool3x3s2_bwd(g, am, x.shape, addend=add2, out=tgt))
        self._defer_backward(sched)

    def _emit_act(self, n: FlatNode, code: int, consumers):
        x = self._input_of(n.inputs[0], False)
        cons = consumers.get(n.output, [])
        fusable = cons and all(isinstance(c.layer, (L.SeparableConv2D, L.DepthwiseConv2D)) for c in cons)
        site = self._tname.get(n.inputs[0], n.layer.name)
        self.act_sites[site] = ("x", code, x)
        if fusable:
            # virtual pre-activation: consumers apply it on load and mask it in their input gradient
            self.values[n.output] = _AliasValue(x, code)
            return
        out = Value(x.shape, self.dt, self._alloc(x.shape, self.dt), n.layer.name)
        out.clog = x.clog
        out.needs_grad = self.training and x.needs_grad
        self.values[n.output] = out
        self.fwd.append(lambda: ops.affine_act(x.buf, x.M, x.C, out.buf, None, None, code))
        self.launches_fwd += 1
        if self.training and x.needs_grad:
            def sched():
                g = self._final_grad(out)
                tgt, add2 = self._grad_target(x)
                self.bwd_seq(lambda: ops.act_bwd(g, x.buf, code, tgt, addend=add2))
            self._defer_backward(sched)

    def _emit_add(self, n: FlatNode):
        a, b = self._input_of(n.inputs[0], False), self._input_of(n.inputs[1], False)
        out = Value(a.shape, self.dt, self._alloc(a.shape, self.dt), n.layer.name)
        out.clog = a.clog
        out.needs_grad = self.training
        self.values[n.output] = out
        self.fwd.append(lambda: ops.add(a.buf, b.buf, out.buf))
        self.launches_fwd += 1
        if self.training:
            def sched():
                g = self._final_grad(out)
                a.pending.append(g)
                b.pending.append(g)
            self._defer_backward(sched)

    def _emit_avgpool(self, n: FlatNode):
        x = self._input_of(n.inputs[0], False)
        k = n.layer.pool_size[0]
        if k == 1:
            self.values[n.output] = x           # identity pool in every shipped config (conf.json:51)
            return
        N, H, W, C = x.shape
        out = Value((N, H // k, W // k, C), self.dt, self._alloc((N, H // k, W // k, C), self.dt), n.layer.name)
        out.clog = x.clog
        out.needs_grad = self.training
        self.values[n.output] = out
        self.fwd.append(lambda: ops.avgpool_fwd(x.buf, k, out=out.buf))
        self.launches_fwd += 1
        self._trace_value(f"{n.layer.name}/out", out.buf, out.clog)
        if self.training and x.needs_grad:
            def sched():
                g = self._final_grad(out)
                self._trace_grad(f"{n.layer.name}/out", g, out.clog)
                tgt, add2 = self._grad_target(x)
                self.bwd_seq(lambda: ops.avgpool_bwd(g, x.shape, k, addend=add2, out=tgt))
            self._defer_backward(sched)

    def _emit_resize(self, n: FlatNode, out_id: int, consumers):
        x = self._input_of(n.inputs[0], False)
        fh, fw = n.layer.factors
        if (fh, fw) == (1, 1):
            self.values[n.output] = x
            return
        N, H, W, C = x.shape
        cons = consumers.get(n.output, [])
        tail = (len(cons) == 1 and isinstance(cons[0].layer, L.Activation) and cons[0].layer.activation == "softmax"
                and cons[0].output == out_id)
        if tail:
            # decoder tail: handled by _emit_softmax_tail (fused with the loss in training)
            self.values[n.output] = _TailResize(x, fh, fw)
            return
        oshape = (N, H * fh, W * fw, C)
        slot_c = self._concat_slot.get(n.output) if (x.clog == C and x.dtype == self.dt and C % 8 == 0) else None
        if slot_c is not None and slot_c[1] % 8 == 0:
            # the resize's only reader is a Concatenate (boundary refinement, ss.py:941-950: the x(OS/2) encoder output and
            # low-level features, 304 channels at 256^2): written straight into its channel slice (ld_y), gradient read
            # from the same slice of the concat gradient (ld_dy) — no slice copies of the largest tensors of the model
            cid, c_off, Ct = slot_c
            if cid not in self._concat_buf:
                self._concat_buf[cid] = self._alloc((N, H * fh, W * fw, Ct), self.dt)
            cbuf = self._concat_buf[cid]
            out = Value(oshape, x.dtype, cbuf.view(N * H * fh * W * fw, Ct)[:, c_off:c_off + C], n.layer.name)
            out.concat_slice = True
            out.clog = x.clog
            out.needs_grad = self.training
            self.values[n.output] = out
            self.fwd.append(lambda: ops.bilinear_fwd(x.buf, fh, fw, out=cbuf, ld_y=Ct, C=C, y_off=c_off))
            self.launches_fwd += 1
            self._trace_value(f"{n.layer.name}/out", out.buf, out.clog, oshape)
            if self.training and x.needs_grad:
                def sched():
                    g = self._final_grad(out)                # the [M, C] slice view of the concat gradient
                    self._trace_grad(f"{n.layer.name}/out", g, out.clog, oshape)
                    ld_g = g.stride(0) if g.dim() == 2 else C
                    tgt, add2 = self._grad_target(x)
                    self.bwd_seq(lambda: ops.bilinear_bwd(g, x.shape, fh, fw, out=tgt, addend=add2, ld_dy=ld_g))
                self._defer_backward(sched)
            return
        out = Value(oshape, x.dtype, self._alloc(oshape, x.dtype), n.layer.name)
        out.clog = x.clog
        out.needs_grad = self.training
        self.values[n.output] = out
        self.fwd.append(lambda: ops.bilinear_fwd(x.buf, fh, fw, out=out.buf))
        self.launches_fwd += 1
        self._trace_value(f"{n.layer.name}/out", out.buf, out.clog)
        if self.training and x.needs_grad:
            def sched():
                g = self._final_grad(out)
                self._trace_grad(f"{n.layer.name}/out", g, out.clog)
                tgt, add2 = self._grad_target(x)
                self.bwd_seq(lambda: ops.bilinear_bwd(g, x.shape, fh, fw, out=tgt, addend=add2))
            self._defer_backward(sched)

    def _emit_concat(self, n: FlatNode):
        ins = [self._input_of(i, False) for i in n.inputs]
        if any(v.clog != v.C for v in ins):
            raise NotImplementedError("Concatenate of a channel-padded tensor (phys_channels) is not on the hot path")
        N, H, W, _ = ins[0].shape
        Ct = sum(v.C for v in ins)
        buf = self._concat_buf.get(n.output)
        if buf is None:
            buf = self._alloc((N, H, W, Ct), self.dt)
        out = Value((N, H, W, Ct), self.dt, buf, n.layer.name)
        out.needs_grad = self.training
        self.values[n.output] = out
        self._trace_value(f"{n.layer.name}/out", buf, Ct)
        M = N * H * W
        off = 0
        for v in ins:
            if not getattr(v, "concat_slice", False):          # (slices written in place by their producer)
                self.fwd.append(lambda v=v, off=off: ops.copy2d(v.buf, v.C, out.buf, Ct, M, v.C, y_off=off))
                self.launches_fwd += 1
            off += v.C
        if self.training:
            def sched():
                g = self._final_grad(out)
                self._trace_grad(f"{n.layer.name}/out", g, Ct)
                o = 0
                for v in ins:
                    if v.needs_grad and getattr(v, "concat_slice", False):
                        v.pending.append(g.view(M, Ct)[:, o:o + v.C])      # read in place (ld_dz = Ct), no copy
                    elif v.needs_grad:
                        tgt, add2 = self._grad_target(v)
                        self.bwd_seq(lambda v=v, o=o, tgt=tgt, add2=add2: ops.copy2d(
                            g, Ct, tgt, v.C, M, v.C, addend=add2, ld_addend=v.C, x_off=o))
                    o += v.C
            self._defer_backward(sched)

    def _emit_dropout(self, n: FlatNode):
        x = self._input_of(n.inputs[0], False)
        rate = n.layer.rate
        if not self.training or rate == 0.0:
            self.values[n.output] = x
            return
        out = Value(x.shape, self.dt, self._alloc(x.shape, self.dt), n.layer.name)
        out.clog = x.clog
        out.needs_grad = True
        self.values[n.output] = out
        seed = self.dropout_seed + 7919 * len(self.fwd)
        ctr = self.step_counter
        self.fwd.append(lambda: ops.dropout(x.buf, rate, seed, out.buf, seed_offset=ctr))
        self.launches_fwd += 1
        self._trace_value(f"{n.layer.name}/out", out.buf, out.clog)
        self.dropout_sites = getattr(self, "dropout_sites", {})
        self.dropout_sites[n.layer.name] = (out, rate, seed)

        def sched():
            g = self._final_grad(out)
            self._trace_grad(f"{n.layer.name}/out", g, out.clog)
            tgt, add2 = self._grad_target(x)
            self.bwd_seq(lambda: ops.dropout(g, rate, seed, tgt, addend=add2, seed_offset=ctr))
        self._defer_backward(sched)

    def _emit_softmax_tail(self, n: FlatNode, producer):
        src = self.values[n.inputs[0]]
        if isinstance(src, _TailResize):
            logits, f = src.x, src.fh
            if src.fh != src.fw:
                raise NotImplementedError("decoder tail: isotropic resize factor expected")
        else:
            logits, f = src, 1
        if logits.dtype != torch.float32:
            raise NotImplementedError("decoder tail expects fp32 logits: the class-score convolution must not be followed by BatchNormalization (ss.py:893-897)")
        self.logits, self.tail_factor = logits, f
        N, H, W, C = logits.shape
        if C > 32:
            raise NotImplementedError("softmax tail supports up to 32 classes (one warp lane per class; VOC 21 / Cityscapes 19)")
        self.out_shape = (N, H * f, W * f, C)
        self.labels = self._alloc((N, H * f, W * f), torch.int32, zero=True)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.reg_sum = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._tail_bufs: Dict[str, torch.Tensor] = {}
        if self.training:
            logits.needs_grad = True
            self._grad_of(logits)
            logits.grad_written = True

    def tail_buf(self, name, shape, dtype):
        if name not in self._tail_bufs:
            self._tail_bufs[name] = self._alloc(shape, dtype)
        return self._tail_bufs[name]

    # ---------------------------------------------------------------------------------------------- running
    def upload_weights(self):
        """Host arrays of the layers -> device (after writing into Model.named_weights() arrays in place)."""
        self.params.upload()
        self.ensure_current()

    def ensure_current(self):
        """Derived operands (bf16 GEMM weights, folded inference BN) follow the shared store: re-derive them when
        another plan's optimizer step, a set_weights or a checkpoint load changed the weights since this plan last
        ran.  A pure host-side check when nothing changed (safe inside CUDA-graph capture after the warm-up)."""
        P = self.params
        P.sync_device()
        if self._prep_version != P.version:
            self.run_prep()
            self._prep_version = P.version

    def run_prep(self):
        if self._wprep:
            if self._wprep_table is None:
                self._wprep_table = ops.weight_prep_table(self._wprep, self.device)
            ops.weight_prep_batch(self._wprep_table, len(self._wprep))
        for fn in self.prep:
            fn()

    def finalize(self):
        """Generate the backward schedule (once)."""
        if getattr(self, "_finalized", False):
            return
        self._finalized = True
        if self.training:
            self.bwd = []
            for t in reversed(self._bwd_thunks):
                t()

    def set_loss(self, pos_weights, neg_weights, epsilon=1e-7):
        C = self.logits.C
        if len(pos_weights) != C or len(neg_weights) != C:
            raise ValueError(f"loss weights have {len(pos_weights)} entries for {C} classes (ss.py:443)")
        self.pw = torch.tensor(pos_weights, dtype=torch.float32, device=self.device)
        self.nw = torch.tensor(neg_weights, dtype=torch.float32, device=self.device)
        self.eps = float(epsilon)

    def zero_grads(self):
        self.params.g.zero_()
        self.stats.zero_()
        self.loss_sum.zero_()
        self.reg_sum.zero_()

    def forward(self):
        self.ensure_current()
        for fn in self.fwd:
            fn()

    def loss_forward_backward(self):
        """Fused decoder tail on the low-resolution logits: loss value + gradient w.r.t. the logits."""
        N, H, W, C = self.logits.shape
        f = self.tail_factor
        P = N * H * f * W * f
        z = self.logits.buf
        if self.fused_tail and f > 1:
            g = self.logits.grad
            g.zero_()
            ops.upsample_softmax_cbloss_fwd_bwd(z, self.labels, self.pw, self.nw, self.eps, N, H, W, C, f, 1.0 / P,
                                                self.loss_sum, g)
        else:
            zh = self.tail_buf("zh", (N, H * f, W * f, C), torch.float32)
            dzh = self.tail_buf("dzh", (N, H * f, W * f, C), torch.float32)
            if f > 1:
                ops.bilinear_fwd(z, f, f, out=zh)
            else:
                zh = z
            ops.softmax_cbloss_fwd(zh, self.labels, self.pw, self.nw, self.eps, P, C, self.loss_sum)
            ops.softmax_cbloss_bwd(zh, self.labels, self.pw, self.nw, self.eps, P, C, 1.0 / P, dzh)
            if f > 1:
                ops.bilinear_bwd(dzh, self.logits.shape, f, f, out=self.logits.grad)
            else:
                self.logits.grad.copy_(dzh)

    def backward(self):
        self.run_bwd_range(0, len(self.bwd))

    def step_fwd_bwd(self):
        """One forward + backward pass over the batch already resident in self.x_in / self.labels."""
        self.finalize()
        self.zero_grads()
        self.forward()
        self.loss_forward_backward()
        self.backward()

    def regularization(self):
        if self.params.l2 and self.params.n_reg:
            ops.sumsq(self.params.w, self.params.n_reg, self.reg_sum)

    def adam_step(self, opt, grad_scale: float = 1.0):
        P = self.params
        lr_t = opt.step_size()
        if P.n_reg:
            ops.adam(P.w, P.g, P.m, P.v, P.n_reg, lr_t, opt.beta_1, opt.beta_2, opt.epsilon, grad_scale, P.l2)
        if P.n_train > P.n_reg:
            ops.adam(P.w, P.g, P.m, P.v, P.n_train - P.n_reg, lr_t, opt.beta_1, opt.beta_2, opt.epsilon, grad_scale,
                     0.0, w_off=P.n_reg)
        opt.iterations += 1
        self.step_counter.add_(1)
        P.mark_updated()
        self.run_prep()
        self._prep_version = P.version

    def loss_value(self) -> float:
        P = self.N * self.out_shape[1] * self.out_shape[2]
        return float(self.loss_sum.item()) / P + self.params.l2 * float(self.reg_sum.item())

    # ---- host-facing API -------------------------------------------------------------------------------
    def load_batch(self, images, labels=None):
        x = torch.as_tensor(images)
        if tuple(x.shape) != self.x_in.shape:
            raise ValueError(f"expected images of shape {self.x_in.shape}, got {tuple(x.shape)}")
        self.x_in.buf.copy_(x.to(self.device, non_blocking=True))
        if labels is not None:
            y = torch.as_tensor(labels)
            if y.dim() == 4:                      # one-hot [B,H,W,C] (reference Sequence output, ss.py:1602)
                y = y.argmax(dim=-1)
            if tuple(y.shape) != tuple(self.labels.shape):
                raise ValueError(f"expected labels of shape {tuple(self.labels.shape)}, got {tuple(y.shape)}")
            self.labels.copy_(y.to(self.device, non_blocking=True))

    def train_on_batch(self, images, labels) -> float:
        m = self.model
        if not hasattr(self, "pw"):
            self.set_loss(m.loss.pos_weights, m.loss.neg_weights, m.loss.epsilon)
        self.load_batch(images, labels)
        self.step_fwd_bwd()
        self.regularization()
        self.adam_step(m.optimizer)
        return self.loss_value()

    def logits_highres(self) -> torch.Tensor:
        N, H, W, C = self.logits.shape
        f = self.tail_factor
        if f == 1:
            return self.logits.buf
        zh = self.tail_buf("zh", (N, H * f, W * f, C), torch.float32)
        ops.bilinear_fwd(self.logits.buf, f, f, out=zh)
        return zh

    def predict_device(self) -> torch.Tensor:
        self.forward()
        zh = self.logits_highres()
        probs = self.tail_buf("probs", self.out_shape, torch.float32)
        ops.softmax_argmax(zh, zh.numel() // zh.shape[-1], zh.shape[-1], probs=probs)
        return probs

    def predict(self, images) -> np.ndarray:
        self.load_batch(images)
        return self.predict_device().cpu().numpy()

    def segment_device(self, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Label maps [N,Ho,Wo] (int32, or the uint8 / int32 tensor passed in) of the batch resident in the input
        buffer: forward, then bilinear up-sampling fused with the channel argmax on the low-resolution logits
        (dlv3p_upsample_argmax) — the high-resolution logits / probabilities are never written."""
        self.forward()
        if labels is None:
            labels = self.tail_buf("lab", self.out_shape[:3], torch.int32)
        f = self.tail_factor
        ops.upsample_argmax(self.logits.buf, f, f, labels)
        return labels

    def segment(self, images) -> np.ndarray:
        self.load_batch(images)
        return self.segment_device().cpu().numpy().astype(np.int64)

    def gradients(self) -> Dict[str, np.ndarray]:
        """{'layer/weight': gradient} after step_fwd_bwd (parity tests)."""
        out = {}
        for l in self.model.flat_layers():
            for n in l.weight_names():
                if l._trainable[n]:
                    out[f"{l.name}/{n}"] = self.params.logical(l, n, grad=True).detach().cpu().numpy().copy()
        return out


class _AliasValue(Value):
    """A tensor that is `act(x)` for a materialised x, never written to memory: consumers fuse the activation."""

    def __init__(self, x: Value, act: int):
        self._x = x
        self.pre_act = act
        self.shape, self.dtype, self.name = x.shape, x.dtype, x.name + "/act"
        self.clog = x.clog

    @property
    def buf(self):
        return self._x.buf

    @property
    def needs_grad(self):
        return self._x.needs_grad

    # gradient writes go to the underlying tensor (the consumer applies the activation mask itself)
    @property
    def grad(self):
        return self._x.grad

    @grad.setter
    def grad(self, v):
        self._x.grad = v

    @property
    def grad_written(self):
        return self._x.grad_written

    @grad_written.setter
    def grad_written(self, v):
        self._x.grad_written = v

    @property
    def pending(self):
        return self._x.pending

    @property
    def fold_hook(self):
        return self._x.fold_hook

    @fold_hook.setter
    def fold_hook(self, v):
        self._x.fold_hook = v


class _BnActValue(Value):
    """`act(scale * y + shift)` for the raw conv output y of a training-mode Conv -> BatchNormalization -> ReLU/ReLU6
    macro-op whose only reader is a dense-tap depthwise stage.  Never written to memory: the reader's forward and
    filter-gradient kernels apply the map on load (dlv3p_dwconv3x3_fwd / _wgrad in_scale/in_shift/in_act), its
    input-gradient kernel masks with act' and writes the gradient w.r.t. the BN OUTPUT into `grad` (and, on the TMA
    path, the two BN-backward reductions into `bn_red`)."""

    def __init__(self, shape, dtype, name, act):
        super().__init__(shape, dtype, None, name + "/bn_act")
        self.pre_act = act
        self.red_done = False
        self.bn_fold = None                 # BN finalize operands when the reader's forward kernel finishes the BN

    def attach(self, y, scale, shift, mean, invstd, red_slot):
        self.buf, self.pre_scale, self.pre_shift = y, scale, shift
        self.bn_mean, self.bn_invstd, self.bn_red = mean, invstd, red_slot


class _BnPoolValue(Value):
    """`scale * y + shift` for the raw conv output y of a training-mode Conv -> BatchNormalization whose only reader is
    a MaxPooling2D(3, strides 2).  Never written: dlv3p_maxpool3x3s2_bn_fwd pools it on the fly (and keeps the raw
    winner values), the pool's backward hands its pooled gradient to the producer (`pool_bwd`), whose backward runs
    the BN reductions on the pooled tensors and dlv3p_maxpool3x3s2_bn_bwd (pool backward + BN input gradient)."""

    def __init__(self, shape, dtype, name):
        super().__init__(shape, dtype, None, name + "/bn_pool")
        self.pre_act = -1                   # no consumer but the fused pool may read it
        self.pool_bwd = None

    def attach(self, y, scale, shift, mean, invstd, red_slot, count):
        self.buf, self.bn_scale, self.bn_shift = y, scale, shift
        self.bn_mean, self.bn_invstd, self.bn_red, self.bn_count = mean, invstd, red_slot, count


FORCE_IMPLICIT = False  # tests: take the implicit-GEMM 3x3 schedule in fp32 too (through tests/fake_ops.py)
FORCE_BNRED = False     # tests: take the fused dgrad+reduction schedule in fp32 too (through tests/fake_ops.py)
FUSE_DW_BWD = os.environ.get("DLV3P_FUSE_DW_BWD", "1") != "0"   # A/B switch of the schedule (host side only)
FOLD_BLOCK_RED = os.environ.get("DLV3P_FOLD_BLOCK_RED", "1") != "0"


class _TailResize:
    def __init__(self, x: Value, fh: int, fw: int):
        self.x, self.fh, self.fw = x, fh, fw
        self.pre_act = ACT_NONE

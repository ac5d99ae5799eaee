"""Per-kernel timing and the algorithmic cost model used for roofline reporting.

`KernelProfiler` brackets every C-ABI call with CUDA events on the launching stream (torch.cuda.current_stream at
call time) and aggregates time, algorithmic bytes and flops per entry point.  The cost model is the "algorithmic
work per unit" of SURVEY.md §8(d): bytes every kernel MUST move (each operand once) and flops it must do, computed
from the call's integer arguments — not from hardware counters, which the ncu captures under profiles/ provide.
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from typing import Callable, Dict, Tuple

import torch

from . import _lib


def _esz(dt: int) -> int:
    return 4 if dt == _lib.F32 else 2


def _opt(p) -> int:
    return 1 if p else 0


# name -> (bound, fn(args) -> (bytes, flops))
COSTS: Dict[str, Tuple[str, Callable]] = {
    "dlv3p_dwconv3x3_fwd": ("hbm", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * a[12] * a[13] * a[6]) * _esz(a[17]) + 36 * a[6],
                                             18 * a[3] * a[12] * a[13] * a[6])),
    "dlv3p_dwconv3x3_fwd_epi": ("hbm", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * a[12] * a[13] * a[6]) * _esz(a[17]) + 44 * a[6],
                                                 20 * a[3] * a[12] * a[13] * a[6])),
    "dlv3p_dwconv3x3_dgrad": ("hbm", lambda a: ((a[3] * a[4] * a[5] * a[6] * (1 + _opt(a[14]) + _opt(a[18])) + a[3] * a[12] * a[13] * a[6]) * _esz(a[19]) + 36 * a[6],
                                               18 * a[3] * a[4] * a[5] * a[6])),
    "dlv3p_dwconv3x3_bn_fwd": ("hbm", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * a[9] * a[10] * a[6]) * _esz(a[25]) + 36 * a[6],
                                                20 * a[3] * a[9] * a[10] * a[6])),
    "dlv3p_dwconv3x3_dgrad_bnred": ("hbm", lambda a: ((2 * a[3] * a[4] * a[5] * a[6] + a[3] * a[9] * a[10] * a[6]) * _esz(a[18]) + 36 * a[6],
                                                     22 * a[3] * a[4] * a[5] * a[6])),
    "dlv3p_dwconv3x3_bwd": ("hbm", lambda a: (a[5] * a[6] * a[7] * a[8] * (3 + _opt(a[12]) + _opt(a[16])) * _esz(a[17]) + 72 * a[8],
                                             (36 + 4 * _opt(a[15])) * a[5] * a[6] * a[7] * a[8])),
    "dlv3p_dwconv3x3_wgrad": ("hbm", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * a[12] * a[13] * a[6]) * _esz(a[17]) + 36 * a[6],
                                               18 * a[3] * a[12] * a[13] * a[6])),
    "dlv3p_gemm_bf16": ("tensor", lambda a: ((a[6] * a[8] + a[7] * a[8]) * 2 + a[6] * a[7] * _esz(a[9]) * (1 + _opt(a[13])),
                                            2 * a[6] * a[7] * a[8])),
    "dlv3p_gemm_wgrad_bf16": ("tensor", lambda a: ((a[6] * a[7] + a[6] * a[8]) * 2 + a[7] * a[8] * 4, 2 * a[6] * a[7] * a[8])),
    "dlv3p_conv3x3_valid_fwd_bf16": ("tensor", lambda a: ((a[4] * a[5] * a[6] * a[7] + a[4] * (a[5] - 2) * (a[6] - 2) * a[8]) * 2 + 18 * a[7] * a[8],
                                                         18 * a[4] * (a[5] - 2) * (a[6] - 2) * a[7] * a[8])),
    "dlv3p_conv3x3_valid_dgrad_bf16": ("tensor", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * (a[4] - 2) * (a[5] - 2) * a[7]) * 2 + 18 * a[6] * a[7],
                                                           18 * a[3] * (a[4] - 2) * (a[5] - 2) * a[6] * a[7])),
    "dlv3p_conv3x3_valid_wgrad_bf16": ("tensor", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * (a[4] - 2) * (a[5] - 2) * a[7]) * 2 + 36 * a[6] * a[7],
                                                           18 * a[3] * (a[4] - 2) * (a[5] - 2) * a[6] * a[7])),
    # SAME 3x3: x once + y once (+ filter); args (x, wt, ldw, y, c_dtype, N, H, W, Cin, Cout, ...)
    "dlv3p_conv3x3_same_fwd_bf16": ("tensor", lambda a: (a[5] * a[6] * a[7] * (a[8] * 2 + a[9] * _esz(a[4])) + 18 * a[8] * a[9],
                                                        18 * a[5] * a[6] * a[7] * a[8] * a[9])),
    # (dy, ld_dy, wd, dx, N, H, W, Cin, Cout)
    "dlv3p_conv3x3_same_dgrad_bf16": ("tensor", lambda a: (a[4] * a[5] * a[6] * (a[7] + a[8]) * 2 + 18 * a[7] * a[8],
                                                          18 * a[4] * a[5] * a[6] * a[7] * a[8])),
    # (x, dy, ld_dy, dw, N, H, W, Cin, Cout)
    "dlv3p_conv3x3_same_wgrad_bf16": ("tensor", lambda a: (a[4] * a[5] * a[6] * (a[7] + a[8]) * 2 + 36 * a[7] * a[8],
                                                          18 * a[4] * a[5] * a[6] * a[7] * a[8])),
    "dlv3p_gemm_simt": ("fp32", lambda a: ((a[8] * a[10] + a[10] * a[9]) * _esz(a[11]) + a[8] * a[9] * _esz(a[12]),
                                          2 * a[8] * a[9] * a[10])),
    "dlv3p_im2col3x3": ("hbm", lambda a: ((a[2] * a[3] * a[4] * a[5] + a[2] * a[10] * a[11] * a[12]) * _esz(a[13]), 0)),
    "dlv3p_col2im3x3": ("hbm", lambda a: ((a[2] * a[3] * a[4] * a[5] * (1 + _opt(a[13])) + a[2] * a[10] * a[11] * 9 * a[5]) * _esz(a[14]), 0)),
    "dlv3p_subsample_fwd": ("hbm", lambda a: (2 * a[2] * a[7] * a[8] * a[5] * _esz(a[9]), 0)),
    "dlv3p_subsample_bwd": ("hbm", lambda a: ((a[2] * a[7] * a[8] * a[5] + a[2] * a[3] * a[4] * a[5] * (1 + _opt(a[9]))) * _esz(a[10]), 0)),
    "dlv3p_weight_prep": ("hbm", lambda a: (a[1] * a[2] * (4 + 2 + (2 if a[5] else 0)), 0)),
    "dlv3p_bn_train_apply": ("hbm", lambda a: ((2 + _opt(a[12])) * a[16] * a[17] * _esz(a[22]), 2 * a[16] * a[17])),
    "dlv3p_weight_prep_batch": ("hbm", lambda a: (0, 0)),
    "dlv3p_bn_stats": ("hbm", lambda a: (a[2] * a[3] * _esz(a[5]), 3 * a[2] * a[3])),
    "dlv3p_affine_act": ("hbm", lambda a: ((2 + _opt(a[5])) * a[9] * a[10] * _esz(a[11]), 2 * a[9] * a[10])),
    "dlv3p_bn_bwd_reduce": ("hbm", lambda a: (2 * a[9] * a[10] * _esz(a[12]), 6 * a[9] * a[10])),
    "dlv3p_bn_bwd_apply": ("hbm", lambda a: (3 * a[10] * a[11] * _esz(a[14]), 8 * a[10] * a[11])),
    "dlv3p_act_bwd": ("hbm", lambda a: ((3 + _opt(a[4])) * a[5] * _esz(a[6]), a[5])),
    "dlv3p_add": ("hbm", lambda a: (3 * a[3] * _esz(a[4]), a[3])),
    "dlv3p_copy2d": ("hbm", lambda a: ((2 + _opt(a[6])) * a[4] * a[5] * _esz(a[8]), 0)),
    "dlv3p_maxpool3x3s2_fwd": ("hbm", lambda a: ((a[3] * a[4] * a[5] * a[6] + a[3] * a[9] * a[10] * a[6] * (1 + _opt(a[11]))) * _esz(a[12]) + a[3] * a[9] * a[10] * a[6] * _opt(a[2]), 0)),
    "dlv3p_maxpool3x3s2_bn_fwd": ("hbm", lambda a: ((a[6] * a[7] * a[8] * a[9] + a[6] * a[12] * a[13] * a[9] * (2 + _opt(a[14]))) * _esz(a[15]) + a[6] * a[12] * a[13] * a[9], 0)),
    "dlv3p_maxpool3x3s2_bn_bwd": ("hbm", lambda a: ((a[9] * a[15] * a[16] * a[12] + 2 * a[9] * a[10] * a[11] * a[12]) * _esz(a[17]) + a[9] * a[15] * a[16] * a[12], 0)),
    "dlv3p_maxpool3x3s2_bwd": ("hbm", lambda a: ((a[3] * a[9] * a[10] * a[6] + a[3] * a[4] * a[5] * a[6] * (1 + _opt(a[11]))) * _esz(a[12]) + a[3] * a[9] * a[10] * a[6], 0)),
    "dlv3p_bilinear_fwd": ("hbm", lambda a: (a[4] * a[5] * a[6] * a[7] * (_esz(a[10]) + a[8] * a[9] * _esz(a[11])), 0)),
    "dlv3p_bilinear_bwd": ("hbm", lambda a: (a[4] * a[5] * a[6] * a[7] * (a[8] * a[9] * _esz(a[11]) + _esz(a[12])), 0)),
    "dlv3p_upsample_softmax_cbloss_fwd": ("hbm", lambda a: (a[5] * a[6] * a[7] * (a[9] * a[9] * 4 + a[8] * 4), 0)),
    "dlv3p_upsample_softmax_cbloss_bwd": ("hbm", lambda a: (a[5] * a[6] * a[7] * (a[9] * a[9] * 4 + 2 * a[8] * 4), 0)),
    "dlv3p_upsample_softmax_cbloss_fwd_bwd": ("hbm", lambda a: (a[5] * a[6] * a[7] * (a[9] * a[9] * 4 + 2 * a[8] * 4), 0)),
    "dlv3p_softmax_cbloss_fwd": ("hbm", lambda a: (a[5] * (a[6] * 4 * (1 + _opt(a[8])) + 4), 0)),
    "dlv3p_softmax_cbloss_bwd": ("hbm", lambda a: (a[5] * (2 * a[6] * 4 + 4), 0)),
    "dlv3p_softmax_argmax": ("hbm", lambda a: (a[1] * a[2] * 4 * (1 + _opt(a[3])) + a[1] * 4 * _opt(a[4]), 0)),
    "dlv3p_adam": ("hbm", lambda a: (7 * a[4] * 4, 0)),
    "dlv3p_sumsq": ("hbm", lambda a: (a[1] * 4, 0)),
    "dlv3p_cast": ("hbm", lambda a: (a[4] * (_esz(a[1]) + _esz(a[3])), 0)),
    "dlv3p_cast2d": ("hbm", lambda a: (a[6] * a[7] * (_esz(a[2]) + _esz(a[5])), 0)),
    "dlv3p_dropout": ("hbm", lambda a: ((2 + _opt(a[6])) * a[2] * _esz(a[7]), 0)),
    "dlv3p_bn_finalize": ("hbm", lambda a: (a[5] * 4 * 12, 0)),
    "dlv3p_bn_fold": ("hbm", lambda a: (a[4] * 4 * 6, 0)),
}


def cost(name: str, args) -> Tuple[int, int]:
    ent = COSTS.get(name)
    if ent is None:
        return 0, 0
    return ent[1](args)


class KernelProfiler:
    """with KernelProfiler() as kp: ...; kp.summary() -> {entry point: dict(calls, ms, bytes, flops)}."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        _lib.PROFILER = self
        return self

    def __exit__(self, *exc):
        _lib.PROFILER = None

    def before(self, name, args):
        s = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        return (name, args, e0, s)

    def after(self, tok):
        name, args, e0, s = tok
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(s)
        self.records.append((name, args, e0, e1))

    def summary(self) -> "OrderedDict[str, dict]":
        torch.cuda.synchronize()
        agg = defaultdict(lambda: dict(calls=0, ms=0.0, bytes=0, flops=0))
        for name, args, e0, e1 in self.records:
            b, f = cost(name, args)
            a = agg[name]
            a["calls"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["bytes"] += b
            a["flops"] += f
        out = OrderedDict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))
        for name, a in out.items():
            a["bound"] = COSTS.get(name, ("hbm",))[0]
            sec = max(a["ms"], 1e-9) / 1e3
            a["GBps"] = a["bytes"] / sec / 1e9
            a["TFLOPs"] = a["flops"] / sec / 1e12
        return out

    def detail(self, top: int = 40):
        """Per (entry point, shape) breakdown: [(name, small-int args, calls, ms, GB/s, TFLOP/s)] sorted by time."""
        torch.cuda.synchronize()
        agg = defaultdict(lambda: [0, 0.0, 0, 0])
        for name, args, e0, e1 in self.records:
            sig = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and 0 <= a < (1 << 24))
            b, f = cost(name, args)
            a = agg[(name, sig)]
            a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += b; a[3] += f
        rows = []
        for (name, sig), (calls, ms, b, f) in agg.items():
            sec = max(ms, 1e-9) / 1e3
            rows.append(dict(kernel=name, args=list(sig), calls=calls, ms=ms, GBps=b / sec / 1e9,
                             TFLOPs=f / sec / 1e12))
        rows.sort(key=lambda r: -r["ms"])
        return rows[:top]

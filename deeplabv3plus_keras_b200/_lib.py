"""ctypes binding of libdlv3p.so — the C-ABI declared in include/dlv3p.h.

This is the only route from the Python host layer to arithmetic: there is no CPU fallback.  If the shared
library (built by ``deeplabv3plus_keras_b200/csrc/build.sh`` / ``__graft_entry__.build()``) is missing, or a
call returns a negative status, a RuntimeError / ValueError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DLV3P_LIB overrides the library (diagnostic scripts point it at libdlv3p_diag.so, built with DLV3P_DIAG=1)
LIB_PATH = os.environ.get("DLV3P_LIB") or os.path.join(_HERE, "libdlv3p.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
ERR_SHAPE, ERR_DTYPE, ERR_ALIGN, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3, -4, -5

_p, _i, _l, _f, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64

# name -> argtypes, mirrors include/dlv3p.h one to one (tests/test_abi.py checks the header against this table)
SIGNATURES = {
    "dlv3p_dwconv3x3_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p],
    "dlv3p_dwconv3x3_fwd_epi": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p],
    "dlv3p_dwconv3x3_dgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p],
    "dlv3p_dwconv3x3_bn_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _d, _f, _f, _i, _i, _p, _p,
                               _p, _p, _i, _p],
    "dlv3p_dwconv3x3_dgrad_bnred": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _i, _p],
    "dlv3p_dwconv3x3_wgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p],
    "dlv3p_dwconv3x3_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _i, _p],
    "dlv3p_gemm_bf16": [_p, _l, _p, _l, _p, _l, _i, _i, _i, _i, _p, _p, _i, _p, _l, _p, _p],
    "dlv3p_gemm_wgrad_bf16": [_p, _l, _p, _l, _p, _l, _i, _i, _i, _p],
    "dlv3p_conv3x3_valid_fwd_bf16": [_p, _p, _l, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p],
    "dlv3p_conv3x3_valid_dgrad_bf16": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dlv3p_conv3x3_valid_wgrad_bf16": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dlv3p_conv3x3_same_fwd_bf16": [_p, _p, _l, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p],
    "dlv3p_conv3x3_same_dgrad_bf16": [_p, _l, _p, _p, _i, _i, _i, _i, _i, _p],
    "dlv3p_conv3x3_same_wgrad_bf16": [_p, _p, _l, _p, _i, _i, _i, _i, _i, _p],
    "dlv3p_gemm_simt": [_p, _l, _l, _p, _l, _l, _p, _l, _i, _i, _i, _i, _i, _p, _p, _i, _p, _l, _i, _p],
    "dlv3p_im2col3x3": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _l, _i, _p],
    "dlv3p_col2im3x3": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _l, _p, _i, _p],
    "dlv3p_subsample_fwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dlv3p_subsample_bwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "dlv3p_weight_prep": [_p, _i, _i, _p, _l, _p, _l, _p],
    "dlv3p_weight_prep_batch": [_p, _i, _i, _p],
    "dlv3p_bn_train_apply": [_p, _l, _p, _p, _p, _p, _p, _d, _f, _f, _i, _i, _p, _l, _p, _l, _l, _i, _p, _p, _p, _p,
                             _i, _p],
    "dlv3p_bn_stats": [_p, _l, _l, _i, _p, _i, _p],
    "dlv3p_bn_finalize": [_p, _p, _p, _p, _p, _i, _d, _f, _f, _p, _p, _p, _p, _i, _p],
    "dlv3p_bn_fold": [_p, _p, _p, _p, _i, _f, _p, _p, _p],
    "dlv3p_affine_act": [_p, _l, _p, _p, _i, _p, _l, _p, _l, _l, _i, _i, _p],
    "dlv3p_bn_bwd_reduce": [_p, _l, _p, _l, _p, _p, _p, _p, _i, _l, _i, _p, _i, _p],
    "dlv3p_bn_bwd_apply": [_p, _l, _p, _l, _p, _p, _p, _p, _i, _p, _l, _i, _p, _l, _i, _p],
    "dlv3p_act_bwd": [_p, _p, _p, _i, _p, _l, _i, _p],
    "dlv3p_add": [_p, _p, _p, _l, _i, _p],
    "dlv3p_copy2d": [_p, _l, _p, _l, _l, _i, _p, _l, _i, _p],
    "dlv3p_maxpool3x3s2_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "dlv3p_maxpool3x3s2_bwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "dlv3p_maxpool3x3s2_bn_fwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "dlv3p_maxpool3x3s2_bn_bwd": [_p, _p, _p, _p, _p, _p, _p, _d, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dlv3p_avgpool_fwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dlv3p_avgpool_bwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "dlv3p_bilinear_fwd": [_p, _l, _p, _l, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dlv3p_bilinear_bwd": [_p, _l, _p, _l, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p],
    "dlv3p_softmax_cbloss_fwd": [_p, _p, _p, _p, _f, _l, _i, _p, _p, _p],
    "dlv3p_softmax_cbloss_bwd": [_p, _p, _p, _p, _f, _l, _i, _f, _p, _p],
    "dlv3p_upsample_softmax_cbloss_fwd": [_p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _p, _p],
    "dlv3p_upsample_softmax_cbloss_bwd": [_p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _f, _p, _p],
    "dlv3p_upsample_softmax_cbloss_fwd_bwd": [_p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _f, _p, _p, _p],
    "dlv3p_softmax_argmax": [_p, _l, _i, _p, _p, _p],
    "dlv3p_cbloss_dense_fwd": [_p, _p, _p, _p, _f, _l, _i, _p, _p],
    "dlv3p_cbloss_dense_bwd": [_p, _p, _p, _p, _f, _l, _i, _f, _p, _p],
    "dlv3p_softmax_bwd": [_p, _p, _l, _i, _p, _p],
    "dlv3p_upsample_argmax": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "dlv3p_confusion_matrix": [_p, _p, _l, _i, _p, _p],
    "dlv3p_dropout": [_p, _p, _l, _f, _u64, _p, _p, _i, _p],
    "dlv3p_adam": [_p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _f, _p],
    "dlv3p_sumsq": [_p, _l, _p, _p],
    "dlv3p_cast": [_p, _i, _p, _i, _l, _p],
    "dlv3p_preprocess_image_batch": [_p, _i, _p, _i, _i, _p],
    "dlv3p_preprocess_label_batch": [_p, _i, _p, _i, _i, _p],
    "dlv3p_cast2d": [_p, _l, _i, _p, _l, _i, _l, _i, _p],
}
_PLAIN = {"dlv3p_version": [], "dlv3p_device_arch": [], "dlv3p_set_pdl": [_i]}

_lib = None
PROFILER = None      # set to a profiler.KernelProfiler to bracket every entry-point call with CUDA events


def load() -> C.CDLL:
    """dlopen libdlv3p.so (once) and declare every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with deeplabv3plus_keras_b200/csrc/build.sh "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in {**SIGNATURES, **_PLAIN}.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.dlv3p_last_error.argtypes = []
    lib.dlv3p_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def set_pdl(enabled: bool) -> bool:
    """Programmatic dependent launch on/off for all subsequent launches (a launch attribute; see include/dlv3p.h).
    Returns the previous setting."""
    return bool(load().dlv3p_set_pdl(1 if enabled else 0))


def last_error() -> str:
    return load().dlv3p_last_error().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    """Invoke an entry point; translate a negative status into the Python exception the reference would raise
    (ValueError for bad shapes/config — cf. ss.py:771,858 — RuntimeError for device failures)."""
    prof = PROFILER
    if prof is not None:
        tok = prof.before(name, args)
        rc = getattr(load(), name)(*args)
        prof.after(tok)
    else:
        rc = getattr(load(), name)(*args)
    if rc != 0:
        msg = f"{name} failed ({rc}): {last_error()}"
        if rc in (ERR_SHAPE, ERR_DTYPE, ERR_ALIGN, ERR_UNSUPPORTED):
            raise ValueError(msg)
        raise RuntimeError(msg)

"""The tf.load_op_library route (tf_ops/): exercised only where TensorFlow exists — it does not in the build image
(SURVEY.md §0.3), so these tests skip here and on the GPU box.  What runs everywhere: the registration source only
names C symbols that include/dlv3p.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tf_ops_source_binds_declared_symbols():
    src = open(os.path.join(ROOT, "tf_ops", "dlv3p_tf_ops.cc")).read()
    hdr = open(os.path.join(ROOT, "include", "dlv3p.h")).read()
    declared = set(re.findall(r"\b(dlv3p_[a-z0-9_]+)\s*\(", hdr))
    used = set(re.findall(r"\b(dlv3p_[a-z0-9_]+)\s*\(", src))
    assert used and used <= declared, sorted(used - declared)


def test_tf_custom_ops_match_stock_tf():
    tf = pytest.importorskip("tensorflow")
    if not os.path.exists(os.path.join(ROOT, "tf_ops", "libdlv3p_tf_ops.so")) or not tf.config.list_physical_devices("GPU"):
        pytest.skip("custom-op library not built / no GPU")
    import numpy as np

    from tf_ops import dlv3p_tf
    x = tf.cast(tf.random.uniform([2, 33, 33, 64], -1, 1, seed=1024), tf.bfloat16)
    ref = tf.keras.layers.SeparableConv2D(128, 3, dilation_rate=(6, 6), padding="same", use_bias=False)
    ref.build(x.shape)
    mine = dlv3p_tf.SeparableConv2D(128, 3, dilation_rate=(6, 6), padding="same", use_bias=False)
    mine.build(x.shape)
    mine.set_weights(ref.get_weights())
    a = tf.cast(ref(tf.cast(x, tf.float32)), tf.float32).numpy()
    b = tf.cast(mine(x), tf.float32).numpy()
    assert np.abs(a - b).max() <= 2e-2 * np.abs(a).max()

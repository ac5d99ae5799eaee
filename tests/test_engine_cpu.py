"""Host-logic parity without a GPU: engine.Plan (graph flattening, fusion, hand-written backward schedule,
gradient accumulation) driven through the tests/fake_ops.py test double, against the independent oracle graph."""
import numpy as np
import pytest
import torch

from oracle import model as OM
from tests import fake_ops, util
from tests.test_ops_gpu import NW, PW


@pytest.fixture
def cpu_engine(monkeypatch):
    from deeplabv3plus_keras_b200 import engine
    monkeypatch.setattr(engine, "ops", fake_ops)
    return engine


CASES = [
    dict(base="xception", output_stride=16, image_size=65),
    dict(base="xception", output_stride=8, image_size=49, refine=True, rate_mult=2),
    dict(base="mobilenetv2", output_stride=16, image_size=65, aspp=util.DEFAULT_ASPP),
    dict(base="mobilenetv2", output_stride=8, image_size=48, refine=True, aspp=util.DEFAULT_ASPP),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['base']}-os{c['output_stride']}-{'br' if c.get('refine') else 'plain'}")
def test_train_step_matches_oracle(cpu_engine, case):
    conf = util.make_conf(width=64, **case)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    B = 2
    plan = cpu_engine.Plan(ss.model, B, training=True)
    x, y = util.synthetic_batch(conf, B, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x, y)
    plan.step_fwd_bwd()
    plan.regularization()

    w = util.torch_weights(ss.model)
    data, l2, grads, out = OM.loss_and_grads(conf, w, torch.from_numpy(x).double(), torch.from_numpy(y), PW, NW)
    assert tuple(out["probs"].shape) == plan.out_shape
    np.testing.assert_allclose(plan.logits.buf.numpy(), out["logits"].detach().numpy(), rtol=2e-3, atol=2e-4)
    assert abs(plan.loss_value() - float(data + l2)) < 1e-4 * max(1.0, float(data))
    got = plan.gradients()
    assert set(got) == set(grads), set(got) ^ set(grads)
    lam = conf["hps"]["weight_decay"]
    regularised = {f"{l.name}/kernel" for l in ss.model.flat_layers() if getattr(l, "kernel_regularizer", None) is not None}
    for k, g in grads.items():
        g = g.numpy().copy()
        if k in regularised:
            g -= 2 * lam * w[k].numpy()       # the engine applies the L2 term inside Adam
        scale = max(np.abs(g).max(), 1e-3)      # gradients that are analytically zero stay at fp32 noise
        # fp32 test double vs fp64 oracle: a ReLU / max-pool decision that flips at a near-tie moves isolated
        # elements, so require 99.9% of each tensor within 3% of its max and the rest within 30%
        err = np.abs(got[k] - g) / scale
        assert (err > 3e-2).mean() <= 1e-3 and err.max() < 0.3, (k, float(err.max()), float((err > 3e-2).mean()))
    # moving statistics (updated twice where the shared base runs twice)
    plan.params.download()
    for k, v in out["new_stats"].items():
        np.testing.assert_allclose(ss.model.named_weights()[k], v.numpy(), rtol=1e-3, atol=1e-4, err_msg=k)


@pytest.mark.parametrize("fuse", [False, True], ids=["materialised-bn", "bn-fused-into-depthwise"])
def test_bn_relu_fused_into_depthwise_schedule(cpu_engine, monkeypatch, fuse):
    """Conv -> BN -> ReLU -> SeparableConv2D: with the fusion on, the BN+ReLU output is virtual (applied on load by the
    reader), the reader's input-gradient launch also produces the BN-backward reductions (dwconv3x3_dgrad_bnred) and
    the separate bn_bwd_reduce disappears; logits, loss and every gradient must equal the materialised schedule."""
    monkeypatch.setattr(cpu_engine, "FORCE_BNRED", True)
    conf = util.make_conf(width=64, base="xception", output_stride=16, image_size=65)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    calls = []
    for name in ("bn_bwd_reduce", "dwconv3x3_dgrad_bnred", "bn_train_apply", "bn_finalize"):
        orig = getattr(fake_ops, name)
        monkeypatch.setattr(fake_ops, name, (lambda orig, name: lambda *a, **k: (calls.append(name), orig(*a, **k))[1])(orig, name))
    plan = cpu_engine.Plan(ss.model, 2, training=True, fuse_bn_dw=fuse)
    x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x, y)
    plan.step_fwd_bwd()
    n_virtual = sum(isinstance(v, cpu_engine._BnActValue) for v in plan.values.values())
    if fuse:
        # Xception(ref-truncated): block2/3/4/13 sepconv1 + 2 per middle block (8 blocks) = 20 virtual tensors
        assert n_virtual == 20 and calls.count("dwconv3x3_dgrad_bnred") == 20
        assert calls.count("bn_finalize") >= 20
    else:
        assert n_virtual == 0 and "dwconv3x3_dgrad_bnred" not in calls
    ga = plan.gradients()               # (the gradient arena is shared by the plans of one model: snapshot first)
    ref = cpu_engine.Plan(ss.model, 2, training=True, fuse_bn_dw=False)
    ref.set_loss(PW, NW)
    ref.load_batch(x, y)
    ref.step_fwd_bwd()
    np.testing.assert_allclose(plan.logits.buf.numpy(), ref.logits.buf.numpy(), rtol=1e-4, atol=1e-5)
    assert abs(plan.loss_value() - ref.loss_value()) < 1e-5
    gb = ref.gradients()
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 2e-3 * scale, k


def test_block_closing_bn_reductions_fold_into_the_readers_backward(cpu_engine, monkeypatch):
    """Xception block output = BN(sepconv3) + residual, read by the next block's first SeparableConv2D through a
    pre-activation ReLU.  When that reader's fused depthwise backward (dlv3p_dwconv3x3_bwd) writes the FINAL gradient of
    the block output, it also produces the BN-backward reductions of sepconv3 (bn_y) and the producer's own
    bn_bwd_reduce pass disappears; every gradient must equal the unfolded schedule."""
    monkeypatch.setattr(cpu_engine, "FORCE_BNRED", True)
    conf = util.make_conf(width=64, base="xception", output_stride=16, image_size=65)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    x = y = None
    res = {}
    for fold in (True, False):
        monkeypatch.setattr(cpu_engine, "FOLD_BLOCK_RED", fold)
        calls = []
        with monkeypatch.context() as mp:
            for name in ("bn_bwd_reduce", "dwconv3x3_bwd"):
                orig = getattr(fake_ops, name)
                mp.setattr(fake_ops, name, (lambda orig, name: lambda *a, **k: (calls.append((name, k.get("bn_y") is not None)),
                                                                                orig(*a, **k))[1])(orig, name))
            plan = cpu_engine.Plan(ss.model, 2, training=True)
            if x is None:
                x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
            plan.set_loss(PW, NW)
            plan.load_batch(x, y)
            plan.step_fwd_bwd()
            res[fold] = (plan.gradients(), plan.loss_value(), calls.count(("dwconv3x3_bwd", True)),
                         sum(1 for c in calls if c[0] == "bn_bwd_reduce"))
    (ga, la, n_y, n_red), (gb, lb, n_y0, n_red0) = res[True], res[False]
    # middle flow: the outputs of blocks 4..11 are read by blocks 5..12 (8 folds); entry flow: whatever closes last
    assert n_y0 == 0 and n_y >= 8 and n_red0 - n_red == n_y, (n_y, n_red, n_red0)
    assert abs(la - lb) < 1e-6
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 2e-3 * scale, k


def test_channel_pitch_padding_is_invisible_and_skips_concatenated_widths(cpu_engine):
    """728-channel tensors live on a 736-channel pitch (pad parameters zero, Keras-shaped views outside); a width that
    reaches a Concatenate (ASPP reduction_size = 72 here) keeps its logical pitch."""
    conf = util.make_conf(width=72, base="xception", output_stride=16, image_size=65)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    plan = cpu_engine.Plan(ss.model, 2, training=True)
    by_name = {v.name: v for v in plan.values.values() if hasattr(v, "name")}
    assert by_name["block5_sepconv1/bn_act"].shape[3] == 736 and by_name["block5_sepconv1/bn_act"].clog == 728
    assert all(v.shape[3] == v.clog for v in plan.values.values() if hasattr(v, "clog") and v.clog == 72)
    lay = [l for l in ss.model.flat_layers() if l.name == "block5_sepconv2"][0]
    assert tuple(plan.params.view(lay, "pointwise_kernel").shape) == (1, 1, 736, 736)
    assert tuple(plan.params.logical(lay, "pointwise_kernel").shape) == (1, 1, 728, 728)
    assert float(plan.params.view(lay, "pointwise_kernel")[0, 0, 728:, :].abs().max()) == 0.0
    x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x, y)
    plan.step_fwd_bwd()
    g = plan.params.view(lay, "pointwise_kernel", grad=True)
    assert float(g[0, 0, 728:, :].abs().max()) == 0.0 and float(g[0, 0, :, 728:].abs().max()) == 0.0
    assert float(g.abs().max()) > 0.0
    assert plan.gradients()["block5_sepconv2/pointwise_kernel"].shape == (1, 1, 728, 728)
    assert plan.params.num_params == sum(int(np.prod(w.shape)) for l in ss.model.flat_layers()
                                         for n, w in l._weights.items() if l._trainable[n])


def test_concat_slices_written_in_place(cpu_engine, monkeypatch):
    """ASPP branches whose only reader is the Concatenate write their BN+ReLU output straight into the concat buffer
    (bn_train_apply ld_out) and read their gradient from the concat gradient in place (ld_dz): four of the five slice
    copies disappear in each direction (the x1 resize of the pooling branch is looked through) (branch 0 also feeds the atrous branches and keeps its copy)."""
    conf = util.make_conf(width=64, base="xception", output_stride=16, image_size=65)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    calls = []
    orig = fake_ops.copy2d
    monkeypatch.setattr(fake_ops, "copy2d", lambda *a, **k: (calls.append("copy2d"), orig(*a, **k))[1])
    plans = {}
    for flag in (True, False):
        calls.clear()
        plan = cpu_engine.Plan(ss.model, 2, training=True, concat_in_place=flag)
        x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
        plan.set_loss(PW, NW)
        plan.load_batch(x, y)
        plan.step_fwd_bwd()
        plans[flag] = (plan, len(calls), plan.gradients())
    assert plans[False][1] == 10 and plans[True][1] == 2, (plans[True][1], plans[False][1])
    a, b = plans[True][0], plans[False][0]
    np.testing.assert_allclose(a.logits.buf.numpy(), b.logits.buf.numpy(), rtol=1e-5, atol=1e-6)
    ga, gb = plans[True][2], plans[False][2]
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 1e-4 * scale, k


def test_boundary_refinement_resizes_write_concat_slices(cpu_engine, monkeypatch):
    """Boundary refinement (ss.py:915-954): the x(OS/2) resizes of the encoder output and of the low-level features are
    read only by the Concatenate, so they write their channel slice of the concat buffer directly (bilinear_fwd ld_y)
    and read their gradient slice in place (bilinear_bwd ld_dy): the four slice copies of the 304-channel tensor go."""
    conf = util.make_conf(width=64, base="xception", output_stride=8, image_size=49, refine=True, rate_mult=2)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    calls = []
    orig = fake_ops.copy2d
    monkeypatch.setattr(fake_ops, "copy2d", lambda *a, **k: (calls.append((a[1], a[3])), orig(*a, **k))[1])
    plans = {}
    for flag in (True, False):
        calls.clear()
        plan = cpu_engine.Plan(ss.model, 2, training=True, concat_in_place=flag)
        x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
        plan.set_loss(PW, NW)
        plan.load_batch(x, y)
        plan.step_fwd_bwd()
        wide = [c for c in calls if 64 + 48 in c]              # copies into / out of the refinement concat (64 + 48 channels)
        plans[flag] = (plan, len(wide), plan.gradients())
    assert plans[False][1] == 4 and plans[True][1] == 0, (plans[True][1], plans[False][1])
    a, b = plans[True][0], plans[False][0]
    np.testing.assert_allclose(a.logits.buf.numpy(), b.logits.buf.numpy(), rtol=1e-5, atol=1e-6)
    ga, gb = plans[True][2], plans[False][2]
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 1e-4 * scale, k


def test_bn_fused_into_maxpool_schedule(cpu_engine, monkeypatch):
    """Conv -> BN -> MaxPooling2D (Xception block2/3/4): the BN output is virtual, the pool applies scale/shift on the fly
    and keeps the raw winners, the producer's backward reduces over the pooled tensors and runs pool backward + BN input
    gradient in one launch.  Must reproduce the materialised schedule."""
    monkeypatch.setattr(cpu_engine, "FORCE_BNRED", True)
    conf = util.make_conf(width=64, base="xception", output_stride=16, image_size=65)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    calls = []
    for name in ("maxpool3x3s2_bn_fwd", "maxpool3x3s2_bn_bwd", "maxpool3x3s2_fwd", "maxpool3x3s2_bwd"):
        orig = getattr(fake_ops, name)
        monkeypatch.setattr(fake_ops, name, (lambda orig, name: lambda *a, **k: (calls.append(name), orig(*a, **k))[1])(orig, name))
    plans = {}
    for flag in (True, False):
        calls.clear()
        ss.model._invalidate()          # both plans start from the initial moving statistics (shared store)
        plan = cpu_engine.Plan(ss.model, 2, training=True, fuse_bn_pool=flag)
        x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
        plan.set_loss(PW, NW)
        plan.load_batch(x, y)
        plan.step_fwd_bwd()
        moving = {k: plan.params.logical(l, n).clone().numpy() for l in ss.model.flat_layers()
                  for n in l.weight_names() if "moving" in n for k in [f"{l.name}/{n}"]}
        plans[flag] = (plan, list(calls), plan.gradients(), moving)
    assert plans[True][1].count("maxpool3x3s2_bn_fwd") == 3 and plans[True][1].count("maxpool3x3s2_bn_bwd") == 3
    assert "maxpool3x3s2_bn_fwd" not in plans[False][1] and plans[False][1].count("maxpool3x3s2_bwd") == 3
    a, b = plans[True][0], plans[False][0]
    np.testing.assert_allclose(a.logits.buf.numpy(), b.logits.buf.numpy(), rtol=1e-4, atol=1e-5)
    assert abs(a.loss_value() - b.loss_value()) < 1e-5
    ga, gb = plans[True][2], plans[False][2]
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 2e-3 * scale, k
    sa, sb = plans[True][3], plans[False][3]
    assert sa and any(np.abs(v - ss.model.named_weights()[k]).max() > 1e-6 for k, v in sa.items())
    for k, v in sb.items():
        np.testing.assert_allclose(sa[k], v, rtol=1e-5, atol=1e-6, err_msg=k)


def test_implicit_conv_schedule(cpu_engine, monkeypatch):
    """Xception block1_conv2 (3x3 VALID stride 1, 32 -> 64): the implicit-GEMM schedule (no im2col / col2im, prepared
    wk / wd filter matrices) must reproduce the im2col + GEMM schedule."""
    monkeypatch.setattr(cpu_engine, "FORCE_IMPLICIT", True)
    conf = util.make_conf(width=64, base="xception", output_stride=16, image_size=65)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    calls = []
    for name in ("im2col3x3", "col2im3x3", "conv3x3_valid_fwd", "conv3x3_valid_dgrad", "conv3x3_valid_wgrad"):
        orig = getattr(fake_ops, name)
        monkeypatch.setattr(fake_ops, name, (lambda orig, name: lambda *a, **k: (calls.append(name), orig(*a, **k))[1])(orig, name))
    plan = cpu_engine.Plan(ss.model, 2, training=True)
    x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x, y)
    plan.step_fwd_bwd()
    assert calls.count("conv3x3_valid_fwd") == 1 and calls.count("conv3x3_valid_dgrad") == 1
    assert calls.count("conv3x3_valid_wgrad") == 1
    n_im2col = calls.count("im2col3x3")
    calls.clear()
    ga = plan.gradients()
    ref = cpu_engine.Plan(ss.model, 2, training=True, implicit_conv=False)
    ref.set_loss(PW, NW)
    ref.load_batch(x, y)
    ref.step_fwd_bwd()
    assert calls.count("im2col3x3") == n_im2col + 2 and "conv3x3_valid_fwd" not in calls      # block1_conv2 + logits conv
    np.testing.assert_allclose(plan.logits.buf.numpy(), ref.logits.buf.numpy(), rtol=1e-4, atol=1e-5)
    gb = ref.gradients()
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 2e-3 * scale, k


@pytest.mark.parametrize("case", [dict(base="xception", output_stride=16, image_size=65),
                                  dict(base="xception", output_stride=8, image_size=49, refine=True, rate_mult=2)],
                         ids=["logits-256ch", "logits-304ch-after-refinement"])
def test_implicit_same_conv_schedule(cpu_engine, monkeypatch, case):
    """The logits convolution (3x3 SAME, ss.py:893-897; 304 input channels after boundary refinement): the implicit-GEMM
    schedule (conv3x3_same_fwd / _dgrad / _wgrad, prepared wt / wd filter matrices, no im2col / col2im) must reproduce
    the im2col + GEMM schedule."""
    monkeypatch.setattr(cpu_engine, "FORCE_IMPLICIT", True)
    conf = util.make_conf(width=64, **case)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    calls = []
    for name in ("im2col3x3", "col2im3x3", "conv3x3_same_fwd", "conv3x3_same_dgrad", "conv3x3_same_wgrad"):
        orig = getattr(fake_ops, name)
        monkeypatch.setattr(fake_ops, name, (lambda orig, name: lambda *a, **k: (calls.append(name), orig(*a, **k))[1])(orig, name))
    plan = cpu_engine.Plan(ss.model, 2, training=True)
    x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x, y)
    plan.step_fwd_bwd()
    assert calls.count("conv3x3_same_fwd") == 1 and calls.count("conv3x3_same_dgrad") == 1
    assert calls.count("conv3x3_same_wgrad") == 1 and "col2im3x3" not in calls
    assert calls.count("im2col3x3") == 1                     # only the 3-channel stride-2 first convolution is left
    ga = plan.gradients()
    calls.clear()
    ref = cpu_engine.Plan(ss.model, 2, training=True, implicit_conv=False)
    ref.set_loss(PW, NW)
    ref.load_batch(x, y)
    ref.step_fwd_bwd()
    assert "conv3x3_same_fwd" not in calls and calls.count("col2im3x3") >= 1
    np.testing.assert_allclose(plan.logits.buf.numpy(), ref.logits.buf.numpy(), rtol=1e-4, atol=1e-5)
    gb = ref.gradients()
    for k in gb:
        scale = max(np.abs(gb[k]).max(), 1e-3)
        assert np.abs(ga[k] - gb[k]).max() <= 2e-3 * scale, k


def test_inference_matches_oracle(cpu_engine):
    conf = util.make_conf(base="mobilenetv2", image_size=65, aspp=util.DEFAULT_ASPP, width=32)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    plan = cpu_engine.Plan(ss.model, 1, training=False)
    x, _ = util.synthetic_batch(conf, 1, plan.out_shape[1:3])
    probs = plan.predict(x)
    out = OM.forward(conf, util.torch_weights(ss.model), torch.from_numpy(x).double(), training=False)
    assert probs.shape == (1, 80, 80, 21)            # 65 -> 5x5 features -> x16 (the reference's own arithmetic)
    np.testing.assert_allclose(probs, out["probs"].numpy(), rtol=1e-3, atol=1e-5)
    assert (plan.segment(x) == out["probs"].argmax(-1).numpy()).mean() >= 0.999


def test_weight_npz_round_trip(tmp_path):
    """save_weights_npz / load_weights_npz move every weight by its tf.keras layer name (checkpoint / resume path of
    the reference: ModelCheckpoint + model_loading, ss.py:482-485, 983-986)."""
    import numpy as np
    from deeplabv3plus_keras_b200 import SemanticSegmentation
    from deeplabv3plus_keras_b200.utils import load_weights_npz, save_weights_npz
    from tests import util
    conf = util.make_conf(base="mobilenetv2", image_size=33, width=32)
    a = util.build(conf)
    util.randomize_weights(a.model, seed=5)
    path = str(tmp_path / "semantic_segmentation_deeplabv3plus.npz")
    save_weights_npz(a.model, path)
    b = util.build(conf)
    rep = load_weights_npz(b.model, path)
    assert set(rep.values()) == {"loaded"} and "Conv1/kernel" in rep and "bn_Conv1/moving_variance" in rep
    for k, v in a.model.named_weights().items():
        np.testing.assert_array_equal(v, b.model.named_weights()[k])
    conf2 = dict(conf, model_loading=True, resource_path=str(tmp_path))
    c = util.build(conf2)                                   # the reference's resume switch
    np.testing.assert_array_equal(c.model.named_weights()["Conv1/kernel"], a.model.named_weights()["Conv1/kernel"])
    assert SemanticSegmentation is type(c)


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[2]], ids=["xception", "xception-br", "mobilenetv2"])
def test_gradient_prefixes_are_final_during_backward(cpu_engine, case):
    """Data-parallel overlap (trainer.py): after the backward launches [0, cut) the arena prefixes reported by
    Plan.final_prefixes must already hold their FINAL values, and the last cut must cover a growing share."""
    conf = util.make_conf(width=32, **case)
    ss = util.build(conf)
    util.randomize_weights(ss.model)
    plan = cpu_engine.Plan(ss.model, 2, training=True)
    x, y = util.synthetic_batch(conf, 2, plan.out_shape[1:3])
    plan.set_loss(PW, NW)
    plan.load_batch(x, y)
    plan.zero_grads()
    plan.forward()
    plan.loss_forward_backward()
    P = plan.params
    n = len(plan.bwd)
    cuts = [round(i * n / 4) for i in range(1, 5)]
    snaps, pos = [], 0
    for c in cuts:
        plan.run_bwd_range(pos, c)
        pos = c
        snaps.append((plan.final_prefixes(c), P.g.clone()))
    final = P.g
    prev = (0, P.n_reg)
    for (a, b), g in snaps:
        assert a >= prev[0] and b >= prev[1]
        torch.testing.assert_close(g[:a], final[:a], rtol=0, atol=0)
        torch.testing.assert_close(g[P.n_reg:b], final[P.n_reg:b], rtol=0, atol=0)
        prev = (a, b)
    assert prev == (P.n_reg, P.n_train), "every trainable parameter has a gradient producer"
    assert snaps[1][0][1] > P.n_reg, "half-way through backward a non-empty prefix is already final"


def test_reduce_lr_on_plateau_keras_semantics():
    """ss.py:978-982: monitor loss, patience 5, min_lr 1e-8 — the learning rate drops by `factor` after `patience`
    epochs without an improvement larger than min_delta, never below min_lr, and Adam's step size follows."""
    from deeplabv3plus_keras_b200.deeplab import Adam, ReduceLROnPlateau
    opt = Adam(lr=1e-4, beta_1=0.5, beta_2=0.99)
    cb = ReduceLROnPlateau(opt, monitor="loss", factor=0.5, patience=5, min_lr=3e-5)
    lrs = [cb.on_epoch_end(v) for v in [1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.89995, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]]
    assert lrs[:6] == [1e-4] * 6                  # epochs 2..6 are the 5 patient ones
    assert lrs[6] == 5e-5 and lrs[7] == 5e-5      # reduced once; 0.89995 is within min_delta of the best
    assert lrs[8] == 5e-5 and lrs[-1] == 3e-5     # improvement resets the wait; second reduction clamps at min_lr
    t1 = opt.step_size()
    opt.lr = 1e-4
    assert abs(opt.step_size() / t1 - 1e-4 / 3e-5) < 1e-9
    with pytest.raises(ValueError):
        ReduceLROnPlateau(opt, factor=1.0)


def test_trained_state_is_shared_by_all_plans_and_checkpointed(cpu_engine, monkeypatch, tmp_path):
    """One parameter store per model: after train_on_batch the inference plan (predict), get_weights, a training
    plan for ANOTHER batch size (last partial batch) and a saved checkpoint all see the trained weights, Adam moments,
    iteration count and BN moving statistics; resuming from the checkpoint continues the optimizer exactly."""
    from deeplabv3plus_keras_b200.utils import load_weights_npz, save_weights_npz
    conf = util.make_conf(base="mobilenetv2", image_size=33, width=32, aspp=util.DEFAULT_ASPP, dropout=0.5)
    conf["hps"]["lr"] = 1e-2

    def fresh():
        ss = util.build(conf)
        util.randomize_weights(ss.model, seed=3)
        return ss

    ss = fresh()
    m = ss.model
    p_inf = m.plan(2, training=False)
    x, y = util.synthetic_batch(conf, 2, p_inf.out_shape[1:3])
    before = m.predict(x, batch_size=2)
    w0 = [w.copy() for w in m.get_weights()]
    losses = [m.train_on_batch(x, y) for _ in range(3)]
    assert np.isfinite(losses).all()
    after = m.predict(x, batch_size=2)
    assert np.abs(after - before).max() > 1e-4, "predict() still sees the initial weights"
    w1 = m.get_weights()
    assert sum(float(np.abs(a - b).max()) > 0 for a, b in zip(w0, w1)) > len(w0) // 2, "get_weights() is stale"
    store = m.plan(2, training=True).params
    assert m.plan(1, training=True).params is store and p_inf.params is store
    assert float(store.m.abs().max()) > 0 and m.optimizer.iterations == 3 and int(store.step_counter) == 3

    # checkpoint -> resume in a fresh model -> the 4th step is the same as continuing in place
    path = str(tmp_path / "ckpt.npz")
    save_weights_npz(m, path)
    cont = m.train_on_batch(x[:1], y[:1])                  # partial batch: another plan, same state
    ss2 = fresh()
    load_weights_npz(ss2.model, path)
    assert ss2.model.optimizer.iterations == 3
    resumed = ss2.model.train_on_batch(x[:1], y[:1])
    assert abs(resumed - cont) < 1e-5 * max(1.0, abs(cont)), (resumed, cont)
    st2 = ss2.model.plan(1, training=True).params
    np.testing.assert_allclose(st2.m.numpy(), store.m.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(st2.w.numpy(), store.w.numpy(), rtol=1e-5, atol=1e-7)
    assert int(st2.step_counter) == 4

    # set_weights on one layer keeps every other layer's trained values
    lay = [l for l in m.flat_layers() if l.name == "Conv1"][0]
    other = [l for l in m.flat_layers() if l.name == "block_3_project"][0]
    trained_other = other.get_weights()[0].copy()
    lay.set_weights([np.zeros_like(v) for v in lay.get_weights()])
    m.predict(x, batch_size=2)
    assert float(store.logical(lay, "kernel").abs().max()) == 0.0
    np.testing.assert_array_equal(store.logical(other, "kernel").numpy(), trained_other)

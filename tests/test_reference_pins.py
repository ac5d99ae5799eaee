"""Pins against outputs of the REFERENCE'S OWN SOURCE (scripts/make_golden_reference_fns.py executes
ss.py:438-447 `class_balanced_loss`, ss.py:410-420 `class_imbalance_loss`, ss.py:290-334 `MeanIoUExt.update_state`
and ss.py:459-525 + 770-954 `SemanticSegmentation.__init__/_make_encoder/_make_decoder/_refine_boundary` and commits
what they return under tests/golden/):

* the oracle restatements (oracle/tf_ops.py) and the CUDA kernels both reproduce the reference functions' values;
* deeplab.py builds, layer for layer, the graph the reference's builder source builds through this repo's keras
  surface (types, tf.keras names, constructor arguments, wiring, output shapes, weight shapes) — the drop-in claim.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from tests import topology, util

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
LOSS_CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz"))
                    if "voc_weights" not in p)
MIOU_CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "miou_*.npz")))


def test_fixtures_present():
    assert len(LOSS_CASES) == 5 and len(MIOU_CASES) == 3
    assert len(glob.glob(os.path.join(GOLDEN, "topology_*.json"))) == len(topology.TOPOLOGY_CASES)


def test_voc_class_weights_are_the_reference_values():
    from deeplabv3plus_keras_b200 import deeplab
    g = np.load(os.path.join(GOLDEN, "loss_voc_weights.npz"))
    np.testing.assert_array_equal(np.asarray(deeplab.ss_pw), g["ss_pw"])
    np.testing.assert_array_equal(np.asarray(deeplab.ss_nw), g["ss_nw"])


@pytest.mark.parametrize("name", LOSS_CASES)
def test_oracle_loss_reproduces_reference_function(name):
    g = np.load(os.path.join(GOLDEN, f"loss_{name}.npz"))
    got = T.class_balanced_loss(torch.from_numpy(g["y_true"]), torch.from_numpy(g["y_pred"]),
                                list(g["pos_weights"]), list(g["neg_weights"]), float(g["epsilon"]))
    assert abs(float(got) - float(g["loss"])) <= 1e-12 * max(1.0, abs(float(g["loss"])))
    got32 = T.class_balanced_loss(torch.from_numpy(g["y_true"]).float(), torch.from_numpy(g["y_pred"]).float(),
                                  [float(np.float32(v)) for v in g["pos_weights"]],
                                  [float(np.float32(v)) for v in g["neg_weights"]], float(np.float32(g["epsilon"])))
    assert abs(float(got32) - float(g["loss_f32"])) <= 2e-6 * max(1.0, abs(float(g["loss_f32"])))


@pytest.mark.parametrize("name", MIOU_CASES)
def test_oracle_confusion_matrix_reproduces_reference_metric(name):
    g = np.load(os.path.join(GOLDEN, f"miou_{name}.npz"))
    C, accum = int(g["num_classes"]), bool(g["accum_enable"])
    total = torch.zeros((C, C), dtype=torch.float64)
    for b in range(g["y_true"].shape[0]):
        cm = T.confusion_matrix(T.argmax_labels(torch.from_numpy(g["y_true"][b])),
                                T.argmax_labels(torch.from_numpy(g["y_pred"][b])), C)
        total = total + cm if accum else cm
        np.testing.assert_array_equal(total.numpy(), g["total_cm"][b])


# ---------------------------------------------------------------------------------------------- topology
def _ours(case):
    conf = util.make_conf(**topology.TOPOLOGY_CASES[case])
    return topology.describe(util.build(conf).model)


@pytest.mark.parametrize("case", sorted(topology.TOPOLOGY_CASES))
def test_deeplab_builds_the_graph_of_the_reference_builder_source(case):
    want = json.load(open(os.path.join(GOLDEN, f"topology_{case}.json")))
    got = json.loads(json.dumps(_ours(case), sort_keys=True))
    assert sorted(got) == sorted(want)                       # deeplabv3plus / encoder / decoder / base
    for model_name in want:
        assert len(got[model_name]) == len(want[model_name]), model_name
        for a, b in zip(got[model_name], want[model_name]):
            assert a == b, (model_name, a.get("layer"), a, b)


@pytest.mark.skipif(not os.path.exists("/root/reference/bodhi/deeplabv3plus_keras/semantic_segmentation.py"),
                    reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("case", ["xception_os8_br", "mobilenetv2_os16_default_aspp"])
def test_topology_fixture_is_what_the_reference_source_builds_today(case):
    """Re-executes the reference's builder source live (build container only) and checks the committed fixture."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_golden_reference_fns", os.path.join(os.path.dirname(GOLDEN), "..", "scripts", "make_golden_reference_fns.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    live, _ = mod.reference_graph(topology.TOPOLOGY_CASES[case])
    want = json.load(open(os.path.join(GOLDEN, f"topology_{case}.json")))
    assert json.loads(json.dumps(live, sort_keys=True)) == want


# ---------------------------------------------------------------------------------------------- CUDA kernels
@pytest.mark.gpu
@pytest.mark.parametrize("name", LOSS_CASES)
def test_cuda_dense_loss_reproduces_reference_function(name):
    """deeplab.class_balanced_loss (Keras signature, dlv3p_cbloss_dense_fwd) against the reference function's value."""
    from deeplabv3plus_keras_b200.deeplab import ClassBalancedLoss, class_balanced_loss
    g = np.load(os.path.join(GOLDEN, f"loss_{name}.npz"))
    got = class_balanced_loss(g["y_true"], g["y_pred"], list(g["pos_weights"]), list(g["neg_weights"]),
                              float(g["epsilon"]))
    assert abs(got - float(g["loss_f32"])) <= 1e-5 * max(1.0, abs(float(g["loss_f32"]))), (got, float(g["loss_f32"]))
    assert abs(got - float(g["loss"])) <= 1e-4 * max(1.0, abs(float(g["loss"])))
    wrapped = ClassBalancedLoss(list(g["pos_weights"]), list(g["neg_weights"]), float(g["epsilon"]))
    assert abs(wrapped(g["y_true"], g["y_pred"]) - got) <= 1e-6 * max(1.0, abs(got))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["voc_soft", "voc_flat", "voc_sharp"])
def test_cuda_fused_softmax_loss_reproduces_reference_function(name):
    """The index-form fused kernel the training step uses (softmax + loss from logits and an integer label map):
    fed z = log(p) and labels = argmax(one-hot), it must give the reference function's value for (one-hot, p)."""
    from deeplabv3plus_keras_b200 import ops
    g = np.load(os.path.join(GOLDEN, f"loss_{name}.npz"))
    p, y = g["y_pred"], g["y_true"]
    C = p.shape[-1]
    P = p.size // C
    z = torch.from_numpy(np.log(np.maximum(p, 1e-300))).float().cuda().contiguous()
    lab = torch.from_numpy(y.argmax(-1).astype(np.int32)).cuda().contiguous()
    pw = torch.tensor(g["pos_weights"], dtype=torch.float32, device="cuda")
    nw = torch.tensor(g["neg_weights"], dtype=torch.float32, device="cuda")
    out = torch.zeros(1, device="cuda")
    ops.softmax_cbloss_fwd(z.view(P, C), lab.view(P), pw, nw, float(g["epsilon"]), P, C, out)
    got = float(out.item()) / P
    assert abs(got - float(g["loss"])) <= 2e-5 * max(1.0, abs(float(g["loss"]))), (got, float(g["loss"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", MIOU_CASES)
def test_cuda_mean_iou_reproduces_reference_metric(name):
    from deeplabv3plus_keras_b200.deeplab import MeanIoUExt
    g = np.load(os.path.join(GOLDEN, f"miou_{name}.npz"))
    C = int(g["num_classes"])
    m = MeanIoUExt(C, accum_enable=bool(g["accum_enable"]))
    for b in range(g["y_true"].shape[0]):
        cm = m.update_state(g["y_true"][b], g["y_pred"][b])
        np.testing.assert_array_equal(cm.cpu().numpy(), g["total_cm"][b])
    want = float(T.mean_iou(torch.from_numpy(g["total_cm"][-1])))
    assert abs(m.result() - want) < 1e-12

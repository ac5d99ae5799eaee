"""The teacher-forced / decision-forced parity harness (tests/teacher.py) driven through the torch-CPU test double of
the kernels (tests/fake_ops.py): checks the HOST side of the real engine schedule — what every macro-op reads and
writes, the hand-written backward wiring, gradient accumulation over shared tensors, concat slices, dropout mask
export — at fp32 round-off against the fp64 oracle, including the ASPP forms the shipped JSONs leave identity (conv
k=1 branch, true image pooling, chained pyramid level) and Dropout(0.5).  The same harness runs against the CUDA kernels
in tests/test_model_gpu.py."""
import pytest

from tests import fake_ops, teacher, util


@pytest.fixture
def cpu_engine(monkeypatch):
    from deeplabv3plus_keras_b200 import engine
    monkeypatch.setattr(engine, "ops", fake_ops)
    return engine


CASES = {
    "xception-os16": dict(base="xception", output_stride=16, image_size=65),
    "xception-os8-br": dict(base="xception", output_stride=8, image_size=49, refine=True, rate_mult=2),
    "mobilenetv2-os16-default-aspp": dict(base="mobilenetv2", output_stride=16, image_size=65, aspp=util.DEFAULT_ASPP),
    "xception-globalpool-dropout": dict(base="xception", output_stride=16, image_size=97, aspp="global_pool", dropout=0.5),
    "mobilenetv2-globalpool-dropout": dict(base="mobilenetv2", output_stride=16, image_size=64, aspp="global_pool",
                                           dropout=0.5),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_every_macro_op_on_identical_inputs_and_whole_graph_with_forced_decisions(cpu_engine, name):
    conf = util.make_conf(width=48, **CASES[name])
    res = teacher.run(conf, Plan=cpu_engine.Plan, step_counter=3)
    assert not res["unused_teacher"] and not res["missing_grad_points"]
    assert not res["unused_sites"] and not res["unforced_sites"]
    for part in ("fwd", "bwd", "param_tf"):
        for k, d in res[part].items():
            assert d["rms"] <= 2e-5 and d["q9999"] <= 1e-4, (part, k, d)
    for k, d in res["param_df"].items():
        assert d["rms"] <= 1e-3, (k, d)
    assert res["logits_df"]["rms"] <= 1e-4
    for k, v in res["param_zero"].items():
        assert v <= 1e-3, (k, v)
    flips = res["flips"]
    assert sum(f["count"] for f in flips.values()) <= 1e-4 * sum(f["total"] for f in flips.values())
    assert all(f["worst_margin"] <= 1e-4 for f in flips.values()), flips
    if "dropout" in name:
        (site,) = res["plan"].dropout_sites
        m = res["plan"].dropout_mask(site)
        keep = float((m > 0).float().mean())
        assert 0.4 < keep < 0.6 and set(m.unique().tolist()) <= {0.0, 2.0}
        assert f"{site}/out" in res["fwd"] and "average_pooling2d/out" in res["fwd"] and "lambda/out" in res["fwd"]

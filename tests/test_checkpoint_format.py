"""Checkpoint interchange: the fused engines write 'optimizer_state_dict' in torch.optim.Adam's own
layout (what the reference's restore feeds to Adam.load_state_dict, torch_model.py:1085-1090), and read it back; the
round-1 raw-slab layout is still accepted.  Host-only: the engine's slabs live on the CPU here, no kernel runs."""
import numpy as np
import pytest
import torch

from deepchem_b200 import _lib


def _engine_model(widths=(8, 12), dense=16, n_tasks=2, n_feat=75):
    from deepchem_b200.engine import FlatEngine
    from deepchem_b200.graphconvmodel import _GraphConvTorchModel
    torch.manual_seed(0)
    m = _GraphConvTorchModel(n_tasks, graph_conv_layers=list(widths), dense_layer_size=dense, mode="regression",
                             number_atom_features=n_feat, batch_size=4)
    return m, FlatEngine(m, "cpu", lr=2e-3)


def _fake_steps(eng, n):
    g = torch.Generator().manual_seed(1)
    for p, view, gview in eng._slots:                        # moments only where a real parameter lives
        ea, eas = eng._moment_views(view)
        ea.copy_(torch.randn(ea.shape, generator=g))
        eas.copy_(torch.rand(eas.shape, generator=g))
    eng.step_count = n


def test_engine_optimizer_state_is_a_torch_adam_state_dict():
    try:
        _lib.lib()
    except Exception as e:              # pragma: no cover
        pytest.skip("libdcgc not built: %s" % e)
    m, eng = _engine_model()
    _fake_steps(eng, 7)
    sd = eng.state_dict()
    params = list(m.parameters())
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(params)))
    assert sd["param_groups"][0]["lr"] == 2e-3 and tuple(sd["param_groups"][0]["betas"]) == (0.9, 0.999)
    for i, p in enumerate(params):
        assert tuple(sd["state"][i]["exp_avg"].shape) == tuple(p.shape) and float(sd["state"][i]["step"]) == 7.0
    # engine -> plain torch.optim.Adam over the same module (the per-layer autograd path / the reference's restore)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.load_state_dict(sd)
    for i, p in enumerate(params):
        assert torch.equal(opt.state[p]["exp_avg"], sd["state"][i]["exp_avg"])
        assert torch.equal(opt.state[p]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
    assert opt.param_groups[0]["lr"] == 2e-3
    # torch.optim.Adam -> engine (a checkpoint written by the autograd path or by the reference)
    m2, eng2 = _engine_model()
    eng2.load_state_dict(opt.state_dict())
    assert eng2.step_count == 7 and torch.equal(eng2.exp_avg, eng.exp_avg) and torch.equal(eng2.exp_avg_sq, eng.exp_avg_sq)
    assert eng2.lr == 2e-3
    # before the first step torch has no per-parameter state; neither does the engine
    m3, eng3 = _engine_model()
    assert eng3.state_dict()["state"] == {}
    eng2.load_state_dict(eng3.state_dict())
    assert eng2.step_count == 0 and float(eng2.exp_avg.abs().max()) == 0.0


def test_round1_slab_checkpoints_still_load():
    try:
        _lib.lib()
    except Exception as e:              # pragma: no cover
        pytest.skip("libdcgc not built: %s" % e)
    m, eng = _engine_model()
    _fake_steps(eng, 3)
    legacy = {"exp_avg": eng.exp_avg.clone(), "exp_avg_sq": eng.exp_avg_sq.clone(), "step": 3}
    m2, eng2 = _engine_model()
    eng2.load_state_dict(legacy)
    assert eng2.step_count == 3 and torch.equal(eng2.exp_avg, eng.exp_avg)


def test_mismatched_optimizer_state_is_refused():
    try:
        _lib.lib()
    except Exception as e:              # pragma: no cover
        pytest.skip("libdcgc not built: %s" % e)
    m, eng = _engine_model()
    _fake_steps(eng, 2)
    other, eng_o = _engine_model(widths=(8,))
    with pytest.raises(ValueError):
        eng_o.load_state_dict(eng.state_dict())


def test_evaluate_undoes_the_transformers_on_labels_and_predictions():
    """Evaluator.compute_model_performance (deepchem/utils/evaluate.py:303-307)."""
    from deepchem_b200.graphconvmodel import evaluate_model

    class Norm(object):                       # NormalizationTransformer-like: y' = (y - 3) / 2
        transform_y = True

        def untransform(self, y):
            return y * 2.0 + 3.0

    class DS(object):
        y = np.array([[0.0], [1.0], [-1.0]])
        w = np.ones((3, 1))

    class M(object):
        def predict(self, ds, transformers=[]):
            out = np.array([[0.5], [1.0], [-1.5]])
            for t in reversed(transformers):
                out = t.untransform(out)
            return out

    def mae(y, p):
        return float(np.mean(np.abs(y - p)))

    class Metric(object):
        name = "mae_metric"

        def compute_metric(self, y, y_pred, w, per_task_metrics=False, n_tasks=None):
            s = mae(y, y_pred)
            return (s, [s]) if per_task_metrics else s
    got = evaluate_model(M(), DS(), [mae, Metric()], [Norm()])
    expect = float(np.mean(np.abs(np.array([0.5, 0.0, -0.5]) * 2.0)))     # errors scale with the transformer
    assert got["mae"] == pytest.approx(expect) and got["mae_metric"] == pytest.approx(expect)
    scores, per_task = evaluate_model(M(), DS(), Metric(), [Norm()], per_task_metrics=True)
    assert scores["mae_metric"] == pytest.approx(expect) and per_task["mae_metric"] == [pytest.approx(expect)]

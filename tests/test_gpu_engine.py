"""The fused whole-model engine (dcgc_gcmodel_*) against the CPU oracle: loss, every gradient,
BatchNorm running statistics, Adam trajectory, eval-mode forward."""
import numpy as np
import pytest
import torch

from helpers import assert_fp64_anchored, oracle_batch, oracle_fp32_fp64, rel_err, torch_args
from oracle import graphconv_torch as O

pytestmark = pytest.mark.gpu
MODEL_TOL = 1e-4     # composite tolerance, see tests/test_gpu_parity.py


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _pair(mode, layers, dense, n_tasks, bsz, bn=True, seed=5):
    from deepchem_b200.graphconvmodel import GraphConvModel
    torch.manual_seed(seed)
    om = O.OracleGraphConvModel(n_tasks, layers, dense, mode=mode, batch_size=bsz, batch_normalize=bn)
    with torch.no_grad():
        for p in om.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    m = GraphConvModel(n_tasks, graph_conv_layers=layers, dense_layer_size=dense, mode=mode, batch_size=bsz,
                       batch_normalize=bn)
    assert m._engine is not None
    m.model.load_state_dict(om.state_dict())
    assert m._engine.aliased()
    return om, m


def _oracle_step(om, mode, pm, y_onehot_or_y, w, n):
    _, mm = oracle_batch(pm.to_list())
    om.train()
    oo = om(torch_args(mm, n))
    lo = O.standard_loss(mode, oo, torch.from_numpy(y_onehot_or_y), torch.from_numpy(w))
    return oo, lo


@pytest.mark.parametrize("mode,layers,bn,shape", [("regression", [64, 64], True, "stress"),
                                                  ("classification", [128, 128, 128], True, "zinc"),
                                                  ("regression", [32, 64], False, "stress"),
                                                  ("classification", [64], True, "delaney"),
                                                  ("classification", [64, 64], True, "tox21")])
def test_engine_train_step_matches_oracle(mode, layers, bn, shape):
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    n_tasks, dense, bsz = (12 if shape == "tox21" else 3), 128, 72     # tox21: BASELINE config 2 (12 tasks, missing labels)
    pm = make_molecules(70, seed=11, shape=shape)
    y, w = make_labels(pm.n_mols, n_tasks, mode, seed=2, missing=0.25)
    om, m = _pair(mode, layers, dense, n_tasks, bsz, bn)
    batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True, pad_batches=False))
    inputs, labels, weights = m._prepare_batch(batch)
    eng = m._engine
    out = torch.empty(pm.n_mols, eng.cfg.n_out, device=m.device)
    loss = eng.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0], weights[0], pm.n_mols, out=out)
    # float64-anchored bound (helpers.py): within 1e-5 of the tensor scale, or as close to float64 as the fp32 oracle;
    # a batch of ~1.7 k atoms gets the discontinuity allowance of one ReLU / argmax decision (flip = 2 / atoms)
    flip = 2.0 / max(1, pm.n_atoms)
    _, mm = oracle_batch(pm.to_list())
    res = oracle_fp32_fp64(om, mode, mm, pm.n_mols, batch[1][0], w)
    o32, l32, g32 = res[torch.float32]
    o64, l64, g64 = res[torch.float64]
    k = 1 if mode == "classification" else 0
    assert_fp64_anchored("output", out.cpu().reshape(o64[k].shape), o32[k], o64[k])
    assert abs(float(loss) - l64) <= max(1e-5, 3 * abs(l32 - l64) / abs(l64)) * abs(l64)
    for name, p in m.model.named_parameters():
        assert_fp64_anchored(name, p.grad, g32[name], g64[name], flip=flip)
    om.train()
    om(torch_args(mm, pm.n_mols))                       # the fp32 oracle's running statistics
    for (k, v), (_, vo) in zip(m.model.state_dict().items(), om.state_dict().items()):
        if "running" in k:
            assert rel_err(v.cpu().numpy(), vo.numpy()) < 1e-5, k


def test_engine_gradients_equal_autograd_path():
    """Two CUDA evaluations of the same math (fused engine vs per-layer autograd ops)."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    pm = make_molecules(200, seed=12, shape="zinc")
    y, w = make_labels(200, 2, "regression", seed=3)
    om, m = _pair("regression", [64, 128], 128, 2, 200)
    batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
    inputs, labels, weights = m._prepare_batch(batch)
    sd = {k: v.clone() for k, v in m.model.state_dict().items()}
    loss_e = float(m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0], weights[0], 200))
    g_engine = {n: p.grad.clone() for n, p in m.model.named_parameters()}
    m.model.load_state_dict(sd)                       # undo the running-stat update
    m._engine.grads.zero_()
    m.model.train()
    outs = m.model(inputs)
    loss_a = m._loss_fn([outs[0]], labels, weights)
    loss_a.backward()
    assert abs(loss_e - float(loss_a)) < 1e-6 * max(1.0, abs(loss_e))
    for n, p in m.model.named_parameters():
        scale = max(float(g_engine[n].abs().max()), 1e-6)
        assert float((p.grad - g_engine[n]).abs().max()) < 5e-5 * scale + 1e-9, n


def test_engine_adam_trajectory_matches_torch_adam():
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    pm = make_molecules(64, seed=13, shape="delaney")
    y, w = make_labels(64, 1, "regression", seed=4)
    om, m = _pair("regression", [32, 32], 64, 1, 64, bn=False)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    ds = PackedDataset(pm, y, w)
    for step in range(5):
        loss = m.fit_on_batch(pm, y, w)
        opt.zero_grad()
        oo, lo = _oracle_step(om, "regression", pm, y, w, 64)
        lo.backward()
        opt.step()
        assert abs(loss - float(lo.detach())) < 2e-5 * max(1.0, abs(float(lo.detach()))), step
    osd = om.state_dict()
    for k, v in m.model.state_dict().items():
        assert rel_err(v.cpu().numpy(), osd[k].numpy()) < 2e-4, k
    assert m.get_global_step() == 5 and m._engine.step_count == 5
    del ds


def test_engine_eval_forward_and_predict():
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.synthetic import make_molecules
    _cuda()
    pm = make_molecules(50, seed=14, shape="stress")
    om, m = _pair("classification", [64, 64], 128, 4, 64)
    with torch.no_grad():                                  # non-trivial running statistics
        for bn in om.batch_norms:
            bn.running_mean.normal_(0, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    m.model.load_state_dict(om.state_dict())
    batch = next(m.default_generator(PackedDataset(pm, n_tasks=4), mode="predict", deterministic=True,
                                     pad_batches=False))
    inputs, _, _ = m._prepare_batch(batch)
    out, probs, fp = m._engine.forward(inputs[1]._dcgc_topology, inputs[0], 50, training=False)
    _, mm = oracle_batch(pm.to_list())
    om.eval()
    oo = om(torch_args(mm, 50))
    assert rel_err(probs.cpu().numpy().reshape(50, 4, 2), oo[0].detach().numpy()) < MODEL_TOL
    assert rel_err(out.cpu().numpy().reshape(50, 4, 2), oo[1].detach().numpy()) < MODEL_TOL
    assert rel_err(fp.cpu().numpy(), oo[2].detach().numpy()) < MODEL_TOL
    assert np.all(fp.cpu().numpy()[50:, :128] == 0) and np.all(fp.cpu().numpy()[50:, 128:] == -1)


def test_state_dict_survives_flat_slab():
    """Parameters are views of one slab; checkpoints keep the reference keys and reload exactly."""
    from deepchem_b200.graphconvmodel import GraphConvModel
    _cuda()
    m = GraphConvModel(12, [64, 64], 128, mode="classification", batch_size=50)
    sd = m.model.state_dict()
    assert sum(1 for _ in sd) == 103                        # 103 tensors for the default model
    assert sum(v.numel() for k, v in sd.items() if "running" not in k and "num_batches" not in k) == 204504
    assert sd["graph_convs.0.W_list.3"].shape == (75, 64) and sd["dense.weight"].shape == (128, 64)
    m2 = GraphConvModel(12, [64, 64], 128, mode="classification", batch_size=50)
    m2.model.load_state_dict(sd)
    assert m2._engine.aliased()
    assert torch.equal(m2._engine.params, m._engine.params)
    m2.model.to("cuda:0")
    m2._engine.adopt()
    assert m2._engine.aliased()


def test_int8_feature_upload_is_exact_and_fit_is_unchanged():
    """The compact int8 copy of an integer-valued feature matrix (PackedMols.compact) converts back to exactly the
    fp32 rows on the device, so a fit through the prefetch pipeline gives bit-identical losses either way."""
    import itertools
    from deepchem_b200 import graphconvmodel as G
    from deepchem_b200 import mol_graphs as MG
    from deepchem_b200 import ops
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.synthetic import make_labels, make_molecules
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda", 0)
    pm = make_molecules(256, seed=11, shape="stress")
    pm.features[5, 40] = -1.0
    assert pm.compact()
    lay = MG.BatchLayout.build(pm, n_segments=256)
    topo = lay.to_device(dev)
    a = ops.permute_rows(torch.from_numpy(pm.features).to(dev), topo.perm)
    b = ops.permute_rows(torch.from_numpy(pm.features_i8).to(dev), topo.perm)
    assert a.shape == b.shape and torch.equal(a, b)
    assert torch.equal(a.cpu(), torch.from_numpy(pm.features[lay.perm]))
    pm.pin_memory()
    y, w = make_labels(256, 2, "regression", seed=3)
    ds = PackedDataset(pm, y, w)
    runs = []
    for use_i8 in (True, False):
        G._USE_I8 = use_i8
        try:
            torch.manual_seed(0)
            m = G.GraphConvModel(2, [64, 64], 128, mode="regression", batch_size=64, device=dev)
            m.log_frequency = 1
            losses = []
            m.fit_generator(itertools.islice(m.default_generator(ds, epochs=3, deterministic=True), 10),
                            checkpoint_interval=0, all_losses=losses)
            runs.append(losses)
        finally:
            G._USE_I8 = True
    assert len(runs[0]) == 10 and runs[0] == runs[1]


def test_shuffled_fit_lazy_gather_equals_eager_batches(monkeypatch):
    """fit(deterministic=False): batches gathered by the layout workers into pinned staging memory (int8 feature copy)
    train exactly like batches gathered eagerly by the dataset iterator (same permutations)."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    pm = make_molecules(300, seed=15, shape="stress")
    assert pm.compact()
    pm.pin_memory()
    y, w = make_labels(300, 2, "regression", seed=5)
    finals = []
    for lazy in ("1", "0"):
        monkeypatch.setenv("DCGC_LAZY_TAKE", lazy)
        torch.manual_seed(3)
        np.random.seed(7)
        m = GraphConvModel(2, [64, 64], 128, mode="regression", batch_size=64)
        m.fit(PackedDataset(pm, y, w), nb_epoch=3, deterministic=False)
        finals.append({k: v.detach().cpu().clone() for k, v in m.model.state_dict().items()})
    for k in finals[0]:
        assert torch.equal(finals[0][k], finals[1][k]), k


def test_uncertainty_mode_against_reference_outputs_and_loss():
    """Rows a11 / a13 on the CUDA path: the five outputs [y, exp(log_var), y, log_var, fingerprint], the uncertainty
    loss, predict_uncertainty and predict_on_generator against the fixture the reference's own
    GraphConvModel(uncertainty=True) produced (tests/golden/make_golden_uncertainty.py)."""
    from helpers import load_golden, unpack_mols
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import PackedMols
    _cuda()
    d = load_golden("ref_model_uncertainty.npz")
    mols = unpack_mols(d)
    n, bsz = len(mols), int(d["batch_size"])
    pm = PackedMols.from_list(mols)
    ds = PackedDataset(pm, d["y"], d["w"])
    sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
    for mode in ("tf32x3", "fp32"):
        m = GraphConvModel(3, [64, 64], 128, mode="regression", uncertainty=True, dropout=0.25, batch_size=bsz,
                           gemm_mode=mode)
        assert m._engine is None and m.output_types == ['prediction', 'variance', 'loss', 'loss', 'embedding']
        m.model.load_state_dict(sd)
        batch = next(m.default_generator(ds, deterministic=True, pad_batches=False))
        inputs, labels, weights = m._prepare_batch(batch)
        m.model.train()
        outs = m.model(inputs)
        assert len(outs) == 5
        for i, r in enumerate(outs):
            assert rel_err(r.detach().cpu().numpy(), d["ref_train_out%d" % i]) < MODEL_TOL, (mode, i)
        loss = m._loss_fn([outs[i] for i in m._loss_outputs], labels, weights)
        ref = float(d["ref_train_loss"])
        assert abs(float(loss) - ref) < 2e-5 * max(1.0, abs(ref)), mode
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.model.parameters())
        # (the fixture's eval-mode outputs were taken after its one train-mode pass: same running statistics here)
        pred, std = m.predict_uncertainty(ds)
        assert rel_err(pred, d["ref_eval_out0"]) < MODEL_TOL and rel_err(std, np.sqrt(d["ref_eval_out1"])) < MODEL_TOL
        assert rel_err(m.predict(ds), d["ref_eval_out0"]) < MODEL_TOL
        gen = m.default_generator(ds, mode='predict', deterministic=True, pad_batches=False)
        var = m.predict_on_generator(gen, output_types='variance')
        assert rel_err(var, d["ref_eval_out1"]) < MODEL_TOL

        class Shift(object):                               # a y-transformer: predictions come back untransformed
            transform_y = True

            def untransform(self, y):
                return y * 2.0 + 1.0
        assert rel_err(m.predict(ds, [Shift()]), d["ref_eval_out0"] * 2.0 + 1.0) < MODEL_TOL


def test_checkpoints_interchange_between_engine_and_autograd_models(tmp_path):
    """'optimizer_state_dict' is torch.optim.Adam's layout on both paths, so a checkpoint written by
    the fused engine restores into the per-layer autograd model and back, with the Adam moments and step count."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    pm = make_molecules(64, seed=4, shape="delaney")
    y, w = make_labels(64, 1, "regression", seed=4)
    ds = PackedDataset(pm, y, w)
    kw = dict(mode="regression", batch_size=32, learning_rate=0.01)
    torch.manual_seed(0)
    a = GraphConvModel(1, [32, 32], 64, model_dir=str(tmp_path / "a"), **kw)
    assert a._engine is not None
    a.fit(ds, nb_epoch=3, deterministic=True)
    # engine checkpoint -> autograd model
    b = GraphConvModel(1, [32, 32], 64, model_dir=str(tmp_path / "a"), use_engine=False, **kw)
    assert b._engine is None
    b.restore()
    assert b.get_global_step() == a.get_global_step()
    st = b._pytorch_optimizer.state_dict()["state"]
    sa = a._engine.state_dict()["state"]
    assert len(st) == len(sa) > 0
    for i in sa:
        assert torch.equal(st[i]["exp_avg"].cpu(), sa[i]["exp_avg"].cpu()) and float(st[i]["step"]) == float(sa[i]["step"])
    # both continue identically for one more epoch (same arithmetic up to fp32 summation order)
    a.fit(ds, nb_epoch=1, deterministic=True, checkpoint_interval=0)
    b.model_dir = str(tmp_path / "b")
    b.fit(ds, nb_epoch=1, deterministic=True)
    assert rel_err(b.predict(ds), a.predict(ds)) < 1e-3
    # autograd checkpoint -> engine model: moments and step arrive
    c = GraphConvModel(1, [32, 32], 64, model_dir=str(tmp_path / "b"), **kw)
    c.restore()
    sb = b._pytorch_optimizer.state_dict()["state"]
    sc = c._engine.state_dict()["state"]
    for i in sb:
        assert torch.equal(sc[i]["exp_avg"].cpu(), sb[i]["exp_avg"].cpu())
    assert c._engine.step_count == int(float(sb[0]["step"]))
    assert np.array_equal(c.predict(ds), b.predict(ds)) or rel_err(c.predict(ds), b.predict(ds)) < 1e-5


def test_materialised_generator_survives_the_staging_ring():
    """batches = list(model.default_generator(ds)) holds more batches than the pinned ring; fitting on
    the list gives the same parameters as fitting on the generator (no batch was overwritten behind the caller)."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    pm = make_molecules(40 * 8, seed=12, shape="delaney")
    y, w = make_labels(40 * 8, 1, "regression", seed=12)
    ds = PackedDataset(pm, y, w)

    def model():
        torch.manual_seed(3)
        return GraphConvModel(1, [32], 32, mode="regression", batch_size=8, batch_normalize=False)
    m1, m2 = model(), model()
    m2.model.load_state_dict(m1.model.state_dict())
    m1._staging.max_slots = m2._staging.max_slots = 6
    batches = list(m1.default_generator(ds, deterministic=True))
    assert len(batches) == 40
    fresh = list(m2.default_generator(ds, deterministic=True, workers=1))
    for (ia, _, _), (ib, _, _) in zip(batches, fresh):
        assert np.array_equal(ia.layout.membership, ib.layout.membership)
        assert np.array_equal(ia.layout.col_idx, ib.layout.col_idx)
    del fresh
    m1.fit_generator(batches, checkpoint_interval=0)
    m2.fit_generator(m2.default_generator(ds, deterministic=True), checkpoint_interval=0)
    for (k, v1), (_, v2) in zip(m1.model.state_dict().items(), m2.model.state_dict().items()):
        assert torch.equal(v1, v2), k


def test_exact_input_path_is_bit_identical():
    """dcgc_gcmodel_config.input_exact: with integer-valued features the first layer's operands are exact in tf32, the
    lo(A) tile is identically zero and the TF32x3 GEMMs skip its term — every gradient must be bit-identical to the
    full three-term evaluation."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    _cuda()
    B = 600
    y, w = make_labels(B, 2, "regression", seed=3)
    res = []
    for compact in (False, True):
        pm = make_molecules(B, seed=17, shape="stress")
        if compact:
            pm = pm.pin_memory()
            assert pm.features_i8 is not None
        torch.manual_seed(0)
        m = GraphConvModel(2, [128, 64], 128, mode="regression", batch_size=B, gemm_mode="tf32x3")
        batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
        inputs, labels, weights = m._prepare_batch(batch)
        assert bool(getattr(inputs[0], "_dcgc_input_exact", False)) == compact
        loss = m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0], weights[0], B)
        assert m._engine.cfg.input_exact == int(compact)
        res.append((float(loss), m._engine.grads.clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])


def test_predict_through_the_staged_result_ring_equals_the_direct_copy(monkeypatch):
    """Large predict() results go through a small ring of reusable pinned buffers + copier threads into pageable
    arrays (cfg5: 1.28 GB of probabilities per GPU); the numbers must be those of the direct single-array path."""
    from deepchem_b200 import graphconvmodel as G
    from deepchem_b200.data import PackedDataset, ReplayDataset
    from deepchem_b200.synthetic import make_molecules
    _cuda()
    shard = make_molecules(700, seed=31, shape="pcba").pin_memory()
    torch.manual_seed(0)
    m = G.GraphConvModel(16, [64, 64], 128, mode="classification", n_classes=2, batch_size=256)
    ds = ReplayDataset(shard, 3000)
    direct = m.predict(ds)
    monkeypatch.setattr(G, "_PINNED_DIRECT_LIMIT", 1 << 10)
    monkeypatch.setattr(G._StagedResult, "BUF_BYTES", 1 << 16)       # several seals per pass
    staged = m.predict(ds)
    assert staged.shape == direct.shape == (3000, 16, 2)
    assert np.array_equal(staged, direct)
    # rank shards of the replayed stream reassemble the whole pass
    parts = [m.predict(ds, shard=(r, 3)) for r in range(3)]
    assert np.array_equal(np.concatenate(parts, axis=0), direct)
    # and the stream really replays the shard
    assert np.array_equal(direct[:700], m.predict(PackedDataset(shard)))
    assert np.array_equal(direct[700:1400], direct[:700])

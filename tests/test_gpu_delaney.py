"""BASELINE config 1 on the real molecules: GraphConvModel regression over the Delaney (ESOL) set read by the RDKit-free
SMILES reader — the CUDA path against the outputs the reference produced on the same molecules
(tests/golden/ref_delaney_real.npz), and a fit over the whole dataset through the public API."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden, rel_err

pytestmark = pytest.mark.gpu
MODEL_TOL = 1e-4     # composite tolerance, see tests/test_gpu_parity.py


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_real_molecules_against_reference_outputs():
    from deepchem_b200.data import CSVLoader
    from deepchem_b200.graphconvmodel import GraphConvModel
    _cuda()
    d = load_golden("ref_delaney_real.npz")
    n = int(d["batch_size"])
    ds = CSVLoader(["y"]).create_dataset(os.path.join(GOLDEN, "delaney.csv")).select_range(0, n)
    assert np.array_equal(ds.X.features, d["features"]) and np.array_equal(ds.y, d["y"])
    sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
    for mode in ("tf32x3", "fp32"):
        m = GraphConvModel(1, [64, 64], 128, mode="regression", batch_size=n, gemm_mode=mode)
        m.model.load_state_dict(sd)
        batch = next(m.default_generator(ds, deterministic=True, pad_batches=False))
        inputs, labels, weights = m._prepare_batch(batch)
        m.model.train()                                   # the fixture ran train mode first: it moves the running statistics
        for i, r in enumerate(m.model(inputs)):
            assert rel_err(r.detach().cpu().numpy(), d["ref_train_out%d" % i]) < MODEL_TOL, (mode, i)
        m.model.eval()
        for i, r in enumerate(m.model(inputs)):
            assert rel_err(r.detach().cpu().numpy(), d["ref_eval_out%d" % i]) < MODEL_TOL, (mode, i)
        m.model.load_state_dict(sd)                       # undo the running-statistics update
        assert m._engine is not None
        loss = m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(),
                                    weights[0].contiguous(), n)
        ref = float(d["ref_train_loss"])
        assert abs(float(loss) - ref) < 2e-5 * abs(ref), mode


def test_fit_whole_dataset():
    """1 128 molecules, batch 100 (12 batches, the last padded), [64, 64] + dense 128: training error falls well below
    the label variance (the reference's own Delaney test asserts the model can fit: models/tests/test_graph_models.py)."""
    from deepchem_b200.data import CSVLoader
    from deepchem_b200.graphconvmodel import GraphConvModel
    _cuda()
    ds = CSVLoader(["y"]).create_dataset(os.path.join(GOLDEN, "delaney.csv"))
    torch.manual_seed(0)
    m = GraphConvModel(1, [64, 64], 128, mode="regression", batch_size=100, learning_rate=1e-3)
    var = float(np.var(ds.y))
    m.fit(ds, nb_epoch=1, deterministic=True)
    first = float(np.mean((m.predict(ds) - ds.y) ** 2))
    m.fit(ds, nb_epoch=40)
    pred = m.predict(ds)
    assert pred.shape == (1128, 1) and np.all(np.isfinite(pred))
    mse = float(np.mean((pred - ds.y) ** 2))
    assert mse < 0.35 * var and mse < first, (mse, first, var)

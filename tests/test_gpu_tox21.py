"""BASELINE config 2 on the real molecules: GraphConvModel 12-task classification over the reference's Tox21 file read
by the RDKit-free SMILES reader — the CUDA path against the outputs the reference produced on rows 0..49
(tests/golden/ref_tox21_real.npz: probabilities, logits, fingerprints, weighted softmax cross-entropy with zero weights
on the missing labels), and an epoch over all 8 014 molecules through the public API."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden, rel_err

pytestmark = pytest.mark.gpu
MODEL_TOL = 1e-4     # composite tolerance, see tests/test_gpu_parity.py
TASKS = ['NR-AR', 'NR-AR-LBD', 'NR-AhR', 'NR-Aromatase', 'NR-ER', 'NR-ER-LBD', 'NR-PPAR-gamma', 'SR-ARE', 'SR-ATAD5',
         'SR-HSE', 'SR-MMP', 'SR-p53']


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


@pytest.fixture(scope="module")
def tox21():
    from deepchem_b200.data import CSVLoader
    return CSVLoader(TASKS).create_dataset(os.path.join(GOLDEN, "tox21.csv.gz"))


def test_real_molecules_against_reference_outputs(tox21):
    from deepchem_b200.graphconvmodel import GraphConvModel
    _cuda()
    d = load_golden("ref_tox21_real.npz")
    n = int(d["batch_size"])
    ds = tox21.select_range(0, n)
    assert np.array_equal(ds.X.features, d["features"]) and np.array_equal(ds.y, d["y"]) and np.array_equal(ds.w, d["w"])
    sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
    for mode in ("tf32x3", "fp32"):
        m = GraphConvModel(12, [64, 64], 128, mode="classification", n_classes=2, batch_size=n, gemm_mode=mode)
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


def test_fit_whole_dataset(tox21):
    """8 014 molecules, batch 50 (161 batches, the last padded with zero weights): the weighted cross-entropy ends well
    below ln 2, predictions are probabilities over 2 classes for every molecule and task."""
    from deepchem_b200.graphconvmodel import GraphConvModel
    _cuda()
    torch.manual_seed(0)
    m = GraphConvModel(12, [64, 64], 128, mode="classification", n_classes=2, batch_size=50, learning_rate=1e-3)
    first = m.fit(tox21, nb_epoch=1, deterministic=True)
    last = m.fit(tox21, nb_epoch=3)
    pred = m.predict(tox21)
    assert pred.shape == (8014, 12, 2) and np.all(np.isfinite(pred))
    assert np.allclose(pred.sum(-1), 1.0, atol=1e-5)
    y1 = np.eye(2, dtype=np.float32)[tox21.y.astype(np.int64)]
    ce = float((-(y1 * np.log(np.clip(pred, 1e-12, 1.0))).sum(-1) * tox21.w).sum() / tox21.w.sum())
    assert ce < 0.45, (ce, first, last)      # untrained: ln 2 = 0.69; the base rate (7.5 % actives) alone gives 0.27

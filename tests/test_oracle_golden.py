"""The oracle (oracle/) pinned against the reference's known-answer vectors and against
reference outputs generated in the build container (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import load_golden, oracle_batch, rel_err, torch_args, unpack_mols
from oracle import graphconv_torch as O
from oracle.convmol_layout import derived_topology


def _check_layout(d, cms, mm):
    assert mm.deg_slice.dtype == d["ref_deg_slice"].dtype == np.int64
    assert np.array_equal(mm.deg_slice, d["ref_deg_slice"])
    assert mm.membership.dtype == d["ref_membership"].dtype == np.int32
    assert np.array_equal(mm.membership, d["ref_membership"])
    for k, a in enumerate(mm.get_deg_adjacency_lists()):
        r = d["ref_deg_adj_%d" % k]
        assert a.dtype == r.dtype == np.int32 and a.shape == r.shape
        assert np.array_equal(a, r)
    assert np.array_equal(mm.get_atom_features(), d["ref_nodes"])
    # per-molecule ConvMol outputs
    ds = np.stack([c.deg_slice for c in cms])
    assert ds.dtype == d["ref_mol_deg_slice"].dtype and np.array_equal(ds, d["ref_mol_deg_slice"])
    assert np.array_equal(np.concatenate([c.deg_block_indices for c in cms]), d["ref_mol_deg_block_indices"])
    assert np.array_equal(np.concatenate([np.asarray(c.degree_list, np.int32) for c in cms]),
                          d["ref_mol_degree_list"])
    flat = np.asarray([k for c in cms for nb in c.get_adjacency_list() for k in nb], np.int32)
    assert np.array_equal(flat, d["ref_mol_canon_adj_flat"])
    assert np.array_equal(np.concatenate([c.get_atom_features() for c in cms]), d["ref_mol_features_sorted"])


@pytest.mark.parametrize("name", ["kat_ccc_c.npz", "ref_layout_stress.npz", "ref_layout_zinc.npz",
                                  "ref_layout_delaney.npz"])
def test_layout_oracle_matches_reference(name):
    d = load_golden(name)
    cms, mm = oracle_batch(unpack_mols(d))
    _check_layout(d, cms, mm)


def test_layout_kat_values():
    """Observed reference values for ['CCC','C']."""
    d = load_golden("kat_ccc_c.npz")
    _, mm = oracle_batch(unpack_mols(d))
    assert mm.deg_slice[:4].tolist() == [[0, 1], [1, 2], [3, 1], [4, 0]]
    assert mm.membership.tolist() == [1, 0, 0, 0]
    assert [a.shape for a in mm.get_deg_adjacency_lists()][:4] == [(1, 0), (2, 1), (1, 2), (0, 3)]


def test_layout_null_mol():
    """One atom of every degree, bonded to itself (mol_graphs.py:236-254), twice."""
    from oracle.convmol_layout import OracleConvMol, agglomerate
    d = load_golden("ref_layout_nullmol.npz")
    adj = [deg * [deg] for deg in range(11)]
    nm = OracleConvMol(d["features"], adj)
    mm = agglomerate([nm, nm])
    assert np.array_equal(mm.deg_slice, d["deg_slice"])
    assert np.array_equal(mm.membership, d["membership"])
    for k, a in enumerate(mm.get_deg_adjacency_lists()):
        assert np.array_equal(a, d["deg_adj_%d" % k])


def test_derived_topology_consistency():
    d = load_golden("ref_layout_stress.npz")
    _, mm = oracle_batch(unpack_mols(d))
    t = derived_topology(mm.deg_slice, mm.membership, mm.get_deg_adjacency_lists(), mm.num_mols)
    n = mm.get_num_atoms()
    # CSR rows reproduce deg_adj lists
    for deg, a in enumerate(mm.get_deg_adjacency_lists()):
        s = int(mm.deg_slice[deg, 0])
        for r in range(a.shape[0]):
            lo, hi = t["row_ptr"][s + r], t["row_ptr"][s + r + 1]
            assert t["col_idx"][lo:hi].tolist() == a[r].tolist()
    # transposed CSR is the exact transpose with slots
    for j in range(n):
        for e in range(t["t_row_ptr"][j], t["t_row_ptr"][j + 1]):
            i, k = t["t_src"][e], t["t_slot"][e]
            assert t["col_idx"][t["row_ptr"][i] + k] == j
    assert t["t_row_ptr"][-1] == t["row_ptr"][-1]
    for g in range(mm.num_mols):
        rows = t["mol_atoms"][t["mol_ptr"][g]:t["mol_ptr"][g + 1]]
        assert np.all(mm.membership[rows] == g) and np.all(np.diff(rows) > 0)


def test_layer_kats_against_reference_assets():
    """models/tests/test_layers.py:1458-1546 replayed on the oracle (their tolerance: 1e-4)."""
    d = load_golden("kat_ccc_c.npz")
    _, mm = oracle_batch(unpack_mols(d))
    args = torch_args(mm)
    W = [torch.from_numpy(w) for w in d["asset_graphconvlayer_weights"]]
    b = [torch.from_numpy(w) for w in d["asset_graphconvlayer_biases"]]
    y = O.graph_conv(args[0], args[1], args[3:], W, b)
    assert np.abs(y.numpy() - d["asset_graphconvlayer_result"]).max() < 1e-6
    p = O.graph_pool(args[0], args[1], args[3:])
    assert np.array_equal(p.numpy(), d["asset_graphpoollayer_result"])
    g = O.graph_gather(args[0], args[2], 2)
    assert np.array_equal(g.numpy(), d["asset_graphgatherlayer_result"])


def _load_asset_model(d):
    m = O.OracleGraphConvModel(2, [64, 64], 128, mode="classification", batch_normalize=False, batch_size=10)
    with torch.no_grad():
        for i in (0, 1):
            for k in range(21):
                m.graph_convs[i].W_list[k].copy_(torch.from_numpy(d["asset_graphconvlayer%d_weights" % i][k]))
                m.graph_convs[i].b_list[k].copy_(torch.from_numpy(d["asset_graphconvlayer%d_biases" % i][k]))
        m.dense.weight.copy_(torch.from_numpy(d["asset_dense_weights"].T))
        m.dense.bias.copy_(torch.from_numpy(d["asset_dense_biases"]))
        m.reshape_dense.weight.copy_(torch.from_numpy(d["asset_reshapedense_weights"].T))
        m.reshape_dense.bias.copy_(torch.from_numpy(d["asset_reshapedense_biases"]))
    return m


def test_model_kat_against_reference_assets():
    """models/tests/test_graphconv_torchmodel.py:15-95 replayed on the oracle model."""
    d = load_golden("kat_ccc_c.npz")
    _, mm = oracle_batch(unpack_mols(d))
    m = _load_asset_model(d)
    out = m(torch_args(mm, 2))
    assert np.abs(out[0].detach().numpy() - d["asset_graphconvmodel_output_classification"]).max() < 2e-6
    assert np.abs(out[1].detach().numpy() - d["asset_graphconvmodel_logits_classification"]).max() < 2e-6
    fp = out[2].detach().numpy()
    assert fp.shape == (10, 256)
    assert np.abs(fp - d["asset_graphconvmodel_neural_classification"]).max() < 2e-6
    # empty segments: 0 in the sum half, tanh(-inf) = -1 in the max half
    assert np.all(fp[2:, :128] == 0) and np.all(fp[2:, 128:] == -1)


def test_layers_against_reference_outputs():
    d = load_golden("ref_layers.npz")
    _, mm = oracle_batch(unpack_mols(d))
    args = torch_args(mm)
    W = [torch.from_numpy(w) for w in d["W"]]
    b = [torch.from_numpy(w) for w in d["b"]]
    assert rel_err(O.graph_conv(args[0], args[1], args[3:], W, b, torch.relu).numpy(), d["ref_conv_relu"]) < 1e-6
    assert rel_err(O.graph_conv(args[0], args[1], args[3:], W, b).numpy(), d["ref_conv_linear"]) < 1e-6
    assert np.array_equal(O.graph_pool(args[0], args[1], args[3:]).numpy(), d["ref_pool"])
    bsz = int(d["gather_batch_size"])
    assert rel_err(O.graph_gather(args[0], args[2], bsz, torch.tanh).numpy(), d["ref_gather_tanh"]) < 1e-6
    g = O.graph_gather(args[0], args[2], bsz).numpy()
    assert rel_err(g, d["ref_gather_linear"]) < 1e-6
    assert np.all(np.isneginf(g[-3:, 75:])) and np.all(g[-3:, :75] == 0)


@pytest.mark.parametrize("mode", ["classification", "regression"])
def test_model_against_reference_outputs(mode):
    d = load_golden("ref_model_%s.npz" % mode)
    mols = unpack_mols(d)
    _, mm = oracle_batch(mols)
    m = O.OracleGraphConvModel(3, [64, 64], 128, mode=mode, batch_size=int(d["batch_size"]))
    sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
    assert set(sd) == set(m.state_dict())          # checkpoint keys identical to the reference
    m.load_state_dict(sd)
    args = torch_args(mm, len(mols))
    m.train()
    out = m(args)
    for i, r in enumerate(out):
        assert rel_err(r.detach().numpy(), d["ref_train_out%d" % i]) < 2e-6, i
    loss = O.standard_loss(mode, out, torch.from_numpy(d["y"]), torch.from_numpy(d["w"]))
    assert abs(float(loss) - float(d["ref_train_loss"])) < 1e-6 * max(1.0, abs(float(d["ref_train_loss"])))
    for k, v in m.state_dict().items():
        if "running" in k:
            assert rel_err(v.numpy(), d["sd_after:" + k]) < 1e-6, k
    m.eval()
    out = m(args)
    for i, r in enumerate(out):
        assert rel_err(r.detach().numpy(), d["ref_eval_out%d" % i]) < 2e-6, i


def test_uncertainty_model_against_reference_outputs_and_loss():
    """Rows a11 / a13: [y, exp(log_var), y, log_var, fingerprint] and the reference's uncertainty loss closure
    (graphconvmodel.py:238-246, 360-372), fixture generated by tests/golden/make_golden_uncertainty.py."""
    d = load_golden("ref_model_uncertainty.npz")
    mols = unpack_mols(d)
    _, mm = oracle_batch(mols)
    m = O.OracleGraphConvModel(3, [64, 64], 128, mode="regression", uncertainty=True, dropout=0.25,
                               batch_size=int(d["batch_size"]))
    sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
    assert set(sd) == set(m.state_dict())
    m.load_state_dict(sd)
    args = torch_args(mm, len(mols))
    m.train()
    out = m(args)
    assert len(out) == 5
    for i, r in enumerate(out):
        assert rel_err(r.detach().numpy(), d["ref_train_out%d" % i]) < 2e-6, i
    loss = O.standard_loss("regression", out, torch.from_numpy(d["y"]), torch.from_numpy(d["w"]), uncertainty=True)
    assert abs(float(loss.detach()) - float(d["ref_train_loss"])) < 1e-6 * max(1.0, abs(float(d["ref_train_loss"])))
    m.eval()
    for i, r in enumerate(m(args)):
        assert rel_err(r.detach().numpy(), d["ref_eval_out%d" % i]) < 2e-6, i


def test_segment_max_ties_first_row_and_empty():
    x = torch.tensor([[1., 5.], [3., 5.], [3., 2.], [0., 0.]], dtype=torch.float64, requires_grad=True)
    ids = torch.tensor([0, 0, 0, 2])
    out = O.segment_max(x, ids, 4)
    assert out[0].tolist() == [3., 5.] and torch.isneginf(out[1]).all() and torch.isneginf(out[3]).all()
    out[torch.isfinite(out)].sum().backward()
    assert x.grad.tolist() == [[0., 1.], [1., 0.], [0., 0.], [1., 1.]]


def test_pool_ties_route_to_self_first():
    x = torch.zeros(3, 2, dtype=torch.float64, requires_grad=True)          # all tie
    deg_slice = torch.tensor([[0, 0], [0, 2], [2, 1]] + [[3, 0]] * 8)
    adjs = [torch.tensor([[2], [2]]), torch.tensor([[0, 1]])] + [torch.zeros(0, k, dtype=torch.long) for k in range(3, 11)]
    O.graph_pool(x, deg_slice, adjs).sum().backward()
    assert x.grad.tolist() == [[1., 1.], [1., 1.], [1., 1.]]


def test_oracle_gradcheck_fp64():
    """No reference test checks a gradient of this path; the oracle's autograd
    is validated against finite differences in float64."""
    d = load_golden("ref_layout_stress.npz")
    mols = unpack_mols(d)[:6]
    _, mm = oracle_batch(mols)
    g = torch.Generator().manual_seed(0)
    n = mm.get_num_atoms()
    x = torch.randn(n, 5, dtype=torch.float64, generator=g, requires_grad=True)
    args = torch_args(mm)
    W = [torch.randn(5, 3, dtype=torch.float64, generator=g, requires_grad=True) for _ in range(21)]
    b = [torch.randn(3, dtype=torch.float64, generator=g, requires_grad=True) for _ in range(21)]

    def f(x, *wb):
        y = O.graph_conv(x, args[1], args[3:], list(wb[:21]), list(wb[21:]), torch.tanh)
        y = O.graph_pool(y, args[1], args[3:])
        return O.graph_gather(y, args[2], len(mols), torch.tanh)
    assert torch.autograd.gradcheck(f, (x, *W, *b), eps=1e-6, atol=1e-5)


def test_segment_ops_reproduce_the_reference_docstring_examples():
    """The worked examples in the reference's own docstrings (utils/pytorch_utils.py:38-53 and :490-506): segment ids
    [0, 1, 0] over three rows."""
    ids = torch.tensor([0, 1, 0])
    data = torch.tensor([[1., 2., 3., 4.], [5., 6., 7., 8.], [4., 3., 2., 1.]])
    assert O.segment_sum(data, ids, 2).tolist() == [[5., 5., 5., 5.], [5., 6., 7., 8.]]
    assert O.segment_max(data, ids, 2).tolist() == [[4., 3., 3., 4.], [5., 6., 7., 8.]]

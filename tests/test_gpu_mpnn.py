"""MPNN edge-network layers on the GPU (deepchem_b200/mpnn.py, csrc/mpnn_kernels.cu + tcgen05 GEMMs) against the
reference-generated goldens and the CPU oracle (oracle/mpnn_torch.py).  fp32 tolerance 1e-5 of the tensor scale per
layer (TF32x3 GEMMs are fp32-grade); the bilinear factorisation of EdgeNetwork changes only the summation order."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err
from oracle import mpnn_torch as M

pytestmark = pytest.mark.gpu
T = torch.from_numpy
TOL = 1e-5


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _all_pairs(sizes):
    a2p, start = [], 0
    for n in sizes:
        C0, C1 = np.meshgrid(np.arange(n), np.arange(n))
        a2p.append(np.transpose(np.array([C1.flatten() + start, C0.flatten() + start])))
        start += n
    return np.concatenate(a2p, axis=0).astype(np.int64)


def test_edge_network_matches_reference_outputs():
    from deepchem_b200.mpnn import EdgeNetwork
    _cuda()
    d = load_golden("ref_mpnn.npz")
    for tag, P, h in (("a", 14, 32), ("b", 8, 75)):
        layer = EdgeNetwork(P, h)
        layer.W = T(d["edge_%s_W" % tag])
        layer.b = T(d["edge_%s_b" % tag])
        out = layer([T(d["edge_%s_pf" % tag]), T(d["edge_%s_x" % tag]), T(d["edge_%s_a2p" % tag])])
        assert out.is_cuda and tuple(out.shape) == d["edge_%s_out" % tag].shape
        assert rel_err(out.cpu().numpy(), d["edge_%s_out" % tag]) < TOL, tag


def test_edge_network_weave_sized_batch_matches_oracle():
    """64 molecules of ~25 atoms, all n^2 pairs each, 14 pair features, hidden 100 (the MPNNModel defaults)."""
    from deepchem_b200.mpnn import EdgeNetwork
    _cuda()
    rng = np.random.default_rng(3)
    sizes = rng.integers(6, 40, size=64).tolist()
    a2p = _all_pairs(sizes)
    n, P, h = sum(sizes), 14, 100
    pf = ((rng.random((a2p.shape[0], P)) < 0.25).astype(np.float32))
    x = rng.standard_normal((n, h)).astype(np.float32)
    torch.manual_seed(3)                      # the weight initialiser draws from torch's generator
    layer = EdgeNetwork(P, h)
    layer.b = T(rng.standard_normal(h * h).astype(np.float32) * 0.02)
    out = layer([T(pf), T(x), T(a2p)])
    # float64 oracle in chunks of molecules (the reference materialises [pairs, h, h])
    want, want32, p0, a0 = [], [], 0, 0
    for s in sizes:
        sl = slice(p0, p0 + s * s)
        loc = a2p[sl] - a0
        want.append(M.edge_network(T(pf[sl]).double(), T(x[a0:a0 + s]).double(), T(loc), layer.W.double(),
                                   layer.b.double(), h))
        want32.append(M.edge_network(T(pf[sl]), T(x[a0:a0 + s]), T(loc), layer.W, layer.b, h))
        p0 += s * s
        a0 += s
    want, want32 = torch.cat(want).numpy(), torch.cat(want32).numpy()
    # sums of ~1500 x 25 fp32 products: the bar is the float64 result, and being at least as close to it as the
    # reference's own fp32 arithmetic is (x3 slack), as for the full-size GraphConv tests
    err, err_ref32 = rel_err(out.cpu().numpy(), want), rel_err(want32, want)
    assert err < max(TOL, 3 * err_ref32), (err, err_ref32)
    assert err < 3e-5
    with pytest.raises(AssertionError, match="sorted"):
        layer([T(pf), T(x), T(a2p[::-1].copy())])


def test_gru_matches_reference_outputs():
    from deepchem_b200.mpnn import GatedRecurrentUnit
    _cuda()
    d = load_golden("ref_mpnn.npz")
    g = GatedRecurrentUnit(32)
    for k in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh"):
        setattr(g, k, T(d["gru_" + k]))
    out = g([T(d["gru_h"]), T(d["gru_x"])])
    assert rel_err(out.cpu().numpy(), d["gru_out"]) < TOL


def test_setgather_known_answer_and_reference_outputs():
    from deepchem_b200.mpnn import SetGather
    _cuda()
    d = load_golden("ref_mpnn.npz")
    sg = SetGather(2, 2, 4)
    sg.U = torch.nn.Parameter(T(d["kat_setgather_U"]))
    out = sg([d["kat_setgather_atom_feat"], np.array([0, 0, 1, 1], dtype=np.int32)])
    # the reference's own tolerance for its TensorFlow vector (models/tests/test_layers.py:1016)
    assert np.allclose(out.cpu().numpy(), d["kat_setgather_result"], atol=1e-4)
    assert rel_err(out.cpu().numpy(), d["kat_setgather_replay"]) < TOL
    sg = SetGather(3, 5, 16)
    sg.U = torch.nn.Parameter(T(d["sg_U"]))
    sg.b = torch.nn.Parameter(T(d["sg_b"]))
    out = sg([d["sg_atom_feat"], d["sg_split"]])
    assert tuple(out.shape) == (5, 32) and rel_err(out.cpu().numpy(), d["sg_out"]) < TOL


def test_setgather_with_an_empty_molecule_and_large_batch_matches_oracle():
    from deepchem_b200.mpnn import SetGather
    _cuda()
    rng = np.random.default_rng(5)
    sizes = rng.integers(1, 60, size=200)
    sizes[17] = 0                                         # an absent molecule in the middle of the batch
    split = np.repeat(np.arange(200), sizes).astype(np.int32)
    h = 100
    x = rng.standard_normal((split.shape[0], h)).astype(np.float32) * 0.3
    sg = SetGather(4, 200, h)
    out = sg([x, split])
    want = M.set_gather(x.astype(np.float64), split, sg.U.detach(), sg.b.detach(), 4, 200, h)   # (the LSTM step is fp32 in the reference)
    assert rel_err(out.cpu().numpy(), want.numpy()) < TOL
    assert float(out[17, h:].abs().max()) == 0.0


def test_message_passing_matches_oracle_composition():
    from deepchem_b200.mpnn import MessagePassing
    _cuda()
    rng = np.random.default_rng(9)
    sizes = [5, 9, 3, 12, 7]
    a2p = _all_pairs(sizes)
    n, P, h, F = sum(sizes), 14, 64, 40
    pf = rng.random((a2p.shape[0], P)).astype(np.float32)
    x = rng.standard_normal((n, F)).astype(np.float32)
    mp = MessagePassing(3, n_hidden=h)
    out = mp([T(x), T(pf), T(a2p)])
    enn, gru = mp.message_function, mp.update_function
    want = M.message_passing(T(x).double(), T(pf).double(), T(a2p), 3, h, (enn.W.double(), enn.b.double()),
                             [getattr(gru, k).double() for k in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh")])
    assert tuple(out.shape) == (n, h)
    assert rel_err(out.cpu().numpy(), want.numpy()) < 5e-5       # three chained layers
    with pytest.raises(ValueError, match="Too large"):
        MessagePassing(1, n_hidden=8)([torch.zeros(3, 9), torch.zeros(9, 2), T(_all_pairs([3]))])

"""The MPNN edge-network oracle (oracle/mpnn_torch.py) against the reference: its own SetGather known answer and
outputs of the reference layers generated in the build container (tests/golden/make_golden_mpnn.py)."""
import numpy as np
import torch

from helpers import load_golden
from oracle import mpnn_torch as M

T = torch.from_numpy


def test_setgather_known_answer_of_the_reference():
    d = load_golden("ref_mpnn.npz")
    b = torch.cat((torch.zeros(4), torch.ones(4), torch.zeros(4), torch.zeros(4)))
    out = M.set_gather(d["kat_setgather_atom_feat"], np.array([0, 0, 1, 1], np.int32), T(d["kat_setgather_U"]), b, 2, 2, 4)
    # the reference's own tolerance for this vector (models/tests/test_layers.py:1016): TensorFlow result, atol 1e-4
    assert np.allclose(out.numpy(), d["kat_setgather_result"], atol=1e-4)
    # and the torch reference replayed here: same arithmetic
    assert np.abs(out.numpy() - d["kat_setgather_replay"]).max() < 1e-6


def test_edge_network_against_reference_outputs():
    d = load_golden("ref_mpnn.npz")
    for tag, h in (("a", 32), ("b", 75)):
        out = M.edge_network(T(d["edge_%s_pf" % tag]), T(d["edge_%s_x" % tag]), T(d["edge_%s_a2p" % tag]),
                             T(d["edge_%s_W" % tag]), T(d["edge_%s_b" % tag]), h)
        ref = d["edge_%s_out" % tag]
        assert out.shape == ref.shape and np.abs(out.numpy() - ref).max() <= 1e-6 * np.abs(ref).max()


def test_gru_and_setgather_against_reference_outputs():
    d = load_golden("ref_mpnn.npz")
    w = [T(d["gru_" + k]) for k in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh")]
    out = M.gated_recurrent_unit(T(d["gru_h"]), T(d["gru_x"]), *w)
    assert np.abs(out.numpy() - d["gru_out"]).max() <= 1e-6
    sg = M.set_gather(d["sg_atom_feat"], d["sg_split"], T(d["sg_U"]), T(d["sg_b"]), 3, 5, 16)
    assert sg.shape == (5, 32) and np.abs(sg.numpy() - d["sg_out"]).max() <= 1e-6


def test_message_passing_pads_and_rejects_wide_inputs():
    rng = np.random.default_rng(0)
    h, P = 8, 3
    a2p = torch.tensor([[0, 0], [0, 1], [1, 0], [1, 1]])
    enn = (T(rng.standard_normal((P, h * h)).astype(np.float32)), torch.zeros(h * h))
    gru = [T(rng.standard_normal((h, h)).astype(np.float32) * 0.3) for _ in range(6)] + [torch.zeros(h)] * 3
    x = T(rng.standard_normal((2, 5)).astype(np.float32))
    out = M.message_passing(x, T(rng.standard_normal((4, P)).astype(np.float32)), a2p, 2, h, enn, gru)
    assert out.shape == (2, h)
    try:
        M.message_passing(torch.zeros(2, h + 1), torch.zeros(4, P), a2p, 1, h, enn, gru)
    except ValueError as e:
        assert "Too large" in str(e)
    else:
        raise AssertionError("expected ValueError")


def test_weave_collation_equals_the_reference_loop():
    """MPNNModel.default_generator's collation (graph_models.py:1214-1246) restated with explicit loops: pairs of a
    molecule in row-major (destination, source) order, destinations ascending, pair features reshaped row-major."""
    from deepchem_b200.mpnn import weave_batch_inputs

    class Mol(object):
        def __init__(self, n, seed):
            r = np.random.RandomState(seed)
            self.n, self.af, self.pf = n, r.randn(n, 4).astype(np.float32), r.randn(n, n, 3).astype(np.float32)

        def get_num_atoms(self):
            return self.n

        def get_atom_features(self):
            return self.af

        def get_pair_features(self):
            return self.pf
    mols = [Mol(3, 0), Mol(1, 1), Mol(4, 2)]
    af, pf, split, a2p = weave_batch_inputs(mols, 3)
    exp_a2p, exp_pf, exp_split, start = [], [], [], 0
    for im, m in enumerate(mols):
        for i in range(m.n):
            exp_split.append(im)
            for j in range(m.n):
                exp_a2p.append([start + i, start + j])
                exp_pf.append(m.pf[i, j])
        start += m.n
    assert np.array_equal(a2p, np.array(exp_a2p)) and np.array_equal(split, np.array(exp_split))
    assert np.array_equal(pf, np.array(exp_pf)) and np.array_equal(af, np.concatenate([m.af for m in mols]))
    assert bool(np.all(np.diff(a2p[:, 0]) >= 0))          # what EdgeNetwork's segment sum needs


def test_set_gather_differentiable_restatement_and_gradcheck():
    """set_gather_d (dtype-generic, differentiable) equals the literal set_gather, and its float64 autograd — the
    gradient oracle of the CUDA backward kernels — passes gradcheck."""
    torch.manual_seed(0)
    h, B = 6, 4
    split = np.array([0, 0, 0, 1, 1, 3, 3, 3, 3])           # molecule 2 is empty
    x = torch.randn(9, h)
    U, b = torch.randn(2 * h, 4 * h) * 0.2, torch.randn(4 * h) * 0.1
    a = M.set_gather(x.numpy(), split, U, b, 3, B, h)
    d = M.set_gather_d(x, split, U, b, 3, B, h)
    assert float((a - d).abs().max()) < 1e-6
    x64, U64, b64 = x.double().requires_grad_(), U.double().requires_grad_(), b.double().requires_grad_()
    assert torch.autograd.gradcheck(lambda xx, uu, bb: M.set_gather_d(xx, split, uu, bb, 2, B, h), (x64, U64, b64),
                                    eps=1e-6, atol=1e-7)

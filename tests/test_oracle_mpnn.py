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

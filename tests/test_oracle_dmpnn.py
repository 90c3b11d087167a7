"""Pin the D-MPNN oracle (oracle/dmpnn_torch.py) to the reference: integer tables of _MapperDMPNN bit
for bit, DMPNNEncoderLayer / PositionwiseFeedForward outputs, and the reference's own known answer
(deepchem/models/tests/test_layers.py:798-827, test_mapper_dmpnn.py:17-111).  Fixture: tests/golden/ref_dmpnn.npz
(generated from the reference by tests/golden/make_golden_dmpnn.py)."""
import os

import numpy as np
import torch

from oracle import dmpnn_torch as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dmpnn.npz"), allow_pickle=False)
N_GRAPHS = len(G["names"])


def graph(i):
    return O.OracleGraph(G["g%d_node_features" % i], G["g%d_edge_index" % i], G["g%d_edge_features" % i])


def test_mapper_tables_bit_exact():
    for i in range(N_GRAPHS):
        af, f_ini, a2b, mapping, _ = O.mapper_values(graph(i))
        assert np.array_equal(a2b, G["g%d_a2b" % i]), G["names"][i]
        assert np.array_equal(mapping, G["g%d_mapping" % i]), G["names"][i]
        assert np.array_equal(f_ini, G["g%d_f_ini" % i]), G["names"][i]


def test_mapper_reference_known_answers():
    """models/tests/test_mapper_dmpnn.py: benzene ring tables (:17-24, :86-111) and the bond-free salt."""
    names = list(G["names"])
    b = names.index("benzene")
    _, _, a2b, mapping, _ = O.mapper_values(graph(b))
    # directed bonds (0->1,1->0,1->2,2->1,...,5->0,0->5): atom 0 receives bonds 1 and 10
    assert a2b.tolist() == [[1, 10], [0, 3], [2, 5], [4, 7], [6, 9], [8, 11]]
    assert mapping.tolist() == [[-1, 10], [-1, 3], [0, -1], [-1, 5], [2, -1], [-1, 7], [4, -1], [-1, 9], [6, -1],
                                [-1, 11], [8, -1], [1, -1], [-1, -1]]
    s = names.index("salt")
    _, f_ini, a2b, mapping, _ = O.mapper_values(graph(s))
    assert f_ini.shape == (1, 147) and not f_ini.any()
    assert a2b.tolist() == [[-1], [-1]] and mapping.tolist() == [[-1]]


def _encoder(prefix, d_hidden, depth, bias, agg, norm=7):
    enc = O.OracleDMPNNEncoder(133, 14, d_hidden, depth, bias, 'relu', agg, norm)
    sd = {k: torch.from_numpy(G[prefix + k]) for k in enc.state_dict().keys()}
    enc.load_state_dict(sd)
    return enc


def test_encoder_matches_reference_single_and_batched():
    values = [O.mapper_values(graph(i)) for i in range(N_GRAPHS)]
    for ci in range(3):
        agg, bias, depth = G["enc%d_cfg" % ci]
        enc = _encoder("enc%d_" % ci, 64, int(depth), bias == "True", str(agg))
        with torch.no_grad():
            singles = np.concatenate([enc(*O.to_torch_batch(O.collate([v]))).numpy() for v in values], 0)
            batch = enc(*O.to_torch_batch(O.collate(values))).numpy()
        assert np.abs(singles - G["enc%d_single" % ci]).max() < 1e-6
        assert np.abs(batch - G["enc%d_batch" % ci]).max() < 1e-6
        if bias != "True":
            # bias-free encoders are collation-invariant (the -1 pads hit all-zero rows): the batched result
            # equals the per-molecule results the reference's tests pin
            assert np.abs(batch - singles).max() < 1e-5


def test_encoder_with_global_features():
    values = [O.mapper_values(graph(i)) for i in range(N_GRAPHS)]
    gvals = [(v[0], v[1], v[2], v[3], G["encg_global"][i]) for i, v in enumerate(values)]
    enc = _encoder("encg_", 32, 3, False, "mean", 100)
    with torch.no_grad():
        out = enc(*O.to_torch_batch(O.collate(gvals))).numpy()
    assert out.shape == (N_GRAPHS, 34)
    assert np.abs(out - G["encg_batch"]).max() < 1e-6


def test_ffn_matches_reference():
    ffn = O.OracleFFN(64, 48, 5, 'relu', 3)
    ffn.load_state_dict({k: torch.from_numpy(G["ffn_" + k]) for k in ffn.state_dict().keys()})
    with torch.no_grad():
        y = ffn(torch.from_numpy(G["ffn_x"])).numpy()
    assert np.abs(y - G["ffn_y"]).max() < 1e-6


def test_encoder_known_answer_cc():
    """deepchem/models/tests/test_layers.py:798-827: 'CC', seed-0 initialisation -> [0.1116, 0.0470]."""
    g = O.OracleGraph(G["kat_atom_features"], np.asarray([[0, 1], [1, 0]]), G["kat_bond_features"])
    enc = O.OracleDMPNNEncoder(133, 14, 2, 3, False, 'relu', 'mean', 100)
    enc.load_state_dict({k: torch.from_numpy(G["kat_" + k]) for k in enc.state_dict().keys()})
    with torch.no_grad():
        out = enc(*O.to_torch_batch(O.collate([O.mapper_values(g)]))).numpy()
    assert np.allclose(out, [[0.1116, 0.0470]], atol=1e-4)
    assert np.abs(out - G["kat_out"]).max() < 1e-7


def test_encoder_gradients_float64_gradcheck():
    values = [O.mapper_values(graph(i)) for i in (1, 3, 6)]
    enc = O.OracleDMPNNEncoder(133, 14, 8, 3, True, 'tanh', 'mean').double()
    af, f_ini, a2b, mapping, gf, key = O.to_torch_batch(O.collate(values), torch.float64)
    f_ini.requires_grad_(True)
    af.requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, f: enc(a, f, a2b, mapping, gf, key), (af, f_ini), atol=1e-6)

"""C++ D-MPNN table builder (dcgc_dmpnn_plan/build) against the reference's _MapperDMPNN tables
(tests/golden/ref_dmpnn.npz, generated from the reference) and the oracle's collation: integer work,
bit-exact.  Host only."""
import os

import numpy as np
import pytest

from deepchem_b200.dmpnn import DmpnnLayout, _MapperDMPNN
from deepchem_b200.dmpnn_data import GraphData, PackedGraphs, make_graphs
from oracle import dmpnn_torch as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dmpnn.npz"), allow_pickle=False)
N_GRAPHS = len(G["names"])


def gdata(i):
    return GraphData(G["g%d_node_features" % i], G["g%d_edge_index" % i], G["g%d_edge_features" % i])


def test_single_molecule_tables_equal_reference():
    for i in range(N_GRAPHS):
        m = _MapperDMPNN(gdata(i))
        af, f_ini, a2b, mapping, gf = m.values
        assert np.array_equal(a2b, G["g%d_a2b" % i]), G["names"][i]
        assert np.array_equal(mapping, G["g%d_mapping" % i]), G["names"][i]
        assert np.array_equal(f_ini, G["g%d_f_ini" % i]), G["names"][i]


@pytest.mark.parametrize("keep_pads", [False, True])
def test_batched_tables_equal_collated_reference(keep_pads):
    graphs = [gdata(i) for i in range(N_GRAPHS)]
    packed = PackedGraphs.from_graphs(graphs)
    lay = DmpnnLayout.build(packed, keep_pads=keep_pads)
    values = [(G["g%d_node_features" % i], G["g%d_f_ini" % i], G["g%d_a2b" % i], G["g%d_mapping" % i], np.empty(0))
              for i in range(N_GRAPHS)]                       # the reference's own per-molecule tables
    af, f_ini, a2b, mapping, gf, key = O.collate(values)
    assert lay.k == a2b.shape[1]
    assert np.array_equal(lay.a2b_ell, a2b) and np.array_equal(lay.map_ell, mapping)
    assert np.diff(lay.mol_ptr).tolist() == key
    R = f_ini.shape[0]
    assert lay.n_rows == R
    # f_ini assembled from bond_src / bond_edge equals the reference's hstack + zero pad rows
    nf = np.concatenate([g.node_features for g in graphs])
    ef = np.concatenate([g.edge_features for g in graphs])
    mine = np.zeros_like(f_ini)
    live = lay.bond_src >= 0
    mine[live] = np.hstack([nf[lay.bond_src[live]], ef[lay.bond_edge[live]]])
    assert np.array_equal(mine, f_ini)
    # CSR forms: gathers through them reproduce the ELL gathers (pads dropped hit zero rows only)
    x = np.random.default_rng(0).standard_normal((R, 3))
    if not keep_pads:
        x[~live] = 0.0                                        # pad rows are zero when bias == False
    ell = lambda tab: x[np.where(tab < 0, tab + R, tab)].sum(1)   # noqa: E731  torch negative indexing
    def csr(ptr, idx):
        out = np.zeros((ptr.shape[0] - 1, 3))
        for r in range(out.shape[0]):
            out[r] = x[idx[ptr[r]:ptr[r + 1]]].sum(0)
        return out
    assert np.allclose(csr(lay.map_ptr, lay.map_idx), ell(mapping), atol=1e-12)
    assert np.allclose(csr(lay.a2b_ptr, lay.a2b_idx), ell(a2b), atol=1e-12)
    # transposes
    for ptr, idx, tptr, tidx, n_in in ((lay.map_ptr, lay.map_idx, lay.map_t_ptr, lay.map_t_idx, R),
                                       (lay.a2b_ptr, lay.a2b_idx, lay.a2b_t_ptr, lay.a2b_t_idx, R)):
        pairs = sorted((int(idx[e]), r) for r in range(ptr.shape[0] - 1) for e in range(ptr[r], ptr[r + 1]))
        tp = [(c, int(tidx[e])) for c in range(n_in) for e in range(tptr[c], tptr[c + 1])]
        assert pairs == tp


def test_larger_synthetic_batch_against_oracle_mapper():
    pg = make_graphs(300, seed=9, shape="qm9", no_bond_fraction=0.05)
    lay = DmpnnLayout.build(pg)
    values = [O.mapper_values(O.OracleGraph(*pg.graph(i)[:3])) for i in range(pg.n_mols)]
    af, f_ini, a2b, mapping, gf, key = O.collate(values)
    assert np.array_equal(lay.a2b_ell, a2b) and np.array_equal(lay.map_ell, mapping)
    assert lay.n_rows == f_ini.shape[0] and lay.n_atoms == af.shape[0]


def test_builder_rejects_bad_bond_index():
    pg = make_graphs(4, seed=1)
    pg.edge_dst = pg.edge_dst.copy()
    pg.edge_dst[0] = 99
    with pytest.raises(ValueError):
        DmpnnLayout.build(pg)


def test_encoder_and_ffn_state_dicts_have_the_reference_layout():
    """Checkpoint compatibility without a device: the encoder / feed-forward modules carry exactly the parameter
    names and shapes of the reference modules whose weights the fixture holds (with and without encoder bias)."""
    import torch
    from deepchem_b200.dmpnn import DMPNNEncoderLayer, PositionwiseFeedForward
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dmpnn.npz"), allow_pickle=False)
    for ci in (0, 1, 2):
        agg, bias, depth = G["enc%d_cfg" % ci]
        enc = DMPNNEncoderLayer(use_default_fdim=False, atom_fdim=133, bond_fdim=14, d_hidden=64, depth=int(depth),
                                bias=bias == "True", aggregation=str(agg), aggregation_norm=7)
        want = {k[len("enc%d_" % ci):]: G[k] for k in G.files
                if k.startswith("enc%d_W_" % ci)}
        own = enc.state_dict()
        assert sorted(own) == sorted(want)
        for k, v in own.items():
            assert tuple(v.shape) == want[k].shape, (ci, k)
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in want.items()}, strict=True)
    want = {k[len("ffn_"):]: G[k] for k in G.files if k.startswith("ffn_linears")}
    d_in, d_hid, d_out = want["linears.0.weight"].shape[1], want["linears.0.weight"].shape[0], want["linears.2.weight"].shape[0]
    ffn = PositionwiseFeedForward(d_input=d_in, d_hidden=d_hid, d_output=d_out, activation='relu', n_layers=3)
    own = ffn.state_dict()
    assert sorted(own) == sorted(want) and all(tuple(own[k].shape) == want[k].shape for k in own)

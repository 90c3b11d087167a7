"""Generate the D-MPNN golden fixture by running the REFERENCE in the build container.

    python tests/golden/make_golden_dmpnn.py      ->  tests/golden/ref_dmpnn.npz

Contents:
  * ``_MapperDMPNN`` (deepchem/models/torch_models/dmpnn.py:38-243) outputs for a seeded set of QM9-shaped
    graphs plus the reference's own test molecules (models/tests/test_mapper_dmpnn.py: C, CC, CCC,
    benzene ring, the two-atom no-bond salt) — integer tables, to be matched bit for bit;
  * ``DMPNNEncoderLayer`` (torch_models/layers.py:1436-1649) forward on every single molecule (batch of one:
    the only thing the reference's tests pin) AND on the whole batch collated by the ``__inc__`` rule
    (dmpnn.py:17-35; torch_geometric itself is not installed), with seeded weights, for the three
    aggregations, with and without bias, depth 3 and 4, with global features;
  * ``PositionwiseFeedForward`` (layers.py:795-910) forward with the DMPNN defaults;
  * the reference's known answer of models/tests/test_layers.py:798-827 replayed through the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _refimport import import_reference  # noqa: E402

dc = import_reference()
import torch  # noqa: E402
from deepchem.feat.graph_data import GraphData  # noqa: E402
from deepchem.models.torch_models.dmpnn import _MapperDMPNN  # noqa: E402
from deepchem.models.torch_models import layers as L  # noqa: E402
from deepchem_b200.dmpnn_data import make_graphs  # noqa: E402
from oracle import dmpnn_torch as O  # noqa: E402   (only its collate(): the PyG collation restatement)


def ref_values(node_features, edge_index, edge_features, global_features):
    g = GraphData(node_features=node_features, edge_index=edge_index, edge_features=edge_features,
                  global_features=global_features)
    return _MapperDMPNN(g).values


def named_graphs(rng, atom_fdim=133, bond_fdim=14):
    """Topologies of models/tests/test_mapper_dmpnn.py with random features."""
    tops = {
        "C": (1, []),
        "CC": (2, [(0, 1)]),
        "CCC": (3, [(0, 1), (1, 2)]),
        "benzene": (6, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0)]),
        "salt": (2, []),
    }
    out = []
    for name, (n, bonds) in tops.items():
        src = [a for b in bonds for a in b]
        dst = [a for b in bonds for a in b[::-1]]
        ei = np.asarray([src, dst], dtype=np.int64).reshape(2, -1)
        nf = rng.random((n, atom_fdim))
        ef = rng.random((ei.shape[1], bond_fdim))
        out.append((name, nf, ei, ef, np.empty(0)))
    return out


def main():
    rng = np.random.default_rng(11)
    store = {}
    graphs = [(n, nf, ei, ef, gf) for n, nf, ei, ef, gf in named_graphs(rng)]
    pg = make_graphs(24, seed=5, shape="qm9")
    for i in range(pg.n_mols):
        nf, ei, ef, gf = pg.graph(i)
        graphs.append(("qm9_%d" % i, nf.astype(np.float64), ei, ef.astype(np.float64), np.empty(0)))
    store["names"] = np.asarray([g[0] for g in graphs])
    values = []
    for i, (name, nf, ei, ef, gf) in enumerate(graphs):
        v = ref_values(nf, ei, ef, gf)
        values.append(v)
        store["g%d_node_features" % i] = nf
        store["g%d_edge_index" % i] = ei
        store["g%d_edge_features" % i] = ef
        store["g%d_f_ini" % i] = v[1]
        store["g%d_a2b" % i] = v[2]
        store["g%d_mapping" % i] = v[3]

    # ---- encoder: reference layer, seeded weights
    cfgs = [("mean", False, 3), ("sum", True, 3), ("norm", False, 4)]
    coll = O.collate(values)
    for ci, (agg, bias, depth) in enumerate(cfgs):
        torch.manual_seed(100 + ci)
        enc = L.DMPNNEncoderLayer(use_default_fdim=False, atom_fdim=133, bond_fdim=14, d_hidden=64, depth=depth,
                                  bias=bias, activation='relu', dropout_p=0.0, aggregation=agg, aggregation_norm=7)
        for k, t in enc.state_dict().items():
            store["enc%d_%s" % (ci, k)] = t.numpy().copy()
        store["enc%d_cfg" % ci] = np.asarray([agg, str(bias), str(depth)])
        singles = []
        for v in values:
            b = O.to_torch_batch(O.collate([v]))
            with torch.no_grad():
                singles.append(enc(*b).numpy())
        store["enc%d_single" % ci] = np.concatenate(singles, 0)
        with torch.no_grad():
            store["enc%d_batch" % ci] = enc(*O.to_torch_batch(coll)).numpy()
    # with global features (2 per molecule), mean aggregation
    torch.manual_seed(7)
    gvals = [(v[0], v[1], v[2], v[3], rng.random(2)) for v in values]
    enc = L.DMPNNEncoderLayer(use_default_fdim=False, atom_fdim=133, bond_fdim=14, d_hidden=32, depth=3)
    for k, t in enc.state_dict().items():
        store["encg_%s" % k] = t.numpy().copy()
    store["encg_global"] = np.stack([g[4] for g in gvals])
    with torch.no_grad():
        store["encg_batch"] = enc(*O.to_torch_batch(O.collate(gvals))).numpy()

    # ---- FFN with the DMPNN defaults (dmpnn.py:392-399)
    torch.manual_seed(3)
    ffn = L.PositionwiseFeedForward(d_input=64, d_hidden=48, d_output=5, activation='relu', n_layers=3,
                                    dropout_p=0.0, dropout_at_input_no_act=True)
    x = torch.randn(9, 64)
    for k, t in ffn.state_dict().items():
        store["ffn_%s" % k] = t.numpy().copy()
    store["ffn_x"] = x.numpy()
    with torch.no_grad():
        store["ffn_y"] = ffn(x).numpy()

    # ---- the reference's own known answer (models/tests/test_layers.py:798-827): 'CC', seed 0
    torch.manual_seed(0)
    af = np.zeros((2, 133))
    for j in (5, 105, 112, 114, 122, 127):
        af[:, j] = 1
    af[:, 132] = 0.12011
    bf = np.zeros((2, 14))
    bf[:, 1] = 1
    bf[:, 7] = 1
    v = ref_values(af, np.asarray([[0, 1], [1, 0]]), bf, np.empty(0))
    layer = L.DMPNNEncoderLayer(use_default_fdim=False, atom_fdim=133, bond_fdim=14, d_hidden=2, depth=3,
                                bias=False, activation='relu', dropout_p=0.0, aggregation='mean', aggregation_norm=100)
    with torch.no_grad():
        out = layer(*O.to_torch_batch(O.collate([v]))).numpy()
    assert np.allclose(out, [[0.1116, 0.0470]], atol=1e-4), out
    store["kat_out"] = out
    store["kat_atom_features"] = af
    store["kat_bond_features"] = bf
    for k, t in layer.state_dict().items():
        store["kat_%s" % k] = t.numpy().copy()

    np.savez_compressed(os.path.join(HERE, "ref_dmpnn.npz"), **store)
    print("wrote ref_dmpnn.npz with %d arrays, %d graphs" % (len(store), len(graphs)))


if __name__ == "__main__":
    main()

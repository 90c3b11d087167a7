"""Generator fixture from the reference checkout.  Run in the build container:

    python tests/golden/make_golden_generator.py

  ref_generator.npz   what the REFERENCE's ``GraphConvModel.default_generator`` (torch_models/graphconvmodel.py:382-422)
        yields for a ``NumpyDataset`` of 23 reference ``ConvMol`` objects with batch_size 10, classification, 2 tasks,
        ``pad_batches=True``: per batch the input list [features, deg_slice, membership, n_samples, deg_adj_1..10],
        the one-hot labels and the weights — the third batch holds 3 molecules and is padded to 10 by
        ``pad_batch`` (data/datasets.py:142-218: the molecules repeated, zero weights on the copies).  Also a 'predict'
        pass with ``pad_batches=False`` (labels stay class indices, the last batch keeps 3 molecules), and a 'tiny' pass:
        the first 3 molecules alone with batch_size 10 (the batch is filled by repeating the dataset 3 1/3 times).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, HERE)
    import make_golden as G            # imports the reference (rdkit stubbed) and the helpers
    import torch
    from deepchem.data import NumpyDataset
    from deepchem.feat.mol_graphs import ConvMol
    from deepchem.models.torch_models.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_molecules
    pm = make_molecules(23, seed=51, shape="stress")
    mols = pm.to_list()
    d = G.pack_mols(mols)
    d["features"] = d["features"].astype(np.float32)
    X = np.empty(len(mols), dtype=object)
    for i, (f, adj) in enumerate(mols):
        X[i] = ConvMol(np.asarray(f, dtype=np.float64), adj)
    rng = np.random.default_rng(5)
    y = rng.integers(0, 2, size=(len(mols), 2)).astype(np.float64)
    w = rng.random((len(mols), 2))
    ds = NumpyDataset(X, y, w)
    ref = GraphConvModel(2, number_input_features=[75, 64], graph_conv_layers=[64, 64], mode="classification",
                         batch_size=10, device=torch.device("cpu"))
    d["y"], d["w"] = y.astype(np.float32), w.astype(np.float32)
    for tag, kw in (("fit", dict(mode="fit", pad_batches=True)), ("predict", dict(mode="predict", pad_batches=False))):
        n = 0
        for inputs, labels, weights in ref.default_generator(ds, epochs=1, deterministic=True, **kw):
            for k, a in enumerate(inputs):
                a = np.asarray(a)
                d["%s_b%d_in%d" % (tag, n, k)] = a.astype(np.float32) if a.dtype == np.float64 else a
            d["%s_b%d_y" % (tag, n)] = np.asarray(labels[0], dtype=np.float32)
            d["%s_b%d_w" % (tag, n)] = np.asarray(weights[0], dtype=np.float32)
            n += 1
        d["%s_batches" % tag] = np.array(n)
        print(tag, "batches", n, "last n_samples", int(inputs[3]), "labels", np.asarray(labels[0]).shape)
    tiny = NumpyDataset(X[:3], y[:3], w[:3])
    got = list(ref.default_generator(tiny, epochs=1, deterministic=True, mode="fit", pad_batches=True))
    assert len(got) == 1
    inputs, labels, weights = got[0]
    for k, a in enumerate(inputs):
        a = np.asarray(a)
        d["tiny_b0_in%d" % k] = a.astype(np.float32) if a.dtype == np.float64 else a
    d["tiny_b0_y"] = np.asarray(labels[0], dtype=np.float32)
    d["tiny_b0_w"] = np.asarray(weights[0], dtype=np.float32)
    d["tiny_batches"] = np.array(1)
    print("tiny: n_samples", int(inputs[3]), "atoms", np.asarray(inputs[0]).shape[0], "weights", np.asarray(weights[0])[:, 0])
    np.savez_compressed(os.path.join(HERE, "ref_generator.npz"), **d)


if __name__ == "__main__":
    main()

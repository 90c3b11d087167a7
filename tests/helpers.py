"""Shared helpers for the test-suite (not product code)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def unpack_mols(d):
    """npz dict -> list of (features, adj_list)."""
    atom_ptr, adj_ptr, adj_idx, feats = d["atom_ptr"], d["adj_ptr"], d["adj_idx"], d["features"]
    mols = []
    for m in range(len(atom_ptr) - 1):
        a0, a1 = int(atom_ptr[m]), int(atom_ptr[m + 1])
        adj = [adj_idx[adj_ptr[a]:adj_ptr[a + 1]].tolist() for a in range(a0, a1)]
        mols.append((feats[a0:a1], adj))
    return mols


def oracle_batch(mols):
    from oracle.convmol_layout import OracleConvMol, agglomerate
    cms = [OracleConvMol(np.asarray(f), adj) for f, adj in mols]
    return cms, agglomerate(cms)


def torch_args(mm, n_samples=None, dtype=torch.float32):
    x = torch.from_numpy(np.asarray(mm.get_atom_features())).to(dtype)
    args = [x, torch.from_numpy(mm.deg_slice), torch.from_numpy(mm.membership)]
    if n_samples is not None:
        args.append(torch.tensor(n_samples))
    return args + [torch.from_numpy(a) for a in mm.get_deg_adjacency_lists()[1:]]


def rel_err(a, b):
    """max|a-b| / max(|b|, tiny): error relative to the reference tensor's scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "non-finite pattern differs"
    assert np.array_equal(a[~fin], b[~fin]), "non-finite values differ"
    if not fin.any():
        return 0.0
    return float(np.abs(a[fin] - b[fin]).max() / max(np.abs(b[fin]).max(), 1e-30))

"""Packed molecular graphs for the D-MPNN path (the output contract of ``DMPNNFeaturizer``:
``GraphData(node_features [n,133], edge_index [2,E], edge_features [E,14], global_features)``,
deepchem/feat/molecule_featurizers/dmpnn_featurizer.py, deepchem/feat/graph_data.py) and a seeded
QM9-shaped synthetic generator (4-9 heavy atoms, tree + at most one ring, in-degree <= 4).

Directed bonds come in (i->j, j->i) pairs at positions (2k, 2k+1): the reference relies on that to find
the reverse bond (``_replace_rev_bonds``, deepchem/models/torch_models/dmpnn.py:234-243).
"""
import numpy as np


class GraphData(object):
    """Minimal stand-in for deepchem.feat.GraphData (no RDKit / PyG on the box)."""

    def __init__(self, node_features, edge_index, edge_features=None, global_features=None, **kwargs):
        self.node_features = np.asarray(node_features)
        self.edge_index = np.asarray(edge_index).reshape(2, -1)
        self.edge_features = None if edge_features is None else np.asarray(edge_features)
        self.global_features = np.empty(0) if global_features is None else np.asarray(global_features)
        self.num_nodes = self.node_features.shape[0]
        self.num_node_features = self.node_features.shape[1]
        self.num_edges = self.edge_index.shape[1]
        self.num_edge_features = 0 if self.edge_features is None else self.edge_features.shape[1]
        for k, v in kwargs.items():
            setattr(self, k, v)


class PackedGraphs(object):
    """A shard of molecular graphs with no Python objects inside.

    node_ptr [B+1] int32, edge_ptr [B+1] int32, edge_src / edge_dst [E] int32 (molecule-local atom ids),
    node_features [A,Fa] f32, edge_features [E,Fb] f32, global_features [B,G] f32 (G may be 0)."""

    __slots__ = ("node_ptr", "edge_ptr", "edge_src", "edge_dst", "node_features", "edge_features",
                 "global_features", "_pin_nf", "_pin_ef")

    def __init__(self, node_ptr, edge_ptr, edge_src, edge_dst, node_features, edge_features, global_features=None):
        self.node_ptr = np.ascontiguousarray(node_ptr, dtype=np.int32)
        self.edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int32)
        self.edge_src = np.ascontiguousarray(edge_src, dtype=np.int32)
        self.edge_dst = np.ascontiguousarray(edge_dst, dtype=np.int32)
        self.node_features = np.ascontiguousarray(node_features, dtype=np.float32)
        self.edge_features = np.ascontiguousarray(edge_features, dtype=np.float32)
        n = self.node_ptr.shape[0] - 1
        if global_features is None:
            global_features = np.zeros((n, 0), np.float32)
        self.global_features = np.ascontiguousarray(global_features, dtype=np.float32).reshape(n, -1)
        self._pin_nf = self._pin_ef = None

    def pin_memory(self):
        """Move the two feature matrices into page-locked host memory (needs torch + CUDA): batches sliced from the
        shard then upload with asynchronous DMA and no staging copy (the role PackedMols.pin_memory plays for
        GraphConv batches)."""
        import torch
        if getattr(self, "_pin_nf", None) is None:
            for name, pin in (("node_features", "_pin_nf"), ("edge_features", "_pin_ef")):
                a = getattr(self, name)
                t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
                t.numpy()[...] = a
                setattr(self, name, t.numpy())
                setattr(self, pin, t)
        return self

    n_mols = property(lambda self: self.node_ptr.shape[0] - 1)
    n_atoms = property(lambda self: int(self.node_ptr[-1]))
    n_bonds = property(lambda self: int(self.edge_ptr[-1]))

    def __len__(self):
        return self.n_mols

    def graph(self, i):
        """(node_features, edge_index [2,E], edge_features, global_features) of molecule i."""
        a0, a1 = int(self.node_ptr[i]), int(self.node_ptr[i + 1])
        e0, e1 = int(self.edge_ptr[i]), int(self.edge_ptr[i + 1])
        ei = np.stack([self.edge_src[e0:e1], self.edge_dst[e0:e1]]).astype(np.int64)
        return self.node_features[a0:a1], ei, self.edge_features[e0:e1], self.global_features[i]

    def to_list(self):
        return [GraphData(*self.graph(i)) for i in range(self.n_mols)]

    def slice(self, lo, hi):
        if lo == 0 and hi == self.n_mols:
            return self
        a0, a1 = int(self.node_ptr[lo]), int(self.node_ptr[hi])
        e0, e1 = int(self.edge_ptr[lo]), int(self.edge_ptr[hi])
        out = PackedGraphs.__new__(PackedGraphs)
        out.node_ptr = self.node_ptr[lo:hi + 1] - np.int32(a0)
        out.edge_ptr = self.edge_ptr[lo:hi + 1] - np.int32(e0)
        out.edge_src, out.edge_dst = self.edge_src[e0:e1], self.edge_dst[e0:e1]
        out.node_features, out.edge_features = self.node_features[a0:a1], self.edge_features[e0:e1]
        out.global_features = self.global_features[lo:hi]
        pn, pe = getattr(self, "_pin_nf", None), getattr(self, "_pin_ef", None)
        out._pin_nf = pn[a0:a1] if pn is not None else None       # torch views of the same pinned rows
        out._pin_ef = pe[e0:e1] if pe is not None else None
        return out

    def take(self, idx):
        """Molecules idx[0], idx[1], ... (repeats allowed) as a new shard; vectorised (a shuffled epoch takes this
        path for every batch)."""
        from .synthetic import _ranges
        idx = np.asarray(idx, dtype=np.int64)
        n_at = (self.node_ptr[idx + 1] - self.node_ptr[idx]).astype(np.int64)
        n_ed = (self.edge_ptr[idx + 1] - self.edge_ptr[idx]).astype(np.int64)
        rows = _ranges(np.asarray(self.node_ptr)[idx], n_at)
        eds = _ranges(np.asarray(self.edge_ptr)[idx], n_ed)
        return PackedGraphs(np.concatenate([[0], np.cumsum(n_at)]), np.concatenate([[0], np.cumsum(n_ed)]),
                            np.asarray(self.edge_src)[eds], np.asarray(self.edge_dst)[eds], self.node_features[rows],
                            self.edge_features[eds], self.global_features[idx])

    @staticmethod
    def from_graphs(graphs, atom_fdim=None, bond_fdim=None):
        """graphs: GraphData-like objects (ours or the reference's)."""
        node_ptr, edge_ptr, src, dst, nf, ef, gf = [0], [0], [], [], [], [], []
        for g in graphs:
            node_ptr.append(node_ptr[-1] + int(g.num_nodes))
            ei = np.asarray(g.edge_index).reshape(2, -1)
            edge_ptr.append(edge_ptr[-1] + ei.shape[1])
            src.append(ei[0])
            dst.append(ei[1])
            nf.append(np.asarray(g.node_features, dtype=np.float32).reshape(int(g.num_nodes), -1))
            e = g.edge_features
            if e is not None and np.asarray(e).size:
                ef.append(np.asarray(e, dtype=np.float32).reshape(ei.shape[1], -1))
            gl = getattr(g, "global_features", None)
            gf.append(np.zeros(0, np.float32) if gl is None else np.asarray(gl, dtype=np.float32).reshape(-1))
        fa = atom_fdim if atom_fdim is not None else (nf[0].shape[1] if nf else 0)
        fb = bond_fdim if bond_fdim is not None else (ef[0].shape[1] if ef else 0)
        cat = lambda xs, w, dt: np.concatenate(xs, 0) if xs else np.zeros((0, w) if w is not None else 0, dt)  # noqa: E731
        glob = np.stack(gf) if gf and gf[0].size else np.zeros((len(graphs), 0), np.float32)
        return PackedGraphs(node_ptr, edge_ptr, cat(src, None, np.int32), cat(dst, None, np.int32),
                            cat(nf, fa, np.float32), cat(ef, fb, np.float32), glob)


def make_graphs(n_mols, seed=0, shape="qm9", atom_fdim=133, bond_fdim=14, global_size=0, no_bond_fraction=0.0):
    """Seeded synthetic PackedGraphs.  ``shape='qm9'``: 4-9 atoms, spanning tree with valence cap 4 plus at
    most one ring closure; 0/1 features (~7 ones per atom row, ~2 per bond row) and a mass-like real column
    at index atom_fdim-1, as DMPNNFeaturizer produces.  ``no_bond_fraction``: share of single-atom,
    bond-free molecules (exercises the zero-bond branch, dmpnn.py:154-161)."""
    from .synthetic import _SHAPES, _random_mol_adj
    rng = np.random.default_rng(seed)
    mean, lo, hi, cap, rings = _SHAPES[shape]
    node_ptr, edge_ptr, src, dst, nf, ef = [0], [0], [], [], [], []
    for _ in range(n_mols):
        if shape == "qm9":
            n = int(rng.integers(lo, hi + 1))
        else:
            n = int(np.clip(rng.poisson(mean), lo, hi))
        if no_bond_fraction > 0 and rng.random() < no_bond_fraction:
            n = 1
        adj = _random_mol_adj(rng, n, cap, rings) if n > 1 else [[]]
        bonds = [(i, j) for i in range(n) for j in adj[i] if i < j]
        for i, j in bonds:                      # (i->j, j->i) at positions (2k, 2k+1)
            src += [i, j]
            dst += [j, i]
        f = np.zeros((n, atom_fdim), np.float32)
        f[np.arange(n)[:, None], rng.integers(0, atom_fdim - 1, size=(n, 7))] = 1.0
        f[:, atom_fdim - 1] = rng.uniform(0.1, 0.35, size=n)
        b = np.zeros((2 * len(bonds), bond_fdim), np.float32)
        if len(bonds):
            one = np.zeros((len(bonds), bond_fdim), np.float32)
            one[np.arange(len(bonds))[:, None], rng.integers(0, bond_fdim, size=(len(bonds), 2))] = 1.0
            b = np.repeat(one, 2, axis=0)       # both directions carry the same bond features
        nf.append(f)
        ef.append(b)
        node_ptr.append(node_ptr[-1] + n)
        edge_ptr.append(edge_ptr[-1] + 2 * len(bonds))
    glob = rng.random((n_mols, global_size)).astype(np.float32) if global_size else None
    return PackedGraphs(node_ptr, edge_ptr, np.asarray(src, np.int32), np.asarray(dst, np.int32),
                        np.concatenate(nf, 0), np.concatenate(ef, 0) if ef else np.zeros((0, bond_fdim), np.float32),
                        glob)

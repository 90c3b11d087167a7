"""Oracle (numpy) for the degree-bucketed ConvMol batch layout.  TEST INFRASTRUCTURE ONLY.

Restates, without sharing code, what the reference computes in
  * ``ConvMol.__init__`` / ``ConvMol._deg_sort``   (deepchem/feat/mol_graphs.py:48-185)
  * ``ConvMol.agglomerate_mols``                   (deepchem/feat/mol_graphs.py:256-349)
  * ``MultiConvMol``                               (deepchem/feat/mol_graphs.py:352-375)

All integer outputs must match the reference bit for bit (dtype included); this is
checked against reference-generated fixtures in tests/test_oracle_golden.py.
"""
import numpy as np

MAX_DEG = 10
MIN_DEG = 0


class OracleConvMol(object):
    """One molecule after the per-molecule stable degree sort (mol_graphs.py:113-185).

    Attributes mirror the reference object so the reference's own
    ``agglomerate_mols`` could consume it: ``atom_features``, ``deg_list``,
    ``canon_adj_list``, ``deg_adj_lists``, ``deg_slice`` (int32, start zeroed for empty
    buckets, mol_graphs.py:184), ``degree_list``, ``deg_block_indices``, ``membership``.
    """

    def __init__(self, atom_features, adj_list, max_deg=MAX_DEG, min_deg=MIN_DEG):
        atom_features = np.asarray(atom_features)
        n = atom_features.shape[0]
        self.n_atoms, self.n_feat = atom_features.shape
        self.max_deg, self.min_deg = max_deg, min_deg
        deg = np.fromiter((len(a) for a in adj_list), dtype=np.int32, count=n)
        # stable sort by degree == lexsort((old_index, degree))  (mol_graphs.py:121)
        perm = np.argsort(deg, kind="stable")
        inv = np.empty(n, dtype=np.int64)
        inv[perm] = np.arange(n)
        self.atom_features = atom_features[perm, :]
        self.deg_list = [int(deg[i]) for i in perm]
        self.membership = n * [0]
        # neighbour lists follow their atom; entries are renumbered, order kept
        # (mol_graphs.py:138-141)
        self.canon_adj_list = [[int(inv[k]) for k in adj_list[i]] for i in perm]
        sdeg = deg[perm]
        nb = max_deg + 1 - min_deg
        self.deg_adj_lists = []
        counts = np.zeros(nb, dtype=np.int32)
        for d in range(min_deg, max_deg + 1):
            rows = np.nonzero(sdeg == d)[0]
            counts[d - min_deg] = rows.size
            if rows.size:
                block = np.array([self.canon_adj_list[i] for i in rows],
                                 dtype=np.int32).reshape(rows.size, d)
            else:
                block = np.zeros((0, d), dtype=np.int32)
            self.deg_adj_lists.append(block)
        ds = np.zeros((nb, 2), dtype=np.int32)
        ds[:, 1] = counts
        ds[1:, 0] = np.cumsum(counts)[:-1]
        ds[:, 0] *= (ds[:, 1] != 0)
        self.deg_slice = ds
        self.degree_list = [int(x) for x in sdeg]
        self.deg_id_list = np.array(self.deg_list) - min_deg
        starts = np.concatenate([[0], np.cumsum(counts)])
        self.deg_start = [int(s) for s in starts]
        self.deg_block_indices = (np.arange(n) - starts[sdeg - min_deg]).astype(np.int32) \
            if n else np.zeros(0, np.int32)

    def get_num_atoms(self):
        return self.n_atoms

    def get_atom_features(self):
        return self.atom_features

    def get_adjacency_list(self):
        return self.canon_adj_list

    def get_deg_adjacency_lists(self):
        return self.deg_adj_lists

    def get_deg_slice(self):
        return self.deg_slice


class OracleMultiConvMol(object):
    """Batch of molecules in degree-major order (mol_graphs.py:352-375)."""

    def __init__(self, nodes, deg_adj_lists, deg_slice, membership, num_mols):
        self.nodes = nodes
        self.deg_adj_lists = deg_adj_lists
        self.deg_slice = deg_slice
        self.membership = membership
        self.num_mols = num_mols
        self.num_atoms = nodes.shape[0]

    def get_deg_adjacency_lists(self):
        return self.deg_adj_lists

    def get_atom_features(self):
        return self.nodes

    def get_num_atoms(self):
        return self.num_atoms

    def get_num_molecules(self):
        return self.num_mols


def agglomerate(mols, max_deg=MAX_DEG, min_deg=MIN_DEG):
    """Oracle for ``ConvMol.agglomerate_mols`` (mol_graphs.py:256-349).

    ``mols`` may be OracleConvMol objects or reference ConvMol objects (only
    ``atom_features``, ``degree_list`` and ``deg_adj_lists`` are read).

    Output contract: atoms ordered by (degree, molecule,
    in-molecule position); ``deg_slice`` int64 [11,2] with *running* starts (not zeroed
    for empty buckets, mol_graphs.py:300-305); ``membership`` int32 [N];
    ``deg_adj_lists[d]`` int32 [N_d, d] of batch-global row ids.
    """
    num_mols = len(mols)
    feats = np.concatenate([m.atom_features for m in mols])
    degs = np.concatenate([np.asarray(m.degree_list, dtype=np.int64) for m in mols], axis=0)
    n_per = np.array([m.atom_features.shape[0] for m in mols], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(n_per)])
    order = np.argsort(degs, kind="stable")          # mergesort in the reference (:274)
    new_of_old = np.empty(order.shape, np.int32)
    new_of_old[order] = np.arange(order.shape[0], dtype=np.int32)
    nodes = feats[order]

    nb = max_deg - min_deg + 1
    deg_sizes = np.bincount(degs - min_deg, minlength=nb).astype(np.int64) \
        if degs.size else np.zeros(nb, np.int64)
    deg_start = np.concatenate([[0], np.cumsum(deg_sizes)[:-1]])
    deg_slice = np.stack([deg_start, deg_sizes], axis=1)   # int64, like np.array(list(zip()))

    membership = np.empty(nodes.shape[0], np.int32)
    membership[new_of_old] = np.repeat(np.arange(num_mols, dtype=np.int32), n_per)

    deg_adj = []
    for d in range(min_deg, max_deg + 1):
        out = np.empty((int(deg_sizes[d - min_deg]), d), dtype=np.int32)
        row = 0
        for mi, m in enumerate(mols):
            nbr = m.deg_adj_lists[d - min_deg]
            k = nbr.shape[0]
            if k:
                out[row:row + k] = new_of_old[offs[mi]:offs[mi + 1]][nbr]
                row += k
        deg_adj.append(out)
    return OracleMultiConvMol(nodes, deg_adj, deg_slice, membership, num_mols)


def model_inputs(multi, n_samples=None):
    """The list ``GraphConvModel.default_generator`` yields
    (torch_models/graphconvmodel.py:414-421): ``[features f64, deg_slice i64,
    membership i32, n_samples 0-d, deg_adj_1 .. deg_adj_10]``."""
    if n_samples is None:
        n_samples = multi.num_mols
    return [multi.get_atom_features(), multi.deg_slice, np.array(multi.membership),
            np.array(n_samples)] + list(multi.get_deg_adjacency_lists()[1:])


def derived_topology(deg_slice, membership, deg_adj_lists, num_mols):
    """Integer structures the CUDA path derives from the reference layout; restated here
    so the host builder can be checked bit-exactly.

    Returns dict with
      row_ptr [N+1] i32, col_idx [E] i32       CSR of the degree-sorted rows (col_idx is the
                                               concatenation of the flattened deg_adj lists)
      t_row_ptr [N+1], t_src [E], t_slot [E]   transposed CSR: for source row j the list of
                                               (row i, slot k) with col_idx[row_ptr[i]+k]==j,
                                               ordered by (i, k)
      mol_ptr [B+1], mol_atoms [N]             atoms of each molecule in ascending row order
    """
    deg_slice = np.asarray(deg_slice)
    n = int(deg_slice[:, 1].sum())
    degs = np.repeat(np.arange(deg_slice.shape[0], dtype=np.int64), deg_slice[:, 1])
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum(degs)
    col_idx = np.concatenate([np.asarray(a, dtype=np.int32).reshape(-1)
                              for a in deg_adj_lists]) if len(deg_adj_lists) else np.zeros(0, np.int32)
    e = col_idx.shape[0]
    src = np.repeat(np.arange(n, dtype=np.int32), degs)
    slot = (np.arange(e, dtype=np.int64) - row_ptr[src]).astype(np.int32)
    order = np.argsort(col_idx, kind="stable")        # (j, then i, then k)
    t_src = src[order]
    t_slot = slot[order]
    t_row_ptr = np.zeros(n + 1, dtype=np.int64)
    t_row_ptr[1:] = np.cumsum(np.bincount(col_idx, minlength=n))
    membership = np.asarray(membership)
    morder = np.argsort(membership, kind="stable").astype(np.int32)
    mol_ptr = np.zeros(num_mols + 1, dtype=np.int64)
    mol_ptr[1:] = np.cumsum(np.bincount(membership, minlength=num_mols))
    return dict(row_ptr=row_ptr.astype(np.int32), col_idx=col_idx.astype(np.int32),
                t_row_ptr=t_row_ptr.astype(np.int32), t_src=t_src.astype(np.int32),
                t_slot=t_slot.astype(np.int32), mol_ptr=mol_ptr.astype(np.int32),
                mol_atoms=morder)

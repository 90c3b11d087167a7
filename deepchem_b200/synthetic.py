"""Seeded synthetic molecule streams shaped like the datasets BASELINE.json names.

No RDKit and no network exist on the build or GPU boxes, so every benchmark / parity input
is generated here directly at the featurizer's *output* contract
(``ConvMol(atom_features, adj_list)``, deepchem/feat/graph_features.py:698 ->
deepchem/feat/mol_graphs.py:48) in the packed shard format the host layout builder
consumes (see ``PackedMols``).

Shapes:
  zinc      atoms/mol ~ clip(Poisson(25), 6, 50), random spanning tree with valence cap 4
            plus 0-3 ring closures (mean degree ~2.1-2.2), 75 0/1 features (~8 ones/row)
  stress    as zinc but with a heavy-degree tail up to 10 and degree-0 singleton
            molecules, so that all 11 degree buckets are populated
  delaney   13.3 atoms/mol,  tox21  18.5 atoms/mol,  qm9  4-9 atoms/mol
"""
import numpy as np


def _ranges(starts, lens):
    """concatenate([arange(s, s + l) for s, l in zip(starts, lens)]) without the Python loop."""
    lens = np.asarray(lens, dtype=np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    first = np.cumsum(lens) - lens                      # position of every range in the output
    return np.arange(total, dtype=np.int64) + np.repeat(np.asarray(starts, dtype=np.int64) - first, lens)


class LazyTake(object):
    """Molecule numbers into a source shard, gathered later by whoever consumes the batch (the layout worker threads
    of the model's host pipeline, into pinned staging memory) instead of by the dataset iterator."""

    def __init__(self, src, idx):
        self.src, self.idx = src, np.ascontiguousarray(idx, dtype=np.int64)

    def __len__(self):
        return int(self.idx.shape[0])

    n_mols = property(__len__)

    def take(self, reps):
        return LazyTake(self.src, self.idx[np.asarray(reps, dtype=np.int64)])

    def resolve(self, **kw):
        return self.src.take(self.idx, **kw)


class PackedMols(object):
    """A shard of molecules with no Python objects inside.

    atom_ptr [B+1] int32   first atom of each molecule in the concatenated arrays
    adj_ptr  [N+1] int32   CSR over atoms (concatenated, in each molecule's own order)
    adj_idx  [E]   int32   neighbour ids, molecule-local, in adjacency-list order
    features [N,F] float32
    features_i8 [N,F] int8 or None: the same matrix when every entry is an integer in [-128, 127] (``compact()``;
             true of every ConvMol feature: one-hots, formal charge, radical electrons) — a quarter of the bytes to
             store and to upload, converted back to fp32 exactly on the device
    """

    __slots__ = ("atom_ptr", "adj_ptr", "adj_idx", "features", "_pin", "features_i8", "_pin_i8")

    def __init__(self, atom_ptr, adj_ptr, adj_idx, features):
        self.atom_ptr = np.ascontiguousarray(atom_ptr, dtype=np.int32)
        self.adj_ptr = np.ascontiguousarray(adj_ptr, dtype=np.int32)
        self.adj_idx = np.ascontiguousarray(adj_idx, dtype=np.int32)
        self.features = np.ascontiguousarray(features, dtype=np.float32)
        self._pin = None
        self.features_i8 = None
        self._pin_i8 = None

    @property
    def n_mols(self):
        return self.atom_ptr.shape[0] - 1

    @property
    def n_atoms(self):
        return int(self.atom_ptr[-1])

    @property
    def n_feat(self):
        return self.features.shape[1]

    def __len__(self):
        return self.n_mols

    def mol(self, i):
        """(features [n,F], adj_list list-of-lists) of molecule i."""
        a0, a1 = int(self.atom_ptr[i]), int(self.atom_ptr[i + 1])
        adj = [self.adj_idx[self.adj_ptr[a]:self.adj_ptr[a + 1]].tolist() for a in range(a0, a1)]
        return self.features[a0:a1], adj

    def to_list(self):
        return [self.mol(i) for i in range(self.n_mols)]

    def slice(self, lo, hi):
        """Molecules [lo, hi) as a new PackedMols.  The feature rows and neighbour ids are VIEWS
        of this shard (a contiguous molecule range is a contiguous row range), so a pinned shard
        yields pinned batches with no host copy; only the two small offset arrays are rebased."""
        if lo == 0 and hi == self.n_mols:
            return self
        a0, a1 = int(self.atom_ptr[lo]), int(self.atom_ptr[hi])
        e0, e1 = int(self.adj_ptr[a0]), int(self.adj_ptr[a1])
        out = PackedMols.__new__(PackedMols)
        out.atom_ptr = self.atom_ptr[lo:hi + 1] - np.int32(a0)
        out.adj_ptr = self.adj_ptr[a0:a1 + 1] - np.int32(e0)
        out.adj_idx = self.adj_idx[e0:e1]
        out.features = self.features[a0:a1]
        pin = getattr(self, "_pin", None)
        out._pin = pin[a0:a1] if pin is not None else None     # torch view of the same pinned rows
        f8, p8 = getattr(self, "features_i8", None), getattr(self, "_pin_i8", None)
        out.features_i8 = f8[a0:a1] if f8 is not None else None
        out._pin_i8 = p8[a0:a1] if p8 is not None else None
        return out

    def compact(self):
        """Add the int8 copy of the feature matrix when it is exact (one pass over the shard, done once when the
        dataset is packed).  Returns True when the shard is compact."""
        if getattr(self, "features_i8", None) is not None:
            return True
        f = self.features
        if f.size == 0:
            return False
        ok = True
        step = 1 << 16                                        # row blocks: no shard-sized temporaries
        out = np.empty(f.shape, dtype=np.int8)
        for r in range(0, f.shape[0], step):
            blk = f[r:r + step]
            if not np.isfinite(blk).all():
                ok = False
                break
            q = np.clip(blk, -128, 127).astype(np.int8)
            if not np.array_equal(q.astype(np.float32), blk):
                ok = False
                break
            out[r:r + step] = q
        if not ok:
            return False
        self.features_i8 = out
        self._pin_i8 = None
        return True

    def pin_memory(self, compact=True):
        """Move the feature matrix into page-locked host memory (needs torch + CUDA) so that H2D
        copies of batches are asynchronous DMA with no staging copy.  ``compact``: also keep the exact int8
        copy (see ``compact()``) pinned; batches then upload a quarter of the bytes."""
        import torch
        if getattr(self, "_pin", None) is None:
            t = torch.empty(self.features.shape, dtype=torch.float32, pin_memory=True)
            t.numpy()[...] = self.features
            self.features = t.numpy()
            self._pin = t
        if compact and self.compact() and getattr(self, "_pin_i8", None) is None:
            t8 = torch.empty(self.features_i8.shape, dtype=torch.int8, pin_memory=True)
            t8.numpy()[...] = self.features_i8
            self.features_i8 = t8.numpy()
            self._pin_i8 = t8
        return self

    def take(self, idx, alloc=None, prefer_i8=False, n_threads=1):
        """Molecules idx[0], idx[1], ... (repeats allowed) as a new PackedMols: one memcpy per molecule and array in C
        (``dcgc_packed_take``) — a shuffled epoch takes this path for every batch.  ``alloc(n_atoms, n_feat, dtype)``
        may return a page-locked torch tensor for the gathered features (the host pipeline's staging ring), which then
        upload with asynchronous DMA; ``prefer_i8``: when the shard has its exact int8 copy, gather only that one (the
        result's ``features`` is then the int8 matrix)."""
        import ctypes
        from . import _lib
        L = _lib.lib()
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        n = int(idx.shape[0])
        atom_ptr = np.ascontiguousarray(self.atom_ptr, dtype=np.int32)
        adj_ptr = np.ascontiguousarray(self.adj_ptr, dtype=np.int32)
        adj_idx = np.ascontiguousarray(self.adj_idx, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)     # noqa: E731
        na, ne = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.dcgc_packed_take_plan(n, p(idx), self.n_mols, p(atom_ptr), p(adj_ptr), ctypes.byref(na),
                                           ctypes.byref(ne)))
        na, ne = na.value, ne.value
        f = self.n_feat
        i8 = self.features_i8 if getattr(self, "features_i8", None) is not None else None
        only_i8 = bool(prefer_i8 and i8 is not None)

        def out(dtype):
            if alloc is not None:
                t = alloc(na, f, dtype)
                return t.numpy(), t
            return np.empty((na, f), dtype), None
        o32 = o8 = t32 = t8 = None
        src32 = src8 = None
        if not only_i8:
            o32, t32 = out(np.float32)
            src32 = np.ascontiguousarray(self.features, dtype=np.float32)
        if i8 is not None:
            o8, t8 = out(np.int8)
            src8 = np.ascontiguousarray(i8)
        o_atom, o_adjp, o_adj = np.empty(n + 1, np.int32), np.empty(na + 1, np.int32), np.empty(ne, np.int32)
        _lib.check(L.dcgc_packed_take(
            n, p(idx), self.n_mols, p(atom_ptr), p(adj_ptr), p(adj_idx), p(src32) if src32 is not None else None, 4 * f,
            p(src8) if src8 is not None else None, f, p(o_atom), p(o_adjp), p(o_adj),
            p(o32) if o32 is not None else None, p(o8) if o8 is not None else None, int(n_threads)))
        res = PackedMols.__new__(PackedMols)
        res.atom_ptr, res.adj_ptr, res.adj_idx = o_atom, o_adjp, o_adj
        res.features = o32 if o32 is not None else o8
        res._pin, res.features_i8, res._pin_i8 = t32, o8, t8
        return res

    def take_lazy(self, idx):
        return LazyTake(self, idx)

    # ------------------------------------------------------------------ on-disk shard format
    # One directory per shard with four raw .npy files (no pickled Python objects, unlike the reference's
    # object-array shard-N-X.npy, deepchem/data/datasets.py:1359-1427): they memory-map, so a DiskDataset-like
    # iterator can hand contiguous molecule ranges to the layout builder without deserialising anything.
    _FILES = ("atom_ptr", "adj_ptr", "adj_idx", "features")

    def save(self, path):
        import os
        os.makedirs(path, exist_ok=True)
        for name in self._FILES:
            np.save(os.path.join(path, name + ".npy"), getattr(self, name))
        if getattr(self, "features_i8", None) is not None:
            np.save(os.path.join(path, "features_i8.npy"), self.features_i8)
        return path

    @staticmethod
    def load(path, mmap=True):
        import os
        arrs = [np.load(os.path.join(path, name + ".npy"), mmap_mode="r" if mmap else None)
                for name in PackedMols._FILES]
        out = PackedMols.__new__(PackedMols)
        out.atom_ptr, out.adj_ptr, out.adj_idx, out.features = arrs
        out.features_i8 = out._pin_i8 = None
        p8 = os.path.join(path, "features_i8.npy")
        if os.path.exists(p8):
            out.features_i8 = np.load(p8, mmap_mode="r" if mmap else None)
            if out.features_i8.dtype != np.int8 or out.features_i8.shape != out.features.shape:
                raise ValueError("not a PackedMols shard: %s" % path)
        if not (out.atom_ptr.dtype == np.int32 and out.adj_ptr.dtype == np.int32 and out.adj_idx.dtype == np.int32
                and out.features.dtype == np.float32 and out.features.ndim == 2):
            raise ValueError("not a PackedMols shard: %s" % path)
        out._pin = None
        return out

    @staticmethod
    def concat(shards):
        """Concatenate shards into one."""
        atom_ptr, adj_ptr = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
        a_off = e_off = 0
        for s in shards:
            atom_ptr.append(s.atom_ptr[1:].astype(np.int64) + a_off)
            adj_ptr.append(s.adj_ptr[1:].astype(np.int64) + e_off)
            a_off += s.n_atoms
            e_off += int(s.adj_ptr[-1])
        return PackedMols(np.concatenate(atom_ptr), np.concatenate(adj_ptr),
                          np.concatenate([s.adj_idx for s in shards]),
                          np.concatenate([s.features for s in shards]))

    @staticmethod
    def from_list(mols, n_feat=None):
        """mols: iterable of (features [n,F], adj_list)."""
        atom_ptr, adj_ptr, adj_idx, feats = [0], [0], [], []
        for f, adj in mols:
            f = np.asarray(f, dtype=np.float32)
            feats.append(f.reshape(len(adj), -1) if f.size else f.reshape(0, n_feat or 0))
            atom_ptr.append(atom_ptr[-1] + len(adj))
            for nb in adj:
                adj_idx.extend(int(k) for k in nb)
                adj_ptr.append(len(adj_idx))
        if feats:
            features = np.concatenate(feats, 0)
        else:
            features = np.zeros((0, n_feat or 0), np.float32)
        return PackedMols(atom_ptr, adj_ptr, np.asarray(adj_idx, dtype=np.int32), features)


def _random_mol_adj(rng, n, cap, ring_max):
    """Random spanning tree with degree cap + ring closures -> adjacency lists."""
    adj = [[] for _ in range(n)]
    for i in range(1, n):
        p = int(rng.integers(0, i))
        if len(adj[p]) >= cap:
            # any earlier atom with free valence (exists: a tree with cap >= 2 has a leaf)
            free = [q for q in range(i) if len(adj[q]) < cap]
            p = free[int(rng.integers(0, len(free)))]
        adj[p].append(i)
        adj[i].append(p)
    n_rings = int(rng.integers(0, ring_max + 1)) if n >= 5 else 0
    for _ in range(n_rings):
        for _try in range(8):
            a, b = (int(v) for v in rng.integers(0, n, size=2))
            if a != b and b not in adj[a] and len(adj[a]) < cap and len(adj[b]) < cap:
                adj[a].append(b)
                adj[b].append(a)
                break
    return adj


def _random_features(rng, n, n_feat, ones_per_row=8):
    f = np.zeros((n, n_feat), dtype=np.float32)
    if n:
        cols = rng.integers(0, n_feat, size=(n, ones_per_row))
        f[np.arange(n)[:, None], cols] = 1.0
    return f


_SHAPES = {
    # name: (mean atoms, min, max, degree cap, max ring closures)
    "zinc": (25.0, 6, 50, 4, 3),
    "pcba": (25.0, 6, 50, 4, 3),
    "tox21": (18.5, 2, 60, 4, 3),
    "delaney": (13.3, 1, 40, 4, 2),
    "qm9": (8.8, 4, 9, 4, 1),
}


def make_molecules(n_mols, seed=0, shape="zinc", n_feat=75, dense_features=False):
    """Return a PackedMols of ``n_mols`` synthetic molecules.

    ``shape='stress'``: zinc-like sizes, degree cap 10, ~5 % single-atom molecules
    (degree-0 bucket) and hub atoms so that degrees 5..10 occur.
    ``dense_features``: N(0,1) features instead of 0/1 (numerics tests).
    """
    rng = np.random.default_rng(seed)
    stress = shape == "stress"
    mean, lo, hi, cap, rings = _SHAPES["zinc" if stress else shape]
    if shape == "qm9":
        sizes = rng.integers(lo, hi + 1, size=n_mols)
    else:
        sizes = np.clip(rng.poisson(mean, size=n_mols), lo, hi)
    mols = []
    for i in range(n_mols):
        n = int(sizes[i])
        if stress and rng.random() < 0.05:
            n = 1
        if n == 1:
            adj = [[]]
        elif stress and rng.random() < 0.3:
            # hub molecule: atom 0 bonded to k others, rest is a capped tree
            adj = _random_mol_adj(rng, n, 4, rings)
            k = int(rng.integers(5, 11))
            others = [j for j in range(1, n) if j not in adj[0]]
            for j in others[:max(0, k - len(adj[0]))]:
                if len(adj[j]) < 10:
                    adj[0].append(j)
                    adj[j].append(0)
        else:
            adj = _random_mol_adj(rng, n, cap, rings)
        if dense_features:
            f = rng.standard_normal((n, n_feat)).astype(np.float32)
        else:
            f = _random_features(rng, n, n_feat)
        mols.append((f, adj))
    return PackedMols.from_list(mols, n_feat)


def make_labels(n_mols, n_tasks, mode="regression", seed=0, n_classes=2, missing=0.0):
    """y [B,T] and w [B,T]; ``missing`` fraction of weights set to 0 (Tox21-like)."""
    rng = np.random.default_rng(seed + 7919)
    if mode == "classification":
        y = rng.integers(0, n_classes, size=(n_mols, n_tasks)).astype(np.float32)
    else:
        y = rng.standard_normal((n_mols, n_tasks)).astype(np.float32)
    w = np.ones((n_mols, n_tasks), dtype=np.float32)
    if missing > 0:
        w[rng.random((n_mols, n_tasks)) < missing] = 0.0
    return y, w

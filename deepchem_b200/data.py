"""Minimal dataset objects with the iteration contract GraphConvModel relies on.

The model accepts any object with DeepChem's ``iterbatches(batch_size, epochs, deterministic,
pad_batches)`` (deepchem/data/datasets.py:843-898) yielding ``(X, y, w, ids)`` where X is an
object array of ConvMol-like molecules — a real ``dc.data.NumpyDataset`` / ``DiskDataset`` works
unchanged.  ``PackedDataset`` is the fast native form: molecules live in a PackedMols shard
(flat arrays, no Python objects) and batches are slices of it.
"""
import math

import numpy as np

from .synthetic import LazyTake, PackedMols


def pad_batch(batch_size, X_b, y_b, w_b, ids_b, lazy=False):
    """Pad by tiling the samples; only the first copy keeps its weights
    (deepchem/data/datasets.py:142-218).  ``lazy``: a packed shard is tiled as an index list (LazyTake) for the
    consumer to gather."""
    n = len(X_b)
    if n == batch_size:
        return X_b, y_b, w_b, ids_b
    reps = np.arange(batch_size) % n
    if isinstance(X_b, PackedMols):
        X_out = X_b.take_lazy(reps) if lazy else X_b.take(reps)
    else:
        X_out = X_b.take(reps) if isinstance(X_b, LazyTake) else X_b[reps]
    y_out = None if y_b is None else y_b[reps]
    ids_out = ids_b[reps]
    if w_b is None:
        w_out = None
    else:
        w_out = np.zeros((batch_size,) + w_b.shape[1:], dtype=w_b.dtype)
        w_out[:n] = w_b
    return X_out, y_out, w_out, ids_out


def pad_features(batch_size, X_b):
    """Tile the features of a short batch up to ``batch_size`` (deepchem/data/datasets.py:83-138)."""
    n = len(X_b)
    if n > batch_size:
        raise ValueError("Cannot pad an array longer than `batch_size`")
    if n == batch_size:
        return X_b
    return X_b[np.arange(batch_size) % n]


class _ArrayDataset(object):
    def _init_arrays(self, n, y, w, ids, n_tasks):
        if y is None:
            y = np.zeros((n, n_tasks or 1), dtype=np.float32)
            w = np.zeros((n, n_tasks or 1), dtype=np.float32) if w is None else w
        y = np.asarray(y)
        if y.ndim == 1:
            y = y.reshape(-1, 1)
        if w is None:
            w = np.ones((n, y.shape[1]) if y.ndim > 1 else (n,), dtype=np.float32)
        w = np.asarray(w)
        if w.ndim == 1:
            w = w.reshape(-1, 1)
        if ids is None:
            ids = np.arange(n)
        self._y, self._w, self._ids = y, w, np.asarray(ids)

    y = property(lambda self: self._y)
    w = property(lambda self: self._w)
    ids = property(lambda self: self._ids)

    def __len__(self):
        return len(self._y)

    def get_shape(self):
        return (len(self),), self._y.shape, self._w.shape, self._ids.shape

    def _take_X(self, idx, contiguous, lazy=False):
        raise NotImplementedError

    def iterbatches(self, batch_size=None, epochs=1, deterministic=False, pad_batches=False, lazy=False):
        """``lazy``: non-contiguous batches of a packed shard come back as ``LazyTake`` (indices only) so that the
        consumer gathers them — the model's layout workers do, in parallel and into pinned memory."""
        n = len(self)
        if batch_size is None:
            batch_size = n
        for _ in range(epochs):
            perm = np.arange(n) if deterministic else np.random.permutation(n)
            for b in range(math.ceil(n / batch_size) if n else 0):
                idx = perm[b * batch_size:min(n, (b + 1) * batch_size)]
                batch = (self._take_X(idx, deterministic, lazy), self._y[idx], self._w[idx], self._ids[idx])
                if pad_batches:
                    batch = pad_batch(batch_size, *batch, lazy=lazy)
                yield batch


class NumpyDataset(_ArrayDataset):
    """X is an object array (or list) of ConvMol-like molecules."""

    def __init__(self, X, y=None, w=None, ids=None, n_tasks=1):
        if not isinstance(X, np.ndarray):
            arr = np.empty(len(X), dtype=object)
            for i, m in enumerate(X):
                arr[i] = m
            X = arr
        self._X = X
        self._init_arrays(len(X), y, w, ids, n_tasks)

    X = property(lambda self: self._X)

    def _take_X(self, idx, contiguous, lazy=False):
        return self._X[idx]

    def select_range(self, lo, hi):
        return NumpyDataset(self._X[lo:hi], self._y[lo:hi], self._w[lo:hi], self._ids[lo:hi])


class PackedDataset(_ArrayDataset):
    """Molecules stored as one PackedMols shard; batches are PackedMols slices."""

    def __init__(self, packed, y=None, w=None, ids=None, n_tasks=1):
        self.packed = packed
        self._init_arrays(packed.n_mols, y, w, ids, n_tasks)

    X = property(lambda self: self.packed)

    def _take_X(self, idx, contiguous, lazy=False):
        if contiguous and len(idx):
            return self.packed.slice(int(idx[0]), int(idx[-1]) + 1)
        return self.packed.take_lazy(idx) if lazy else self.packed.take(idx)

    def save(self, path):
        """Write the dataset as a packed on-disk shard (PackedMols.save + y / w / ids)."""
        import os
        self.packed.save(path)
        for name, a in (("y", self._y), ("w", self._w), ("ids", self._ids)):
            np.save(os.path.join(path, name + ".npy"), a)
        return path

    @staticmethod
    def from_disk(path, mmap=True):
        import os
        ld = lambda n: np.load(os.path.join(path, n + ".npy"), mmap_mode="r" if mmap else None)  # noqa: E731
        return PackedDataset(PackedMols.load(path, mmap), ld("y"), ld("w"), ld("ids"))

    def select_range(self, lo, hi):
        """Molecules [lo, hi) as a dataset sharing this one's memory (inference sharding)."""
        return PackedDataset(self.packed.slice(lo, hi), self._y[lo:hi], self._w[lo:hi], self._ids[lo:hi])


class CSVLoader(object):
    """``dc.data.CSVLoader(tasks, feature_field, featurizer).create_dataset(csv)`` (deepchem/data/data_loader.py:
    ``CSVLoader``, labels / weights as ``_convert_df_to_numpy`` ``:35-69``: weights 1, missing labels -> y = 0, w = 0)
    for SMILES files, over the RDKit-free reader: the molecules go straight into one ``PackedMols`` shard (no pickled
    ``ConvMol`` objects), rows whose SMILES fail to parse are dropped as ``DataLoader`` drops failed datapoints."""

    def __init__(self, tasks, featurizer=None, feature_field="smiles", id_field=None, smiles_field=None, n_jobs=1):
        self.tasks = list(tasks)
        self.n_jobs = n_jobs                  # worker processes of the SMILES reader (smiles.featurize_smiles_packed)
        self.feature_field = smiles_field if smiles_field is not None else feature_field
        self.id_field = id_field
        self.featurizer = featurizer          # accepted for signature compatibility; the 75-dim ConvMol features are built in

    def create_dataset(self, inputs, data_dir=None, shard_size=None):
        import csv
        from .smiles import featurize_smiles_packed
        paths = [inputs] if isinstance(inputs, str) else list(inputs)
        rows = []
        for path in paths:
            if path.endswith(".gz"):                   # pandas (the reference's reader) infers gzip from the suffix
                import gzip
                fh = gzip.open(path, "rt", newline="")
            else:
                fh = open(path, newline="")
            with fh:
                rows.extend(csv.DictReader(fh))
        smiles = [r[self.feature_field].strip() for r in rows]
        packed, bad = featurize_smiles_packed(smiles, n_jobs=self.n_jobs)
        keep = np.setdiff1d(np.arange(len(rows)), np.asarray(bad, dtype=np.int64))
        y = np.zeros((len(keep), len(self.tasks)), np.float32)
        w = np.ones((len(keep), len(self.tasks)), np.float32)
        for i, k in enumerate(keep):
            for t, task in enumerate(self.tasks):
                v = (rows[k].get(task) or "").strip()
                if v == "":
                    w[i, t] = 0.0
                else:
                    y[i, t] = float(v)
        ids = np.asarray([rows[k][self.id_field] if self.id_field else smiles[k] for k in keep], dtype=str)
        ds = PackedDataset(packed, y, w, ids, n_tasks=len(self.tasks))
        if data_dir is not None:
            ds.save(data_dir)
        return ds

    featurize = create_dataset


class ReplayDataset(object):
    """``n`` molecules streamed from a small packed shard replayed cyclically (molecule i of the dataset is molecule
    ``(start + i) % len(shard)`` of the shard): the inference stream of BASELINE configs[4] — 10 M PCBA-shaped
    molecules — without holding 10 M molecules in host memory.  Supports what ``predict`` needs: ``len``,
    ``iterbatches`` (labels are ``None``) and ``select_range`` for rank sharding."""

    def __init__(self, shard, n, start=0):
        self.shard, self.n, self.start = shard, int(n), int(start)

    def __len__(self):
        return self.n

    y = w = None
    ids = property(lambda self: np.arange(self.start, self.start + self.n))

    def select_range(self, lo, hi):
        return ReplayDataset(self.shard, max(0, int(hi) - int(lo)), self.start + int(lo))

    def iterbatches(self, batch_size=None, epochs=1, deterministic=True, pad_batches=False, lazy=False):
        m = self.shard.n_mols
        bs = self.n if batch_size is None else int(batch_size)
        for _ in range(epochs):
            done = 0
            while done < self.n:
                k = min(bs, self.n - done)
                off = (self.start + done) % m
                if off + k <= m:
                    X = self.shard.slice(off, off + k)
                else:
                    idx = (off + np.arange(k)) % m
                    X = self.shard.take_lazy(idx) if lazy else self.shard.take(idx)
                yield X, None, None, np.arange(self.start + done, self.start + done + k)
                done += k

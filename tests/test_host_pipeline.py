"""Host side of the batch pipeline (no GPU): multi-worker layout building keeps dataset order and
produces the same layouts as the single-threaded generator (graphconvmodel.py:382-422 contract)."""
import threading

import numpy as np
import pytest

from deepchem_b200 import graphconvmodel as G
from deepchem_b200.data import PackedDataset
from deepchem_b200.synthetic import make_molecules

FIELDS = ("deg_slice", "membership", "perm", "row_ptr", "col_idx", "t_row_ptr", "t_src", "t_slot", "mol_ptr",
          "mol_atoms", "tiles")


def _host_model(batch_size, workers):
    m = G.GraphConvModel.__new__(G.GraphConvModel)      # host-only: no CUDA objects are created
    m.batch_size, m.mode, m.n_tasks, m.n_classes = batch_size, 'regression', 1, 2
    m.host_workers = workers
    m._staging, m._feat_staging = G._PinnedRing(pin=False), G._PinnedRing(pin=False)
    return m


def test_parallel_generator_matches_serial_and_keeps_order():
    pm = make_molecules(1000, seed=2, shape="stress")
    ds = PackedDataset(pm, np.arange(1000, dtype=np.float32).reshape(-1, 1), np.ones((1000, 1), np.float32))
    m = _host_model(128, 3)
    serial = list(m.default_generator(ds, epochs=2, workers=1))
    par = list(m.default_generator(ds, epochs=2, workers=3))
    assert len(serial) == len(par) == 16
    for (ia, ya, wa), (ib, yb, wb) in zip(serial, par):
        assert np.array_equal(ya[0], yb[0]) and np.array_equal(wa[0], wb[0])
        assert int(ia[3]) == int(ib[3])
        for f in FIELDS:
            assert np.array_equal(getattr(ia.layout, f), getattr(ib.layout, f)), f
    # last batch of each epoch is padded to batch_size with zero weights (datasets.py:142-218)
    assert serial[7][1][0].shape[0] == 128 and float(serial[7][2][0][1000 - 7 * 128:].sum()) == 0.0


def test_parallel_generator_propagates_errors():
    class Bad(object):
        def iterbatches(self, **kw):
            pm = make_molecules(8, seed=1)
            pm.adj_idx = pm.adj_idx.copy()
            pm.adj_idx[0] = 10 ** 6                      # neighbour outside its molecule
            yield pm, np.zeros((8, 1), np.float32), np.ones((8, 1), np.float32), np.arange(8)
    m = _host_model(8, 2)
    try:
        list(m.default_generator(Bad(), workers=2))
    except Exception as e:
        assert "neighbour" in str(e) or "index" in str(e).lower()
    else:
        raise AssertionError("expected the layout builder's index error")


def test_packed_shard_round_trips_through_disk(tmp_path):
    """The packed on-disk shard format (raw .npy arrays, memory-mapped on load) feeds the layout builder with
    the same integers as the in-memory shard."""
    from deepchem_b200.mol_graphs import BatchLayout
    pm = make_molecules(300, seed=4, shape="stress")
    y = np.arange(600, dtype=np.float32).reshape(300, 2)
    ds = PackedDataset(pm, y, np.ones((300, 2), np.float32))
    ds.save(str(tmp_path / "shard0"))
    back = PackedDataset.from_disk(str(tmp_path / "shard0"))
    assert len(back) == 300 and np.array_equal(back.y, y)
    m = _host_model(64, 1)
    a = list(m.default_generator(ds, workers=1))
    b = list(m.default_generator(back, workers=1))
    for (ia, ya, _), (ib, yb, _) in zip(a, b):
        assert np.array_equal(ya[0], yb[0])
        for f in FIELDS:
            assert np.array_equal(getattr(ia.layout, f), getattr(ib.layout, f)), f
        assert np.array_equal(ia.packed_features, ib.packed_features)
    assert np.array_equal(BatchLayout.build(back.packed).col_idx, BatchLayout.build(pm).col_idx)


def test_compact_int8_features_are_exact_or_refused(tmp_path):
    """PackedMols.compact(): the int8 copy exists only when it reproduces the fp32 matrix exactly (every ConvMol
    feature is a one-hot, a formal charge or a radical count: graph_features.py:282-391); slices are views of it
    and it round-trips through the on-disk shard."""
    from deepchem_b200.synthetic import PackedMols
    pm = make_molecules(200, seed=7, shape="stress")
    pm.features[3, 10] = -1.0                                  # a formal charge
    assert pm.compact() and pm.features_i8.dtype == np.int8
    assert np.array_equal(pm.features_i8.astype(np.float32), pm.features)
    part = pm.slice(17, 90)
    assert np.array_equal(part.features_i8.astype(np.float32), part.features)
    assert part.features_i8.base is not None                   # a view, not a copy
    pm.save(str(tmp_path / "s"))
    back = PackedMols.load(str(tmp_path / "s"))
    assert np.array_equal(back.features_i8, pm.features_i8)
    for bad in (0.5, 200.0, -129.0, float("nan")):
        q = make_molecules(50, seed=8)
        q.features[5, 5] = bad
        assert not q.compact() and q.features_i8 is None


def test_packed_take_matches_per_molecule_gather():
    """PackedMols.take (dcgc_packed_take: one memcpy per molecule) against a per-molecule rebuild: repeats, slices of a
    shard, the int8-only form, the empty selection, an out-of-range index."""
    from deepchem_b200.synthetic import PackedMols, make_molecules
    pm = make_molecules(600, seed=2, shape="stress")
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 600, size=777)
    ref = PackedMols.from_list([pm.mol(int(i)) for i in idx], 75)
    for threads in (1, 3):
        got = pm.take(idx, n_threads=threads)
        for k in ("atom_ptr", "adj_ptr", "adj_idx", "features"):
            assert np.array_equal(getattr(got, k), getattr(ref, k)) and getattr(got, k).dtype == getattr(ref, k).dtype, k
    pc = make_molecules(600, seed=2, shape="stress")
    assert pc.compact()
    only8 = pc.take(idx, prefer_i8=True)
    assert only8.features.dtype == np.int8 and np.array_equal(only8.features.astype(np.float32), ref.features)
    both = pc.take(idx)
    assert both.features.dtype == np.float32 and np.array_equal(both.features_i8, only8.features)
    sl = pm.slice(100, 400)
    sub = sl.take(np.array([5, 5, 0, 299]))
    want = PackedMols.from_list([sl.mol(i) for i in (5, 5, 0, 299)], 75)
    for k in ("atom_ptr", "adj_ptr", "adj_idx", "features"):
        assert np.array_equal(getattr(sub, k), getattr(want, k)), k
    assert pm.take(np.zeros(0, np.int64)).n_mols == 0
    with pytest.raises(ValueError):
        pm.take(np.array([600]))


def test_lazy_shuffled_batches_equal_eager_ones():
    """iterbatches(lazy=True) hands out index lists; resolving them (what the layout workers do) gives the batches of
    the eager iterator, padding included."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.synthetic import LazyTake, make_labels, make_molecules
    pm = make_molecules(230, seed=5, shape="zinc")
    y, w = make_labels(230, 2, "regression", seed=1)
    ds = PackedDataset(pm, y, w)
    np.random.seed(11)
    eager = list(ds.iterbatches(batch_size=64, deterministic=False, pad_batches=True))
    np.random.seed(11)
    lazy = list(ds.iterbatches(batch_size=64, deterministic=False, pad_batches=True, lazy=True))
    assert len(eager) == len(lazy) == 4
    for (Xe, ye, we, ie), (Xl, yl, wl, il) in zip(eager, lazy):
        assert isinstance(Xl, LazyTake) and len(Xl) == 64
        Xr = Xl.resolve()
        for k in ("atom_ptr", "adj_ptr", "adj_idx", "features"):
            assert np.array_equal(getattr(Xr, k), getattr(Xe, k)), k
        assert np.array_equal(ye, yl) and np.array_equal(we, wl) and np.array_equal(ie, il)
    assert float(lazy[-1][2][230 - 192:].sum()) == 0.0            # the padded copies carry zero weights


def test_undo_transforms_follows_the_reference_order_and_limits():
    """deepchem.trans.undo_transforms as TorchModel._predict applies it (torch_model.py:625-634): y-transformers are
    undone last-to-first, X-transformers are skipped, several outputs cannot be untransformed."""
    class T(object):
        def __init__(self, transform_y, scale, shift):
            self.transform_y, self.scale, self.shift = transform_y, scale, shift

        def untransform(self, y):
            return y * self.scale + self.shift
    y = np.arange(6, dtype=np.float32).reshape(3, 2)
    assert G._undo_transforms(y, []) is y
    chain = [T(True, 2.0, 1.0), T(False, 100.0, 100.0), T(True, 3.0, -1.0)]      # forward order of the transforms
    assert np.array_equal(G._undo_transforms(y, chain), (y * 3.0 - 1.0) * 2.0 + 1.0)
    assert np.array_equal(G._undo_transforms([y], chain)[0], (y * 3.0 - 1.0) * 2.0 + 1.0)   # a single output in a list
    with pytest.raises(ValueError):
        G._undo_transforms([y, y], chain)


@pytest.mark.parametrize("source", ["convmol_objects", "packed_shard"])
def test_default_generator_equals_the_reference_generator(source):
    """Row a3 against the reference itself (tests/golden/make_golden_generator.py): 23 molecules, batch 10,
    classification with 2 tasks.  Every array the reference's ``default_generator`` yields — permuted features,
    deg_slice (int64), membership, n_samples, deg_adj_1..10, one-hot labels, weights, the third batch padded by
    ``pad_batch`` — is reproduced bit for bit, from an object array of ConvMol and from a packed shard; in 'predict' mode
    with ``pad_batches=False`` the labels stay class indices and the last batch keeps its 3 molecules."""
    from helpers import load_golden, unpack_mols
    from deepchem_b200.data import NumpyDataset
    from deepchem_b200.mol_graphs import ConvMol
    from deepchem_b200.synthetic import PackedMols
    d = load_golden("ref_generator.npz")
    mols = unpack_mols(d)
    if source == "packed_shard":
        ds = PackedDataset(PackedMols.from_list(mols), d["y"], d["w"])
    else:
        X = np.empty(len(mols), dtype=object)
        for i, (f, adj) in enumerate(mols):
            X[i] = ConvMol(np.asarray(f, dtype=np.float64), adj)
        ds = NumpyDataset(X, d["y"], d["w"])
    m = _host_model(10, 1)
    m.mode, m.n_tasks, m.n_classes = 'classification', 2, 2
    for tag, kw in (("fit", dict(mode="fit", pad_batches=True)), ("predict", dict(mode="predict", pad_batches=False)),
                    ("tiny", dict(mode="fit", pad_batches=True))):
        src = ds.select_range(0, 3) if tag == "tiny" else ds      # 3 molecules in a batch of 10: repeated 3 1/3 times
        got = list(m.default_generator(src, epochs=1, deterministic=True, **kw))
        assert len(got) == int(d["%s_batches" % tag])
        for n, (inputs, labels, weights) in enumerate(got):
            lay = inputs.layout
            feats = lay.permute_features(np.asarray(inputs.packed_features, dtype=np.float32))
            assert np.array_equal(feats, d["%s_b%d_in0" % (tag, n)]), (tag, n)
            ref_slice = d["%s_b%d_in1" % (tag, n)]
            assert inputs[1].dtype == ref_slice.dtype == np.int64 and np.array_equal(inputs[1], ref_slice)
            assert np.array_equal(inputs[2], d["%s_b%d_in2" % (tag, n)])
            assert int(inputs[3]) == int(d["%s_b%d_in3" % (tag, n)])
            assert len(inputs) == 14
            for k in range(4, 14):
                ref_adj = d["%s_b%d_in%d" % (tag, n, k)]
                assert inputs[k].shape == ref_adj.shape and np.array_equal(inputs[k], ref_adj), (tag, n, k)
            assert np.array_equal(np.asarray(labels[0], dtype=np.float32), d["%s_b%d_y" % (tag, n)]), (tag, n)
            assert np.array_equal(np.asarray(weights[0], dtype=np.float32), d["%s_b%d_w" % (tag, n)]), (tag, n)


@pytest.mark.parametrize("fixture,kw", [("ref_model_classification.npz", dict(mode="classification")),
                                        ("ref_model_regression.npz", dict(mode="regression")),
                                        ("ref_model_uncertainty.npz", dict(mode="regression", uncertainty=True, dropout=0.25)),
                                        ("ref_tox21_real.npz", dict(mode="classification", n_tasks=12))])
def test_module_state_dict_is_the_reference_checkpoint_layout(fixture, kw):
    """SURVEY 8b checkpoint compatibility, without a device: the module's state_dict has exactly the reference's keys,
    shapes and dtypes (fixtures hold the reference model's own state_dict) and loads it strictly — also with the
    opt-in synchronised BatchNorm."""
    import torch
    from helpers import load_golden
    d = load_golden(fixture)
    sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
    kw = dict(kw)
    n_tasks = kw.pop("n_tasks", 3)
    for sync in (False, True):
        m = G._GraphConvTorchModel(n_tasks, graph_conv_layers=[64, 64], dense_layer_size=128, batch_size=32,
                                   sync_batch_norm=sync, **kw)
        own = m.state_dict()
        assert list(own.keys()) == list(sd.keys())
        for k, v in own.items():
            assert tuple(v.shape) == tuple(sd[k].shape) and v.dtype == sd[k].dtype, k
        m.load_state_dict(sd, strict=True)
        assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())


def test_constructor_errors_follow_the_reference():
    """graphconvmodel.py:121-140: ValueError for an unknown mode, a dropout list of the wrong length, uncertainty
    outside regression and uncertainty without dropout in every layer."""
    make = G._GraphConvTorchModel
    with pytest.raises(ValueError, match="mode must be either"):
        make(2, mode="ranking")
    with pytest.raises(ValueError, match="Wrong number of dropout probabilities"):
        make(2, graph_conv_layers=[64, 64], dropout=[0.1, 0.1])
    with pytest.raises(ValueError, match="only supported in regression"):
        make(2, mode="classification", uncertainty=True, dropout=0.2)
    with pytest.raises(ValueError, match="Dropout must be included"):
        make(2, mode="regression", uncertainty=True, dropout=[0.2, 0.0, 0.2])
    m = make(2, graph_conv_layers=[64, 64], dropout=[0.1, 0.2, 0.3], mode="regression", uncertainty=True)
    assert [d.p for d in m.dropouts] == [0.1, 0.2, 0.3] and hasattr(m, "uncertainty_dense")
    assert len(m.graph_convs[0].W_list) == 21 and tuple(m.graph_convs[0].W_list[0].shape) == (75, 64)    # layers.py:6140-6151


def test_staging_ring_never_overwrites_a_batch_that_is_still_referenced():
    """The reference's generator yields independent arrays (graphconvmodel.py:382-422): a materialised list of
    batches, or a consumer lagging behind, must not see earlier batches change.  The pinned staging ring hands a
    slot out again only after every array built in it is gone."""
    import gc
    ring = G._PinnedRing(max_slots=4, pin=False)
    held = []
    for i in range(4):
        slot, root = ring.take(1024)
        root[:8] = i
        held.append(root[:8])                     # a VIEW keeps the slot busy, not only the root
        del root
    assert len(ring.slots) == 4
    assert ring.take(1024) == (None, None)        # every slot referenced: the caller must use ordinary memory
    assert [int(h[0]) for h in held] == [0, 1, 2, 3]
    held.pop(1)
    gc.collect()
    slot, root = ring.take(1024)                  # the released slot is found again, the others are untouched
    assert slot is ring.slots[1]
    root[:8] = 9
    assert [int(h[0]) for h in held] == [0, 2, 3]
    # a slot grows when a larger batch needs it
    del root
    gc.collect()
    slot, root = ring.take(1 << 20)
    assert root.shape[0] >= 1 << 20


def test_materialised_generator_keeps_every_batch_intact():
    """list(default_generator(ds)) with more batches than the ring holds: every layout still equals a fresh build."""
    from deepchem_b200.mol_graphs import BatchLayout
    pm = make_molecules(600, seed=9, shape="stress")
    ds = PackedDataset(pm, np.zeros((600, 1), np.float32), np.ones((600, 1), np.float32))
    m = _host_model(16, 2)
    m._staging = G._PinnedRing(max_slots=6, pin=False)
    real_take = m._staging.take
    m.batch_inputs = lambda X_b, pinned=True: _batch_inputs_with_ring(m, X_b)
    batches = list(m.default_generator(ds, workers=2))
    assert len(batches) == 38 > 6
    for k, (inp, _, _) in enumerate(batches):
        fresh = BatchLayout.build(pm.take(np.arange(16 * k, min(600, 16 * k + 16))) if 16 * k + 16 <= 600
                                  else inp.layout_source, n_segments=16)
        for f in FIELDS:
            assert np.array_equal(getattr(inp.layout, f), getattr(fresh, f)), (k, f)
    assert real_take is not None


def _batch_inputs_with_ring(m, X_b):
    """batch_inputs with the ring in use although no GPU is present (the product only uses it for pinned memory)."""
    from deepchem_b200.mol_graphs import BatchLayout
    packed = X_b.resolve() if isinstance(X_b, G.LazyTake) else X_b
    slot, root = m._staging.take(1 << 16)
    layout = BatchLayout.build(packed, n_segments=max(m.batch_size, packed.n_mols),
                               staging=slot[0] if slot is not None else None, staging_root=root)
    inputs = G.BatchInputs([None, layout.deg_slice, layout.membership, np.array(packed.n_mols)]
                           + layout.deg_adjacency_lists()[1:])
    inputs.layout = layout
    inputs.layout_source = packed
    return inputs


def test_replay_dataset_streams_and_shards():
    """ReplayDataset (the 10 M-molecule inference stream of BASELINE configs[4] without 10 M molecules in memory):
    molecule i is molecule (start + i) % len(shard); rank shards are contiguous and cover the stream exactly once."""
    from deepchem_b200.data import ReplayDataset
    from deepchem_b200.parallel import shard_range
    shard = make_molecules(24, seed=3, shape="delaney")
    ds = ReplayDataset(shard, 100)
    assert len(ds) == 100
    seen = []
    for rank in range(3):
        lo, hi = shard_range(len(ds), rank, 3)
        part = ds.select_range(lo, hi)
        assert len(part) == hi - lo
        for X, y, w, ids in part.iterbatches(batch_size=16, deterministic=True):
            assert y is None and w is None and X.n_mols == len(ids) <= 16
            for j, i in enumerate(ids):
                a = X.mol(j)
                b = shard.mol(int(i) % 24)
                assert np.array_equal(a[0], b[0]) and a[1] == b[1]
            seen.extend(int(i) for i in ids)
    assert seen == list(range(100))

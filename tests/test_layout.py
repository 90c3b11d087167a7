"""Host layout builder (C++ behind the C ABI) against the reference goldens and the oracle:
integer outputs must match bit for bit (dtype included)."""
import numpy as np
import pytest

from helpers import load_golden, oracle_batch, unpack_mols
from deepchem_b200 import mol_graphs as MG
from deepchem_b200.synthetic import PackedMols, make_molecules
from oracle.convmol_layout import derived_topology


def _packed(d):
    return PackedMols(d["atom_ptr"], d["adj_ptr"], d["adj_idx"], d["features"].astype(np.float32))


@pytest.mark.parametrize("name", ["kat_ccc_c.npz", "ref_layout_stress.npz", "ref_layout_zinc.npz",
                                  "ref_layout_delaney.npz"])
def test_builder_matches_reference_golden(name):
    d = load_golden(name)
    lay = MG.BatchLayout.build(_packed(d))
    assert lay.deg_slice.dtype == np.int64 and np.array_equal(lay.deg_slice, d["ref_deg_slice"])
    assert lay.membership.dtype == np.int32 and np.array_equal(lay.membership, d["ref_membership"])
    for k, a in enumerate(lay.deg_adjacency_lists()):
        r = d["ref_deg_adj_%d" % k]
        assert a.dtype == np.int32 and a.shape == r.shape and np.array_equal(a, r), k
    feats = lay.permute_features(d["features"].astype(np.float32))
    assert np.array_equal(feats, d["ref_nodes"].astype(np.float32))
    padded = lay.permute_features(d["features"].astype(np.float32), ld_out=76)
    assert np.array_equal(padded[:, :75], feats) and np.all(padded[:, 75] == 0)


@pytest.mark.parametrize("name", ["kat_ccc_c.npz", "ref_layout_stress.npz"])
def test_convmol_class_matches_reference_golden(name):
    """ConvMol.__init__ outputs (a1) and the agglomerate_mols drop-in (a2)."""
    d = load_golden(name)
    cms = [MG.ConvMol(f, adj) for f, adj in unpack_mols(d)]
    ds = np.stack([c.deg_slice for c in cms])
    assert ds.dtype == d["ref_mol_deg_slice"].dtype and np.array_equal(ds, d["ref_mol_deg_slice"])
    assert np.array_equal(np.concatenate([c.deg_block_indices for c in cms]), d["ref_mol_deg_block_indices"])
    assert np.array_equal(np.concatenate([np.asarray(c.degree_list, np.int32) for c in cms]), d["ref_mol_degree_list"])
    flat = np.asarray([k for c in cms for nb in c.get_adjacency_list() for k in nb], np.int32)
    assert np.array_equal(flat, d["ref_mol_canon_adj_flat"])
    assert np.array_equal(np.concatenate([c.get_atom_features() for c in cms]), d["ref_mol_features_sorted"])
    mm = MG.ConvMol.agglomerate_mols(cms)
    assert np.array_equal(mm.deg_slice, d["ref_deg_slice"]) and mm.deg_slice.dtype == np.int64
    assert np.array_equal(mm.membership, d["ref_membership"])
    assert np.array_equal(mm.get_atom_features(), d["ref_nodes"])
    assert mm.get_atom_features().dtype == d["ref_nodes"].dtype
    for k, a in enumerate(mm.get_deg_adjacency_lists()):
        assert np.array_equal(a, d["ref_deg_adj_%d" % k])
    assert mm.get_num_atoms() == d["ref_nodes"].shape[0] and mm.get_num_molecules() == len(cms)


def test_null_mol_self_loops():
    d = load_golden("ref_layout_nullmol.npz")
    adj = [deg * [deg] for deg in range(11)]
    nm = MG.ConvMol(d["features"], adj)
    mm = MG.ConvMol.agglomerate_mols([nm, nm])
    assert np.array_equal(mm.deg_slice, d["deg_slice"]) and np.array_equal(mm.membership, d["membership"])
    for k, a in enumerate(mm.get_deg_adjacency_lists()):
        assert np.array_equal(a, d["deg_adj_%d" % k])


@pytest.mark.parametrize("shape,n,seed", [("stress", 300, 1), ("zinc", 257, 2), ("qm9", 100, 3)])
def test_builder_matches_oracle_on_random_batches(shape, n, seed):
    pm = make_molecules(n, seed=seed, shape=shape)
    _, mm = oracle_batch(pm.to_list())
    nseg = n + 5
    lay = MG.BatchLayout.build(pm, n_segments=nseg)
    assert np.array_equal(lay.deg_slice, mm.deg_slice) and np.array_equal(lay.membership, mm.membership)
    for a, r in zip(lay.deg_adjacency_lists(), mm.get_deg_adjacency_lists()):
        assert np.array_equal(a, r)
    t = derived_topology(mm.deg_slice, mm.membership, mm.get_deg_adjacency_lists(), nseg)
    for k in ("row_ptr", "col_idx", "t_row_ptr", "t_src", "t_slot", "mol_ptr", "mol_atoms"):
        assert np.array_equal(getattr(lay, k), t[k]), k
    # perm: row i of the batch is atom perm[i] of the shard
    assert np.array_equal(pm.features[lay.perm], np.asarray(mm.get_atom_features(), np.float32))
    # tiles: cover every row exactly once, never straddle a degree bucket
    rows = np.concatenate([np.arange(r0, r0 + nr) for r0, nr, _, _ in lay.tiles])
    assert np.array_equal(rows, np.arange(lay.n_atoms))
    deg_of_row = np.repeat(np.arange(11), lay.deg_slice[:, 1])
    for r0, nr, dg, _ in lay.tiles:
        assert 0 < nr <= 128 and np.all(deg_of_row[r0:r0 + nr] == dg)
    # the same slab derived from the reference arrays alone
    lay2 = MG.BatchLayout.from_reference_arrays(mm.deg_slice, mm.membership, mm.get_deg_adjacency_lists(), nseg)
    for k in ("row_ptr", "col_idx", "t_row_ptr", "t_src", "t_slot", "mol_ptr", "mol_atoms", "tiles", "deg_slice"):
        assert np.array_equal(getattr(lay, k), getattr(lay2, k)), k


def test_empty_and_degenerate_batches():
    empty = PackedMols([0], [0], [], np.zeros((0, 75), np.float32))
    lay = MG.BatchLayout.build(empty, n_segments=4)
    assert lay.n_atoms == 0 and lay.n_tiles == 0 and lay.mol_ptr.tolist() == [0] * 5
    assert lay.deg_slice.tolist() == [[0, 0]] * 11
    single = PackedMols.from_list([(np.ones((1, 75), np.float32), [[]])])
    lay = MG.BatchLayout.build(single)
    assert lay.deg_slice[0].tolist() == [0, 1] and lay.membership.tolist() == [0]
    assert [a.shape for a in lay.deg_adjacency_lists()][:2] == [(1, 0), (0, 1)]


def test_builder_errors():
    f = np.zeros((12, 75), np.float32)
    too_many = [list(range(1, 12))] + [[0]] * 11          # atom 0 has degree 11
    with pytest.raises(ValueError, match="degree"):
        MG.BatchLayout.build(PackedMols.from_list([(f, too_many)]))
    bad_index = PackedMols.from_list([(np.zeros((2, 75), np.float32), [[5], [0]])])
    with pytest.raises(ValueError, match="neighbour index"):
        MG.BatchLayout.build(bad_index)
    with pytest.raises(ValueError):
        MG.BatchLayout.build(PackedMols.from_list([(np.zeros((1, 75), np.float32), [[]])]), n_segments=0)


def test_asymmetric_adjacency_master_atom():
    """A fake atom appended to every real atom's list but with no list of its own
    (feat/graph_features.py:906-909): CSR^T must still be the exact transpose."""
    adj = [[1, 3], [0, 2, 3], [1, 3], []]
    lay = MG.BatchLayout.build(PackedMols.from_list([(np.zeros((4, 75), np.float32), adj)]))
    for j in range(lay.n_atoms):
        for e in range(lay.t_row_ptr[j], lay.t_row_ptr[j + 1]):
            assert lay.col_idx[lay.row_ptr[lay.t_src[e]] + lay.t_slot[e]] == j
    assert lay.t_row_ptr[-1] == lay.n_edges == 7


# ---------------------------------------------------------------------------- molecule-group table
def _check_groups(lay, R):
    """The table the staged kernels rely on: every group is one contiguous row range per degree bucket, those
    ranges hold exactly the rows of the group's molecules, and every neighbour of a row is in its group."""
    g = lay.groups
    n_groups, max_rows, max_entries, valid, rows_budget = (int(v) for v in g[0, :5])
    assert valid == 1 and rows_budget == R
    N = lay.n_atoms
    tab = g[1:n_groups + 2, :11].astype(np.int64)
    starts = lay.deg_slice[:, 0] if N else np.zeros(11, np.int64)
    counts = np.asarray(lay.deg_count, np.int64)
    bstart = np.concatenate([[0], np.cumsum(counts)])[:11]
    assert np.array_equal(tab[0], bstart) and np.array_equal(tab[-1], bstart + counts)
    assert (np.diff(tab, axis=0) >= 0).all()
    group_of_row = np.full(N, -1, np.int64)
    seen_rows = seen_entries = 0
    for k in range(n_groups):
        rows = np.concatenate([np.arange(tab[k, d], tab[k + 1, d]) for d in range(11)])
        group_of_row[rows] = k
        n_ent = int(sum((tab[k + 1, d] - tab[k, d]) * d for d in range(11)))
        seen_rows, seen_entries = max(seen_rows, rows.size), max(seen_entries, n_ent)
    assert (group_of_row >= 0).all() and seen_rows == max_rows and seen_entries == max_entries
    # greedy packing of consecutive molecules: the group is a non-decreasing function of the molecule, a group
    # holds at most R rows unless it is a single molecule, and two consecutive groups never fit in one
    sizes = np.diff(lay.mol_ptr).astype(np.int64)
    mol_group = np.full(sizes.size, -1, np.int64)
    mol_group[lay.membership] = group_of_row
    assert all(np.unique(group_of_row[lay.membership == m]).size == 1 for m in np.unique(lay.membership)[:50])
    mg = mol_group[sizes > 0]
    assert (np.diff(mg) >= 0).all() and (n_groups == 0 or (mg[0] == 0 and mg[-1] == n_groups - 1))
    rows_of_group = np.bincount(group_of_row, minlength=n_groups)
    mols_of_group = np.bincount(mg, minlength=n_groups)
    assert ((rows_of_group <= R) | (mols_of_group == 1)).all() and (rows_of_group > 0).all()
    first_mol_rows = sizes[sizes > 0][np.searchsorted(mg, np.arange(n_groups))]
    assert (rows_of_group[:-1] + first_mol_rows[1:] > R).all()
    assert np.array_equal(g[1:n_groups + 2, 11], np.concatenate([np.flatnonzero(sizes > 0)[np.searchsorted(mg, np.arange(n_groups))], [lay.n_segments]]))
    rows_of_entry = np.repeat(np.arange(N), np.diff(lay.row_ptr))
    assert np.array_equal(group_of_row[lay.col_idx], group_of_row[rows_of_entry])
    assert np.array_equal(group_of_row[lay.t_src], group_of_row[np.repeat(np.arange(N), np.diff(lay.t_row_ptr))])
    del starts


@pytest.mark.parametrize("shape,n,segs", [("zinc", 300, 300), ("stress", 257, 300), ("delaney", 5, 5),
                                          ("zinc", 1, 2)])
def test_group_table(shape, n, segs):
    from deepchem_b200.synthetic import make_molecules
    from deepchem_b200.mol_graphs import BatchLayout
    pm = make_molecules(n, seed=11, shape=shape)
    lay = BatchLayout.build(pm, n_segments=segs)
    _check_groups(lay, int(lay.info.group_rows))
    # the same table from the reference-shaped arrays (GraphConv layers handed plain tensors)
    lay2 = BatchLayout.from_reference_arrays(lay.deg_slice, lay.membership, lay.deg_adjacency_lists(), segs)
    assert np.array_equal(lay2.groups, lay.groups)


def test_group_table_rejects_foreign_layouts():
    """Hand-made layouts without the (degree, molecule) row order, or with an edge between molecules, get no
    group table (n_groups == 0), which switches the staged kernels off."""
    from deepchem_b200.synthetic import make_molecules
    from deepchem_b200.mol_graphs import BatchLayout
    pm = make_molecules(40, seed=3, shape="zinc")
    lay = BatchLayout.build(pm)
    mem = lay.membership.copy()
    d2 = np.flatnonzero(np.diff(mem[:lay.deg_count[0] + lay.deg_count[1]]) > 0)
    i = int(d2[0])
    mem[i], mem[i + 1] = mem[i + 1], mem[i]          # rows of a bucket no longer ordered by molecule
    bad = BatchLayout.from_reference_arrays(lay.deg_slice, mem, lay.deg_adjacency_lists(), 40)
    assert bad.n_groups == 0 and int(bad.groups[0, 3]) == 0
    adj = [a.copy() for a in lay.deg_adjacency_lists()]
    first = int(lay.deg_slice[1, 0])
    other = int(np.flatnonzero(lay.membership != lay.membership[first])[0])
    adj[1][0, 0] = other                             # an edge into another molecule
    bad = BatchLayout.from_reference_arrays(lay.deg_slice, lay.membership, adj, 40)
    assert bad.n_groups == 0


def test_empty_batch_has_no_groups():
    from deepchem_b200.synthetic import PackedMols
    from deepchem_b200.mol_graphs import BatchLayout
    pm = PackedMols(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros((0, 75), np.float32))
    lay = BatchLayout.build(pm, n_segments=2)
    assert lay.n_groups == 0 and lay.group_max_rows == 0


def _all_fields(lay):
    names = ("deg_slice", "membership", "perm", "row_ptr", "col_idx", "t_row_ptr", "t_src", "t_slot", "mol_ptr",
             "mol_atoms", "tiles", "groups")
    return {f: np.array(getattr(lay, f)) for f in names}


@pytest.mark.parametrize("shape,n,seed,segs", [("zinc", 2048, 0, 2048), ("stress", 1000, 1, 1003), ("delaney", 64, 2, 64),
                                               ("qm9", 500, 3, 500), ("tox21", 777, 4, 800), ("zinc", 3, 5, 3),
                                               ("stress", 50, 6, 50)])
def test_single_pass_builder_equals_general_builder(shape, n, seed, segs, monkeypatch):
    """dcgc_layout_build's molecule-local single pass (symmetric adjacencies) writes the same integers into every
    section of the slab as the general multi-pass builder, which is pinned to the reference's arrays above."""
    from deepchem_b200.mol_graphs import BatchLayout
    from deepchem_b200.synthetic import make_molecules
    pm = make_molecules(n, seed=seed, shape=shape)
    monkeypatch.delenv("DCGC_LAYOUT_GENERAL", raising=False)
    fast = _all_fields(BatchLayout.build(pm, n_segments=segs))
    monkeypatch.setenv("DCGC_LAYOUT_GENERAL", "1")
    general = _all_fields(BatchLayout.build(pm, n_segments=segs))
    for f in fast:
        assert np.array_equal(fast[f], general[f]), f


def test_single_pass_builder_hands_asymmetric_batches_to_the_general_path(monkeypatch):
    """An adjacency in which an atom is listed more (or less) often than it lists others is not a molecular graph;
    the single pass detects it and the general builder answers, so both settings agree and `symmetric` is False."""
    from deepchem_b200.mol_graphs import BatchLayout
    from deepchem_b200.synthetic import PackedMols, make_molecules
    pm = make_molecules(40, seed=9, shape="zinc")
    # molecule 7: drop the last entry of its last atom that has >= 2 neighbours (its partner still lists it)
    a0, a1 = int(pm.atom_ptr[7]), int(pm.atom_ptr[8])
    deg = np.diff(pm.adj_ptr)
    victim = max(a for a in range(a0, a1) if deg[a] >= 2)
    cut = int(pm.adj_ptr[victim + 1]) - 1
    adj_idx = np.delete(pm.adj_idx, cut)
    adj_ptr = pm.adj_ptr.copy()
    adj_ptr[victim + 1:] -= 1
    bad = PackedMols(pm.atom_ptr, adj_ptr, adj_idx, pm.features)
    monkeypatch.delenv("DCGC_LAYOUT_GENERAL", raising=False)
    a = BatchLayout.build(bad)
    monkeypatch.setenv("DCGC_LAYOUT_GENERAL", "1")
    b = _all_fields(BatchLayout.build(bad))
    assert not a.symmetric
    fa = _all_fields(a)
    for f in fa:
        assert np.array_equal(fa[f], b[f]), f


def test_device_topology_views_are_lazy_and_equal_the_host_slab():
    """DeviceTopology (here on the CPU device: the class is device-agnostic) makes the typed views of the slab sections
    on first use; ptr() gives their addresses without a view; the C struct is filled from addresses alone; the
    model-input list materialises membership / adjacency entries only when they are asked for."""
    import torch
    from deepchem_b200.engine import topology_struct
    pm = make_molecules(50, seed=1, shape="tox21")
    lay = MG.BatchLayout.build(pm)
    topo = lay.to_device("cpu")
    assert not any(name in topo.__dict__ for name, *_ in MG._SLAB_FIELDS)          # nothing made yet
    st = topology_struct(topo)
    assert not any(name in topo.__dict__ for name, *_ in MG._SLAB_FIELDS)          # the struct needs addresses only
    for name, *_ in MG._SLAB_FIELDS:
        view, host = getattr(topo, name), getattr(lay, name)
        assert name in topo.__dict__ and getattr(topo, name) is view                 # cached after the first access
        assert np.array_equal(view.numpy().reshape(-1), np.asarray(host).reshape(-1)), name
        assert view.numel() == 0 or view.data_ptr() == topo.ptr(name), name
    assert topo.deg_slice.shape == (11, 2) and topo.deg_slice.dtype == torch.int64
    assert topo.tiles.shape[1] == 4 and topo.groups.shape[1] == MG.GROUP_STRIDE
    assert st.membership == topo.membership.data_ptr() and st.col_idx == topo.ptr("col_idx")
    assert st.groups == topo.groups[1:].data_ptr()                                    # past the header row
    with pytest.raises(AttributeError):
        topo.no_such_section
    x = torch.zeros(lay.n_atoms, 76)
    mi = topo.model_inputs(x, n_samples=50)
    assert len(mi) == 14 and mi[0] is x and int(mi[3]) == 50 and mi[1]._dcgc_topology is topo
    assert list.__getitem__(mi, 2) is None and list.__getitem__(mi, 13) is None     # engine path: never materialised
    assert mi[2].shape[0] == lay.n_atoms and list.__getitem__(mi, 13) is not None   # first other access fills all
    for make in (lambda m: list(m), lambda m: m[4:], lambda m: m + [], lambda m: [t for t in m][4:], lambda m: m[-10:]):
        got = make(topo.model_inputs(x))
        adj = got[-10:]
        assert all(a is not None for a in got) and [a.shape[1] for a in adj] == list(range(1, 11))
        assert sum(a.shape[0] for a in adj) + lay.deg_count[0] == lay.n_atoms


def test_reference_mol_graphs_known_answers():
    """The hand-written expectations of the reference's own tests (feat/tests/test_mol_graphs.py:21-141) against the
    ConvMol class here and the C++ batch builder: deg_slice of a 4-ring, degree-sorted features, renumbered adjacency,
    the three-molecule agglomeration golden, the null molecule."""
    from deepchem_b200.mol_graphs import ConvMol
    ring = ConvMol(np.array([[20, 21, 22, 23], [24, 25, 26, 27], [28, 29, 30, 31], [32, 33, 34, 35]]),
                   [[1, 2], [0, 3], [0, 3], [1, 2]])
    assert np.array_equal(ring.get_deg_slice(), np.array([[0, 0], [0, 0], [0, 4]] + [[0, 0]] * 8))      # :21-42
    f5 = np.array([[40, 41, 42, 43], [44, 45, 46, 47], [48, 49, 50, 51], [52, 53, 54, 55], [56, 57, 58, 59]])
    five = ConvMol(f5, [[1, 2], [0, 3], [0, 3], [1, 2, 4], [3]])
    assert np.array_equal(five.get_atom_features(), f5[[4, 0, 1, 2, 3]])                                 # :44-60
    assert five.get_adjacency_list() == [[4], [2, 3], [1, 4], [1, 4], [2, 3, 0]]                          # :62-75
    chain = ConvMol(np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]), [[1], [0, 2], [1]])
    for batch in (ConvMol.agglomerate_mols([chain, ring, five]),                                          # :77-126
                  MG.BatchLayout.build(MG.pack_convmols([chain, ring, five])).multi_conv_mol(
                      np.concatenate([m.atom_features for m in (chain, ring, five)]).astype(np.float32))):
        assert batch.get_num_atoms() == 12 and batch.get_num_molecules() == 3
        x = batch.get_atom_features()
        assert np.array_equal(x[0], [1, 2, 3, 4]) and np.array_equal(x[2], [56, 57, 58, 59])
        assert np.array_equal(x[11], [52, 53, 54, 55]) and np.array_equal(x[4], [20, 21, 22, 23])
        adj = batch.get_deg_adjacency_lists()
        assert adj[0].shape == (0, 0)
        assert np.array_equal(adj[1], [[3], [3], [11]])
        assert np.array_equal(adj[2], [[0, 1], [5, 6], [4, 7], [4, 7], [5, 6], [9, 10], [8, 11], [8, 11]])
        assert np.array_equal(adj[3], [[9, 10, 2]])
        assert adj[4].shape == (0, 4) and adj[5].shape == (0, 5)
    null = ConvMol.get_null_mol(4)                                                                       # :128-141
    adj = null.get_deg_adjacency_lists()
    assert np.array_equal(adj[10], [[10] * 10]) and np.array_equal(adj[1], [[1]])
    assert np.array_equal(null.get_deg_slice(), [[d, 1] for d in range(11)])

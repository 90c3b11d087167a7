// Host layout builder: packed molecule shard -> degree-major batch slab.
//
// Replaces ConvMol._deg_sort + ConvMol.agglomerate_mols
// (deepchem/feat/mol_graphs.py:113-185, 256-349).  The reference sorts every molecule's atoms
// stably by degree, then sorts the concatenation stably by degree again; the composite is one
// stable counting sort of all atoms keyed by degree, i.e. the batch order is
// (degree, molecule, position inside the molecule).  Everything here is integer work and must
// match the reference bit for bit (tests/test_layout.py).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.h"

static thread_local char g_err[512] = "";

void dcgc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* dcgc_last_error(void) { return g_err; }
extern "C" int dcgc_version(void) { return 100; }

// row budget of a molecule group (dcgc.h); DCGC_GROUP_ROWS overrides it (read at every plan: tuning runs and the
// wide-row tests change it inside one process)
static int group_rows_setting() {
  const char* e = getenv("DCGC_GROUP_ROWS");
  const int x = e ? atoi(e) : 0;
  return x >= 8 && x <= 4096 ? x : DCGC_GROUP_ROWS_DEFAULT;
}

static void plan_offsets(dcgc_layout_info* info) {
  const int64_t N = info->n_atoms, E = info->n_edges, S = info->n_segments;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = dcgc_align_up(off + bytes, 256);
    return o;
  };
  info->off_deg_slice = take(DCGC_N_DEG * 2 * 8);
  info->off_membership = take(N * 4);
  info->off_perm = take(N * 4);
  info->off_row_ptr = take((N + 1) * 4);
  info->off_col_idx = take(E * 4);
  info->off_t_row_ptr = take((N + 1) * 4);
  info->off_t_src = take(E * 4);
  info->off_t_slot = take(E * 4);
  info->off_mol_ptr = take((S + 1) * 4);
  info->off_mol_atoms = take(N * 4);
  info->off_tiles = take(info->n_tiles * 16);
  info->off_groups = take((int64_t)DCGC_GROUP_STRIDE * (info->n_groups_alloc + 2) * 4);
  info->slab_bytes = off;
}

static int finish_plan(dcgc_layout_info* info, int64_t n_segments, int32_t tile_rows) {
  int64_t N = 0, E = 0, T = 0;
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    N += info->deg_count[d];
    E += (int64_t)d * info->deg_count[d];
    T += (info->deg_count[d] + tile_rows - 1) / tile_rows;
  }
  if (N >= (int64_t)1 << 31 || E >= (int64_t)1 << 31) {
    dcgc_set_error("batch too large for int32 indices (N=%lld, E=%lld)", (long long)N, (long long)E);
    return DCGC_ERR_INVALID;
  }
  info->n_atoms = N;
  info->n_edges = E;
  info->n_tiles = T;
  info->tile_rows = tile_rows;
  info->reserved = 0;
  info->n_segments = n_segments;
  info->group_rows = group_rows_setting();
  info->n_groups_alloc = (int32_t)(2 * N / info->group_rows + 2);
  plan_offsets(info);
  return DCGC_OK;
}

extern "C" int dcgc_layout_plan(int64_t n_mols, const int32_t* atom_ptr, const int32_t* adj_ptr,
                                int64_t n_segments, int32_t tile_rows, dcgc_layout_info* info) {
  DCGC_CHECK_ARG(n_mols >= 0 && atom_ptr && adj_ptr && info, "dcgc_layout_plan: null argument");
  DCGC_CHECK_ARG(tile_rows > 0, "dcgc_layout_plan: tile_rows must be positive");
  DCGC_CHECK_ARG(n_segments >= n_mols, "dcgc_layout_plan: n_segments (%lld) < n_mols (%lld)",
                 (long long)n_segments, (long long)n_mols);
  memset(info, 0, sizeof(*info));
  info->n_mols = n_mols;
  const int64_t N = atom_ptr[n_mols];
  DCGC_CHECK_ARG(atom_ptr[0] == 0 && N >= 0, "dcgc_layout_plan: atom_ptr must start at 0");
  for (int64_t a = 0; a < N; ++a) {
    const int64_t d = (int64_t)adj_ptr[a + 1] - adj_ptr[a];
    if (d < 0 || d > DCGC_MAX_DEG) {
      dcgc_set_error("atom %lld has degree %lld; the layout supports degrees 0..%d", (long long)a,
                     (long long)d, DCGC_MAX_DEG);
      return DCGC_ERR_DEGREE;
    }
    info->deg_count[d]++;
  }
  return finish_plan(info, n_segments, tile_rows);
}

// Shared tail: everything derivable from (deg_count, membership, col_idx).
static void build_derived(const dcgc_layout_info* info, char* slab, bool check_neighbours) {
  const int64_t N = info->n_atoms, E = info->n_edges, S = info->n_segments;
  int64_t* deg_slice = (int64_t*)(slab + info->off_deg_slice);
  const int32_t* membership = (const int32_t*)(slab + info->off_membership);
  int32_t* row_ptr = (int32_t*)(slab + info->off_row_ptr);
  const int32_t* col_idx = (const int32_t*)(slab + info->off_col_idx);
  int32_t* t_row_ptr = (int32_t*)(slab + info->off_t_row_ptr);
  int32_t* t_src = (int32_t*)(slab + info->off_t_src);
  int32_t* t_slot = (int32_t*)(slab + info->off_t_slot);
  int32_t* mol_ptr = (int32_t*)(slab + info->off_mol_ptr);
  int32_t* mol_atoms = (int32_t*)(slab + info->off_mol_atoms);
  int32_t* tiles = (int32_t*)(slab + info->off_tiles);

  // deg_slice: running starts, not zeroed for empty buckets (mol_graphs.py:295-305)
  int64_t start = 0, t = 0, e = 0;
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    const int64_t cnt = info->deg_count[d];
    deg_slice[2 * d] = start;
    deg_slice[2 * d + 1] = cnt;
    for (int64_t r = 0; r < cnt; ++r) {
      row_ptr[start + r] = (int32_t)e;
      e += d;
    }
    for (int64_t r = 0; r < cnt; r += info->tile_rows) {
      tiles[4 * t + 0] = (int32_t)(start + r);
      tiles[4 * t + 1] = (int32_t)std::min<int64_t>(info->tile_rows, cnt - r);
      tiles[4 * t + 2] = d;
      tiles[4 * t + 3] = 0;
      ++t;
    }
    start += cnt;
  }
  row_ptr[N] = (int32_t)e;

  // transposed CSR by counting sort on the column id; entries of one column end up ordered by
  // (row, slot) because rows and slots are visited in ascending order.
  memset(t_row_ptr, 0, (size_t)(N + 1) * 4);
  for (int64_t k = 0; k < E; ++k) t_row_ptr[col_idx[k] + 1]++;
  for (int64_t j = 0; j < N; ++j) t_row_ptr[j + 1] += t_row_ptr[j];
  {
    std::vector<int32_t> cur(t_row_ptr, t_row_ptr + N);
    for (int64_t i = 0; i < N; ++i) {
      for (int32_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
        const int32_t p = cur[col_idx[k]]++;
        t_src[p] = (int32_t)i;
        t_slot[p] = k - row_ptr[i];
      }
    }
  }
  // molecule -> rows CSR (membership is not globally sorted)
  memset(mol_ptr, 0, (size_t)(S + 1) * 4);
  for (int64_t i = 0; i < N; ++i) mol_ptr[membership[i] + 1]++;
  for (int64_t g = 0; g < S; ++g) mol_ptr[g + 1] += mol_ptr[g];
  {
    std::vector<int32_t> cur(mol_ptr, mol_ptr + S);
    for (int64_t i = 0; i < N; ++i) mol_atoms[cur[membership[i]]++] = (int32_t)i;
  }

  // molecule groups (dcgc.h): consecutive molecules are packed greedily into groups of at most R rows (a
  // molecule larger than R is a group of its own).  Walking the molecules in order, every row of molecule m
  // must be the next unconsumed row of its degree bucket (rows of a bucket are ordered by molecule); the table
  // row of a group is the snapshot of the 11 bucket cursors when it starts.  Two consecutive groups always hold
  // more than R rows together, hence at most 2N/R + 1 groups.
  int32_t* gt = (int32_t*)(slab + info->off_groups);
  const int R = info->group_rows;
  memset(gt, 0, (size_t)DCGC_GROUP_STRIDE * (info->n_groups_alloc + 2) * 4);
  int32_t* tab = gt + DCGC_GROUP_STRIDE;
  int64_t cursor[DCGC_N_DEG + 1], bstart[DCGC_N_DEG + 1];
  {
    int64_t s0 = 0;
    for (int d = 0; d < DCGC_N_DEG; ++d) { bstart[d] = cursor[d] = s0; s0 += info->deg_count[d]; }
    bstart[DCGC_N_DEG] = s0;
  }
  bool valid = true;
  int64_t G = 0, cur_rows = 0;   // groups opened so far, rows of the open group
  auto snapshot = [&](int64_t mol) {
    int32_t* row = tab + G * DCGC_GROUP_STRIDE;
    for (int d = 0; d < DCGC_N_DEG; ++d) row[d] = (int32_t)cursor[d];
    row[DCGC_N_DEG] = (int32_t)mol;
  };
  for (int64_t m = 0; m < S && valid; ++m) {
    const int64_t n_m = mol_ptr[m + 1] - mol_ptr[m];
    if (n_m == 0) continue;
    if (G == 0 || cur_rows + n_m > R) {
      if (G >= info->n_groups_alloc) { valid = false; break; }   // cannot happen (bound above)
      snapshot(m);
      ++G;
      cur_rows = 0;
    }
    cur_rows += n_m;
    int d = 0;
    for (int32_t k = mol_ptr[m]; k < mol_ptr[m + 1]; ++k) {
      const int64_t r = mol_atoms[k];
      while (r >= bstart[d + 1]) ++d;
      if (r != cursor[d]) { valid = false; break; }
      ++cursor[d];
    }
  }
  snapshot(S);   // closing row: every cursor at its bucket end
  if (valid && check_neighbours) {
    for (int64_t i = 0; i < N && valid; ++i)
      for (int32_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
        if (membership[col_idx[k]] != membership[i]) { valid = false; break; }
  }
  int64_t max_rows = 0, max_entries = 0;
  for (int64_t g = 0; g < G; ++g) {
    int64_t rows = 0, entries = 0;
    for (int d = 0; d < DCGC_N_DEG; ++d) {
      const int64_t n = tab[(g + 1) * DCGC_GROUP_STRIDE + d] - tab[g * DCGC_GROUP_STRIDE + d];
      rows += n;
      entries += n * d;
    }
    max_rows = std::max(max_rows, rows);
    max_entries = std::max(max_entries, entries);
  }
  gt[0] = valid ? (int32_t)G : 0;
  gt[1] = (int32_t)max_rows;
  gt[2] = (int32_t)max_entries;
  gt[3] = valid ? 1 : 0;
  gt[4] = R;
}

// One pass over the molecules that writes EVERY section of the slab (the general path below makes five passes over
// the batch with batch-sized scatter targets: 2.9 ms for 4096 molecules on one core; this one 1.2 ms).  It relies
// on two facts: (1) the rows of one molecule, visited by (degree, position), are visited in ascending batch row
// order, so the molecule's row list (mol_atoms), its transposed entries ordered by (source row, slot) and the group
// cursors all fall out of a molecule-local walk with molecule-sized scratch; (2) for an adjacency in which every
// atom is listed by exactly as many atoms as it lists itself (every molecular graph), the transposed CSR has the
// row offsets of the forward CSR, which are arithmetic in the degree buckets.  Returns 1 when the batch has an
// atom for which (2) fails (the caller then runs the general path), 0 on success, < 0 on error.
static int build_single_pass(int64_t n_mols, const int32_t* atom_ptr, const int32_t* adj_ptr, const int32_t* adj_idx,
                             const dcgc_layout_info* info, char* slab) {
  const int64_t N = info->n_atoms, S = info->n_segments;
  int64_t* deg_slice = (int64_t*)(slab + info->off_deg_slice);
  int32_t* membership = (int32_t*)(slab + info->off_membership);
  int32_t* perm = (int32_t*)(slab + info->off_perm);
  int32_t* row_ptr = (int32_t*)(slab + info->off_row_ptr);
  int32_t* col_idx = (int32_t*)(slab + info->off_col_idx);
  int32_t* t_row_ptr = (int32_t*)(slab + info->off_t_row_ptr);
  int32_t* t_src = (int32_t*)(slab + info->off_t_src);
  int32_t* t_slot = (int32_t*)(slab + info->off_t_slot);
  int32_t* mol_ptr = (int32_t*)(slab + info->off_mol_ptr);
  int32_t* mol_atoms = (int32_t*)(slab + info->off_mol_atoms);
  int32_t* tiles = (int32_t*)(slab + info->off_tiles);
  int32_t* gt = (int32_t*)(slab + info->off_groups);
  const int R = info->group_rows;

  int64_t bstart[DCGC_N_DEG + 1], estart[DCGC_N_DEG + 1], cursor[DCGC_N_DEG];
  {
    int64_t s0 = 0, e0 = 0, t = 0;
    for (int d = 0; d < DCGC_N_DEG; ++d) {
      const int64_t cnt = info->deg_count[d];
      bstart[d] = cursor[d] = s0;
      estart[d] = e0;
      deg_slice[2 * d] = s0;          // running starts, not zeroed for empty buckets (mol_graphs.py:295-305)
      deg_slice[2 * d + 1] = cnt;
      int32_t e = (int32_t)e0;
      for (int64_t r = 0; r < cnt; ++r, e += d) row_ptr[s0 + r] = e;
      for (int64_t r = 0; r < cnt; r += info->tile_rows) {
        tiles[4 * t + 0] = (int32_t)(s0 + r);
        tiles[4 * t + 1] = (int32_t)std::min<int64_t>(info->tile_rows, cnt - r);
        tiles[4 * t + 2] = d;
        tiles[4 * t + 3] = 0;
        ++t;
      }
      s0 += cnt;
      e0 += (int64_t)d * cnt;
    }
    bstart[DCGC_N_DEG] = s0;
    estart[DCGC_N_DEG] = e0;
    row_ptr[N] = (int32_t)e0;
  }
  memset(gt, 0, (size_t)DCGC_GROUP_STRIDE * (info->n_groups_alloc + 2) * 4);
  int32_t* tab = gt + DCGC_GROUP_STRIDE;
  int64_t G = 0, cur_rows = 0;
  auto snapshot = [&](int64_t mol) {
    int32_t* row = tab + G * DCGC_GROUP_STRIDE;
    for (int d = 0; d < DCGC_N_DEG; ++d) row[d] = (int32_t)cursor[d];
    row[DCGC_N_DEG] = (int32_t)mol;
  };
  // molecule-local scratch, grown to the largest molecule
  std::vector<int32_t> new_id, order, fill, tbase, degs;   // (int32 degrees: a uint8 array would alias every store)
  for (int64_t m = 0; m < n_mols; ++m) {
    const int64_t base = atom_ptr[m];
    const int n = (int)(atom_ptr[m + 1] - base);
    mol_ptr[m] = (int32_t)base;
    if (n <= 0) {
      if (n < 0) { dcgc_set_error("dcgc_layout_build: atom_ptr decreases at molecule %lld", (long long)m); return DCGC_ERR_INVALID; }
      continue;
    }
    if (G == 0 || cur_rows + n > R) {
      if (G >= info->n_groups_alloc) { dcgc_set_error("dcgc_layout_build: group table overflow"); return DCGC_ERR_INVALID; }
      snapshot(m);
      ++G;
      cur_rows = 0;
    }
    cur_rows += n;
    if ((int)new_id.size() < n) { new_id.resize(n); order.resize(n); fill.resize(n); tbase.resize(n); degs.resize(n); }
    // rows of this molecule: stable counting sort by degree, continuing the batch cursors
    int cnt[DCGC_N_DEG + 1] = {0};
    const int32_t* ap = adj_ptr + base;
    for (int a = 0; a < n; ++a) {
      const int d = ap[a + 1] - ap[a];
      if (d < 0 || d > DCGC_MAX_DEG) {
        dcgc_set_error("atom %lld has degree %d; the layout supports degrees 0..%d", (long long)(base + a), d, DCGC_MAX_DEG);
        return DCGC_ERR_DEGREE;
      }
      degs[a] = d;
      ++cnt[d + 1];
    }
    for (int d = 0; d < DCGC_N_DEG; ++d) cnt[d + 1] += cnt[d];        // cnt[d] = first position of degree d
    int32_t* ma = mol_atoms + base;
    for (int a = 0; a < n; ++a) {
      const int d = degs[a];
      const int64_t r = cursor[d]++;
      if (r >= bstart[d + 1]) { dcgc_set_error("dcgc_layout_build: degree histogram changed since dcgc_layout_plan"); return DCGC_ERR_INVALID; }
      new_id[a] = (int32_t)r;
      perm[r] = (int32_t)(base + a);
      membership[r] = (int32_t)m;
      const int pos = cnt[d]++;
      order[pos] = a;
      ma[pos] = (int32_t)r;                                           // ascending: buckets ascend, cursors ascend
      fill[a] = 0;
      tbase[a] = (int32_t)(estart[d] + (r - bstart[d]) * d);          // first entry of row r, forward == transposed
    }
    // neighbour lists renumbered to batch rows in list order (mol_graphs.py:139-141, 327-336), and their transpose:
    // sources are visited in ascending row order and slots in order, so every transposed list ends up ordered by
    // (source row, slot) — the order of the counting sort of the general path
    for (int pos = 0; pos < n; ++pos) {
      const int a = order[pos];
      const int d = degs[a];
      const int32_t r = new_id[a];
      int32_t* dst = col_idx + tbase[a];
      const int32_t* src = adj_idx + ap[a];
      for (int k = 0; k < d; ++k) {
        const int64_t nb = src[k];
        if (nb < 0 || nb >= n) {
          dcgc_set_error("molecule %lld atom %lld: neighbour index %lld outside [0,%lld)", (long long)m, (long long)a,
                         (long long)nb, (long long)n);
          return DCGC_ERR_INDEX;
        }
        dst[k] = new_id[nb];
        const int f = fill[nb]++;
        if (f >= degs[nb]) return 1;                                  // listed more often than it lists: general path
        const int32_t p = tbase[nb] + f;
        t_src[p] = r;
        t_slot[p] = k;
      }
    }
    for (int a = 0; a < n; ++a)
      if (fill[a] != degs[a]) return 1;
  }
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    if (cursor[d] != bstart[d + 1]) {
      dcgc_set_error("dcgc_layout_build: degree histogram changed since dcgc_layout_plan");
      return DCGC_ERR_INVALID;
    }
  }
  for (int64_t g = n_mols; g <= S; ++g) mol_ptr[g] = (int32_t)N;
  memcpy(t_row_ptr, row_ptr, (size_t)(N + 1) * 4);
  snapshot(S);   // closing row: every cursor at its bucket end
  int64_t max_rows = 0, max_entries = 0;
  for (int64_t g = 0; g < G; ++g) {
    int64_t rows = 0, entries = 0;
    for (int d = 0; d < DCGC_N_DEG; ++d) {
      const int64_t c = tab[(g + 1) * DCGC_GROUP_STRIDE + d] - tab[g * DCGC_GROUP_STRIDE + d];
      rows += c;
      entries += c * d;
    }
    max_rows = std::max(max_rows, rows);
    max_entries = std::max(max_entries, entries);
  }
  gt[0] = (int32_t)G;
  gt[1] = (int32_t)max_rows;
  gt[2] = (int32_t)max_entries;
  gt[3] = 1;
  gt[4] = R;
  return 0;
}

extern "C" int dcgc_layout_build(int64_t n_mols, const int32_t* atom_ptr, const int32_t* adj_ptr,
                                 const int32_t* adj_idx, const dcgc_layout_info* info, void* slab_v) {
  DCGC_CHECK_ARG(atom_ptr && adj_ptr && info && slab_v, "dcgc_layout_build: null argument");
  DCGC_CHECK_ARG(n_mols == info->n_mols && atom_ptr[n_mols] == info->n_atoms,
                 "dcgc_layout_build: info does not match the inputs");
  DCGC_CHECK_ARG(adj_idx || info->n_edges == 0, "dcgc_layout_build: null adj_idx");
  char* slab = (char*)slab_v;
  {
    const char* e = getenv("DCGC_LAYOUT_GENERAL");     // =1: always the general path (tests compare the two)
    const bool general_only = e && e[0] == '1';
    if (!general_only) {
      const int rc = build_single_pass(n_mols, atom_ptr, adj_ptr, adj_idx, info, slab);
      if (rc <= 0) return rc;     // done, or an input error; 1: not a symmetric adjacency -> general path below
    }
  }
  const int64_t N = info->n_atoms;
  int32_t* membership = (int32_t*)(slab + info->off_membership);
  int32_t* perm = (int32_t*)(slab + info->off_perm);
  int32_t* col_idx = (int32_t*)(slab + info->off_col_idx);

  // stable counting sort by degree
  int64_t bucket_start[DCGC_N_DEG], edge_start[DCGC_N_DEG], cursor[DCGC_N_DEG];
  int64_t s = 0, es = 0;
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    bucket_start[d] = cursor[d] = s;
    edge_start[d] = es;
    s += info->deg_count[d];
    es += (int64_t)d * info->deg_count[d];
  }
  std::vector<int32_t> new_of_old((size_t)N);
  for (int64_t m = 0; m < n_mols; ++m) {
    for (int64_t a = atom_ptr[m]; a < atom_ptr[m + 1]; ++a) {
      const int d = adj_ptr[a + 1] - adj_ptr[a];
      if (d < 0 || d > DCGC_MAX_DEG) {
        dcgc_set_error("atom %lld has degree %d; the layout supports degrees 0..%d", (long long)a, d,
                       DCGC_MAX_DEG);
        return DCGC_ERR_DEGREE;
      }
      const int64_t r = cursor[d]++;
      new_of_old[a] = (int32_t)r;
      perm[r] = (int32_t)a;
      membership[r] = (int32_t)m;
    }
  }
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    if (cursor[d] != bucket_start[d] + info->deg_count[d]) {
      dcgc_set_error("dcgc_layout_build: degree histogram changed since dcgc_layout_plan");
      return DCGC_ERR_INVALID;
    }
  }
  // neighbour lists: renumber to batch rows, keep list order (mol_graphs.py:139-141, 327-336)
  for (int64_t m = 0; m < n_mols; ++m) {
    const int64_t base = atom_ptr[m], n_local = atom_ptr[m + 1] - base;
    for (int64_t a = base; a < base + n_local; ++a) {
      const int d = adj_ptr[a + 1] - adj_ptr[a];
      const int64_t r = new_of_old[a];
      int32_t* dst = col_idx + edge_start[d] + (r - bucket_start[d]) * d;
      const int32_t* src = adj_idx + adj_ptr[a];
      for (int k = 0; k < d; ++k) {
        const int64_t nb = src[k];
        if (nb < 0 || nb >= n_local) {
          dcgc_set_error("molecule %lld atom %lld: neighbour index %lld outside [0,%lld)", (long long)m,
                         (long long)(a - base), (long long)nb, (long long)n_local);
          return DCGC_ERR_INDEX;
        }
        dst[k] = new_of_old[base + nb];
      }
    }
  }
  build_derived(info, slab, false);
  return DCGC_OK;
}

extern "C" int dcgc_layout_plan_from_deg(const int64_t* deg_slice, int64_t n_segments,
                                         int32_t tile_rows, dcgc_layout_info* info) {
  DCGC_CHECK_ARG(deg_slice && info, "dcgc_layout_plan_from_deg: null argument");
  DCGC_CHECK_ARG(tile_rows > 0 && n_segments >= 0, "dcgc_layout_plan_from_deg: bad sizes");
  memset(info, 0, sizeof(*info));
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    DCGC_CHECK_ARG(deg_slice[2 * d + 1] >= 0, "dcgc_layout_plan_from_deg: negative bucket size");
    info->deg_count[d] = deg_slice[2 * d + 1];
  }
  info->n_mols = n_segments;
  return finish_plan(info, n_segments, tile_rows);
}

extern "C" int dcgc_layout_build_from_deg(const int64_t* deg_slice, const int32_t* membership,
                                          const int32_t* col_idx, const dcgc_layout_info* info,
                                          void* slab_v) {
  DCGC_CHECK_ARG(deg_slice && info && slab_v, "dcgc_layout_build_from_deg: null argument");
  DCGC_CHECK_ARG((membership || info->n_atoms == 0) && (col_idx || info->n_edges == 0),
                 "dcgc_layout_build_from_deg: null index array");
  char* slab = (char*)slab_v;
  const int64_t N = info->n_atoms, E = info->n_edges, S = info->n_segments;
  for (int64_t i = 0; i < N; ++i) {
    if (membership[i] < 0 || membership[i] >= S) {
      dcgc_set_error("membership[%lld] = %d outside [0,%lld)", (long long)i, membership[i], (long long)S);
      return DCGC_ERR_INDEX;
    }
  }
  for (int64_t k = 0; k < E; ++k) {
    if (col_idx[k] < 0 || col_idx[k] >= N) {
      dcgc_set_error("adjacency entry %lld = %d outside [0,%lld)", (long long)k, col_idx[k], (long long)N);
      return DCGC_ERR_INDEX;
    }
  }
  memcpy(slab + info->off_membership, membership, (size_t)N * 4);
  memcpy(slab + info->off_col_idx, col_idx, (size_t)E * 4);
  int32_t* perm = (int32_t*)(slab + info->off_perm);
  for (int64_t i = 0; i < N; ++i) perm[i] = (int32_t)i;
  build_derived(info, slab, true);
  return DCGC_OK;
}

extern "C" int dcgc_layout_permute_features_host(const float* src, int64_t ld_src, const int32_t* perm,
                                                 int64_t n_atoms, int32_t n_feat, float* dst,
                                                 int64_t ld_dst, int32_t n_threads) {
  DCGC_CHECK_ARG((src && perm && dst) || n_atoms == 0, "dcgc_layout_permute_features_host: null argument");
  DCGC_CHECK_ARG(n_feat >= 0 && ld_src >= n_feat && ld_dst >= n_feat,
                 "dcgc_layout_permute_features_host: leading dimension smaller than n_feat");
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      float* d = dst + i * ld_dst;
      memcpy(d, src + (int64_t)perm[i] * ld_src, (size_t)n_feat * 4);
      for (int64_t c = n_feat; c < ld_dst; ++c) d[c] = 0.f;
    }
  };
  if (n_threads <= 1 || n_atoms < 4096) {
    work(0, n_atoms);
    return DCGC_OK;
  }
  std::vector<std::thread> pool;
  const int64_t chunk = (n_atoms + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t lo = t * chunk, hi = std::min(n_atoms, lo + chunk);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return DCGC_OK;
}

// ------------------------------------------------------------------------------------------
// Gather of molecules out of a packed shard (PackedMols.take): what a SHUFFLED epoch does for every batch
// (DiskDataset.iterbatches with deterministic=False, deepchem/data/datasets.py:1598-1730, permutes the sample
// indices of a shard).  Molecules are contiguous blocks of the shard, so the gather is one memcpy per molecule
// and array; adjacency entries are molecule-local and are copied unchanged.
// ------------------------------------------------------------------------------------------
extern "C" int dcgc_packed_take_plan(int64_t n_take, const int64_t* idx, int64_t n_src_mols, const int32_t* atom_ptr,
                                     const int32_t* adj_ptr, int64_t* n_atoms_out, int64_t* n_entries_out) {
  DCGC_CHECK_ARG(n_take >= 0 && n_src_mols >= 0 && n_atoms_out && n_entries_out, "dcgc_packed_take_plan: bad arguments");
  DCGC_CHECK_ARG((idx && atom_ptr && adj_ptr) || n_take == 0, "dcgc_packed_take_plan: null argument");
  int64_t na = 0, ne = 0;
  for (int64_t i = 0; i < n_take; ++i) {
    const int64_t m = idx[i];
    if (m < 0 || m >= n_src_mols) {
      dcgc_set_error("dcgc_packed_take_plan: molecule index %lld out of range [0, %lld)", (long long)m, (long long)n_src_mols);
      return DCGC_ERR_INDEX;
    }
    na += atom_ptr[m + 1] - atom_ptr[m];
    ne += adj_ptr[atom_ptr[m + 1]] - adj_ptr[atom_ptr[m]];
  }
  DCGC_CHECK_ARG(na < ((int64_t)1 << 31) && ne < ((int64_t)1 << 31), "dcgc_packed_take_plan: batch too large for int32 offsets");
  *n_atoms_out = na;
  *n_entries_out = ne;
  return DCGC_OK;
}

// features / features2 (either may be null): row-major per-atom matrices with row_bytes / row_bytes2 bytes per atom
// (the fp32 feature matrix and its exact int8 copy); outputs sized by dcgc_packed_take_plan.
extern "C" int dcgc_packed_take(int64_t n_take, const int64_t* idx, int64_t n_src_mols, const int32_t* atom_ptr,
                                const int32_t* adj_ptr, const int32_t* adj_idx, const void* features, int64_t row_bytes,
                                const void* features2, int64_t row_bytes2, int32_t* out_atom_ptr, int32_t* out_adj_ptr,
                                int32_t* out_adj_idx, void* out_features, void* out_features2, int32_t n_threads) {
  DCGC_CHECK_ARG(n_take >= 0 && out_atom_ptr && out_adj_ptr, "dcgc_packed_take: bad arguments");
  DCGC_CHECK_ARG((idx && atom_ptr && adj_ptr) || n_take == 0, "dcgc_packed_take: null argument");
  DCGC_CHECK_ARG((features == nullptr) == (out_features == nullptr) && (features2 == nullptr) == (out_features2 == nullptr),
                 "dcgc_packed_take: feature input / output mismatch");
  std::vector<int64_t> a0((size_t)n_take + 1), e0((size_t)n_take + 1);
  a0[0] = e0[0] = 0;
  for (int64_t i = 0; i < n_take; ++i) {
    const int64_t m = idx[i];
    if (m < 0 || m >= n_src_mols) {
      dcgc_set_error("dcgc_packed_take: molecule index %lld out of range [0, %lld)", (long long)m, (long long)n_src_mols);
      return DCGC_ERR_INDEX;
    }
    a0[i + 1] = a0[i] + (atom_ptr[m + 1] - atom_ptr[m]);
    e0[i + 1] = e0[i] + (adj_ptr[atom_ptr[m + 1]] - adj_ptr[atom_ptr[m]]);
    out_atom_ptr[i] = (int32_t)a0[i];
  }
  out_atom_ptr[n_take] = (int32_t)a0[n_take];
  out_adj_ptr[a0[n_take]] = (int32_t)e0[n_take];
  DCGC_CHECK_ARG(e0[n_take] == 0 || (adj_idx && out_adj_idx), "dcgc_packed_take: null adjacency");
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t m = idx[i], s = atom_ptr[m], n = atom_ptr[m + 1] - s;
      const int64_t es = adj_ptr[s], shift = e0[i] - es;
      for (int64_t r = 0; r < n; ++r) out_adj_ptr[a0[i] + r] = (int32_t)(adj_ptr[s + r] + shift);
      if (e0[i + 1] > e0[i]) memcpy(out_adj_idx + e0[i], adj_idx + es, (size_t)(e0[i + 1] - e0[i]) * 4);
      if (features)
        memcpy(static_cast<char*>(out_features) + a0[i] * row_bytes, static_cast<const char*>(features) + s * row_bytes,
               (size_t)(n * row_bytes));
      if (features2)
        memcpy(static_cast<char*>(out_features2) + a0[i] * row_bytes2, static_cast<const char*>(features2) + s * row_bytes2,
               (size_t)(n * row_bytes2));
    }
  };
  if (n_threads <= 1 || n_take < 512) {
    work(0, n_take);
    return DCGC_OK;
  }
  std::vector<std::thread> pool;
  const int64_t chunk = (n_take + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t lo = t * chunk, hi = std::min(n_take, lo + chunk);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return DCGC_OK;
}

// Host table builder of the D-MPNN path: packed molecular graphs -> batched index tables.
//
// Replaces _MapperDMPNN (deepchem/models/torch_models/dmpnn.py:123-243: per-molecule Python loops), the
// batch-wide -1 padding of DMPNNModel.default_generator (dmpnn.py:741-753) and the PyG collation with
// _ModData.__inc__ (dmpnn.py:17-35) by one pass over flat arrays.  Integer work only; the two ELL tables
// (atom_to_incoming_bonds, mapping) must equal the reference's batched tensors bit for bit
// (tests/test_dmpnn_layout.py against tests/golden/ref_dmpnn.npz).
//
// Row space of the bond tensors (f_ini, message, h_message): molecule k owns rows
// [off_k, off_k + E_k] = its E_k directed bonds followed by one zero "pad" row, off_k = sum_{j<k}(E_j+1).
// The reference adds off_k to EVERY table entry, pads included, so a pad entry of molecule k reads row
// off_k - 1 (the pad row of molecule k-1; for k = 0 torch's negative index -1 = the last row of the batch).
// The CSR forms given to the gather kernels either drop the pad entries (exact when the pad rows are zero:
// bias == False) or keep them resolved to that row (keep_pads, exact for bias == True too).
#include <string.h>

#include <vector>

#include "common.h"

static void plan_offsets(dcgc_dmpnn_info* info) {
  const int64_t A = info->n_atoms, R = info->n_rows, B = info->n_mols, K = info->k;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = dcgc_align_up(off + bytes, 256);
    return o;
  };
  info->off_row_of_mol = take((B + 1) * 4);
  info->off_mol_ptr = take((B + 1) * 4);
  info->off_bond_src = take(R * 4);
  info->off_bond_edge = take(R * 4);
  info->off_a2b_ell = take(A * K * 4);
  info->off_map_ell = take(R * K * 4);
  info->off_a2b_ptr = take((A + 1) * 4);
  info->off_a2b_idx = take(info->n_a2b_entries * 4);
  info->off_a2b_t_ptr = take((R + 1) * 4);
  info->off_a2b_t_idx = take(info->n_a2b_entries * 4);
  info->off_map_ptr = take((R + 1) * 4);
  info->off_map_idx = take(info->n_map_entries * 4);
  info->off_map_t_ptr = take((R + 1) * 4);
  info->off_map_t_idx = take(info->n_map_entries * 4);
  info->slab_bytes = off;
}

// in-degree of every atom of molecule m (+ validation); returns the maximum
static int mol_indegree(int64_t m, const int32_t* node_ptr, const int32_t* edge_ptr, const int32_t* edge_src,
                        const int32_t* edge_dst, std::vector<int32_t>& indeg, int* status) {
  const int64_t n = node_ptr[m + 1] - node_ptr[m];
  indeg.assign((size_t)n, 0);
  int mx = 0;
  for (int64_t e = edge_ptr[m]; e < edge_ptr[m + 1]; ++e) {
    const int64_t s = edge_src[e], d = edge_dst[e];
    if (s < 0 || s >= n || d < 0 || d >= n) {
      dcgc_set_error("molecule %lld bond %lld: atom index outside [0,%lld)", (long long)m,
                     (long long)(e - edge_ptr[m]), (long long)n);
      *status = DCGC_ERR_INDEX;
      return 0;
    }
    const int v = ++indeg[(size_t)d];
    if (v > mx) mx = v;
  }
  return mx;
}

extern "C" int dcgc_dmpnn_plan(int64_t n_mols, const int32_t* node_ptr, const int32_t* edge_ptr,
                               const int32_t* edge_src, const int32_t* edge_dst, int32_t keep_pads,
                               dcgc_dmpnn_info* info) {
  DCGC_CHECK_ARG(n_mols >= 0 && node_ptr && edge_ptr && info, "dcgc_dmpnn_plan: null argument");
  DCGC_CHECK_ARG(node_ptr[0] == 0 && edge_ptr[0] == 0, "dcgc_dmpnn_plan: offset arrays must start at 0");
  memset(info, 0, sizeof(*info));
  info->n_mols = n_mols;
  info->n_atoms = node_ptr[n_mols];
  info->n_bonds = edge_ptr[n_mols];
  info->n_rows = info->n_bonds + n_mols;
  info->keep_pads = keep_pads ? 1 : 0;
  DCGC_CHECK_ARG(info->n_bonds == 0 || (edge_src && edge_dst), "dcgc_dmpnn_plan: null bond arrays");
  DCGC_CHECK_ARG(info->n_rows < ((int64_t)1 << 31) && info->n_atoms < ((int64_t)1 << 31),
                 "dcgc_dmpnn_plan: batch too large for int32 indices");
  std::vector<int32_t> indeg;
  int k = 1;                                    // max(1, ...) as in dmpnn.py:222-223 and :729
  int64_t live_map = 0;
  for (int64_t m = 0; m < n_mols; ++m) {
    DCGC_CHECK_ARG(node_ptr[m + 1] >= node_ptr[m] && edge_ptr[m + 1] >= edge_ptr[m],
                   "dcgc_dmpnn_plan: offsets must be non-decreasing");
    int st = DCGC_OK;
    const int mx = mol_indegree(m, node_ptr, edge_ptr, edge_src, edge_dst, indeg, &st);
    if (st != DCGC_OK) return st;
    if (mx > k) k = mx;
    // live mapping entries of bond b: incoming bonds of src(b) other than the reverse bond b ^ 1
    const int64_t e0 = edge_ptr[m], E = edge_ptr[m + 1] - e0;
    for (int64_t b = 0; b < E; ++b) {
      const int64_t s = edge_src[e0 + b], rev = b ^ 1;
      int64_t cnt = indeg[(size_t)s];
      if (rev < E && edge_dst[e0 + rev] == s) cnt -= 1;
      live_map += cnt;
    }
  }
  info->k = k;
  if (keep_pads) {
    info->n_a2b_entries = info->n_atoms * k;
    info->n_map_entries = info->n_rows * k;
  } else {
    info->n_a2b_entries = info->n_bonds;
    info->n_map_entries = live_map;
  }
  plan_offsets(info);
  return DCGC_OK;
}

// transpose of a CSR (ptr, idx) with n_src columns by counting sort; entries of one column end up in
// ascending row order
static void transpose_csr(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int64_t n_cols, int32_t* t_ptr,
                          int32_t* t_idx) {
  memset(t_ptr, 0, (size_t)(n_cols + 1) * 4);
  const int64_t nnz = ptr[n_rows];
  for (int64_t e = 0; e < nnz; ++e) t_ptr[idx[e] + 1]++;
  for (int64_t c = 0; c < n_cols; ++c) t_ptr[c + 1] += t_ptr[c];
  std::vector<int32_t> cur(t_ptr, t_ptr + n_cols);
  for (int64_t r = 0; r < n_rows; ++r)
    for (int32_t e = ptr[r]; e < ptr[r + 1]; ++e) t_idx[cur[(size_t)idx[e]]++] = (int32_t)r;
}

extern "C" int dcgc_dmpnn_build(int64_t n_mols, const int32_t* node_ptr, const int32_t* edge_ptr,
                                const int32_t* edge_src, const int32_t* edge_dst, const dcgc_dmpnn_info* info,
                                void* slab_v) {
  DCGC_CHECK_ARG(node_ptr && edge_ptr && info && slab_v, "dcgc_dmpnn_build: null argument");
  DCGC_CHECK_ARG(n_mols == info->n_mols && node_ptr[n_mols] == info->n_atoms && edge_ptr[n_mols] == info->n_bonds,
                 "dcgc_dmpnn_build: info does not match the inputs");
  char* slab = (char*)slab_v;
  const int64_t A = info->n_atoms, R = info->n_rows, K = info->k;
  int32_t* row_of_mol = (int32_t*)(slab + info->off_row_of_mol);
  int32_t* mol_ptr = (int32_t*)(slab + info->off_mol_ptr);
  int32_t* bond_src = (int32_t*)(slab + info->off_bond_src);
  int32_t* bond_edge = (int32_t*)(slab + info->off_bond_edge);
  int32_t* a2b_ell = (int32_t*)(slab + info->off_a2b_ell);
  int32_t* map_ell = (int32_t*)(slab + info->off_map_ell);
  int32_t* a2b_ptr = (int32_t*)(slab + info->off_a2b_ptr);
  int32_t* a2b_idx = (int32_t*)(slab + info->off_a2b_idx);
  int32_t* map_ptr = (int32_t*)(slab + info->off_map_ptr);
  int32_t* map_idx = (int32_t*)(slab + info->off_map_idx);

  // the row a -1 entry of molecule k refers to (torch negative indexing for k = 0)
  auto resolve = [&](int64_t v) -> int32_t { return (int32_t)(v < 0 ? v + R : v); };

  int64_t off = 0, na = 0, nm = 0;
  std::vector<int32_t> fill;
  for (int64_t m = 0; m < n_mols; ++m) {
    const int64_t a0 = node_ptr[m], n = node_ptr[m + 1] - a0;
    const int64_t e0 = edge_ptr[m], E = edge_ptr[m + 1] - e0;
    row_of_mol[m] = (int32_t)off;
    mol_ptr[m] = (int32_t)a0;
    // a2b: incoming bonds of every atom in ascending bond order (np.where, dmpnn.py:217-219), -1 padded,
    // every entry shifted by the molecule's row offset (dmpnn.py:27-35)
    for (int64_t a = 0; a < n; ++a)
      for (int64_t j = 0; j < K; ++j) a2b_ell[(a0 + a) * K + j] = (int32_t)(off - 1);
    fill.assign((size_t)n, 0);
    for (int64_t b = 0; b < E; ++b) {
      const int64_t d = edge_dst[e0 + b];
      if (d < 0 || d >= n || edge_src[e0 + b] < 0 || edge_src[e0 + b] >= n) {
        dcgc_set_error("molecule %lld bond %lld: atom index outside [0,%lld)", (long long)m, (long long)b,
                       (long long)n);
        return DCGC_ERR_INDEX;
      }
      int32_t& f = fill[(size_t)d];
      if (f >= K) {
        dcgc_set_error("dcgc_dmpnn_build: in-degree changed since dcgc_dmpnn_plan");
        return DCGC_ERR_INVALID;
      }
      a2b_ell[(a0 + d) * K + f++] = (int32_t)(off + b);
    }
    // mapping: a2b row of the bond's initial atom with the reverse bond (b ^ 1) masked (dmpnn.py:203, 234-243)
    for (int64_t b = 0; b < E; ++b) {
      const int64_t s = edge_src[e0 + b];
      const int32_t rev = (int32_t)(off + (b ^ 1));
      bond_src[off + b] = (int32_t)(a0 + s);
      bond_edge[off + b] = (int32_t)(e0 + b);
      for (int64_t j = 0; j < K; ++j) {
        const int32_t v = a2b_ell[(a0 + s) * K + j];
        map_ell[(off + b) * K + j] = (v == rev && (b ^ 1) < E) ? (int32_t)(off - 1) : v;
      }
    }
    bond_src[off + E] = -1;                         // the zero pad row (dmpnn.py:187-188, 207-208)
    bond_edge[off + E] = -1;
    for (int64_t j = 0; j < K; ++j) map_ell[(off + E) * K + j] = (int32_t)(off - 1);
    // CSR forms
    const int32_t pad = (int32_t)(off - 1);
    for (int64_t a = 0; a < n; ++a) {
      a2b_ptr[a0 + a] = (int32_t)na;
      for (int64_t j = 0; j < K; ++j) {
        const int32_t v = a2b_ell[(a0 + a) * K + j];
        if (v != pad || j < fill[(size_t)a]) { a2b_idx[na++] = v; }
        else if (info->keep_pads) a2b_idx[na++] = resolve(v);
      }
    }
    for (int64_t b = 0; b <= E; ++b) {
      map_ptr[off + b] = (int32_t)nm;
      for (int64_t j = 0; j < K; ++j) {
        const int32_t v = map_ell[(off + b) * K + j];
        if (v != pad) map_idx[nm++] = v;
        else if (info->keep_pads) map_idx[nm++] = resolve(v);
      }
    }
    off += E + 1;
  }
  row_of_mol[n_mols] = (int32_t)off;
  mol_ptr[n_mols] = (int32_t)A;
  a2b_ptr[A] = (int32_t)na;
  map_ptr[R] = (int32_t)nm;
  if (na != info->n_a2b_entries || nm != info->n_map_entries) {
    dcgc_set_error("dcgc_dmpnn_build: entry counts differ from dcgc_dmpnn_plan (%lld/%lld vs %lld/%lld)",
                   (long long)na, (long long)nm, (long long)info->n_a2b_entries, (long long)info->n_map_entries);
    return DCGC_ERR_INVALID;
  }
  transpose_csr(a2b_ptr, a2b_idx, A, R, (int32_t*)(slab + info->off_a2b_t_ptr), (int32_t*)(slab + info->off_a2b_t_idx));
  transpose_csr(map_ptr, map_idx, R, R, (int32_t*)(slab + info->off_map_t_ptr), (int32_t*)(slab + info->off_map_t_idx));
  return DCGC_OK;
}

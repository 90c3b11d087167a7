/*
 * dcgc.h — C ABI of libdcgc: the B200 (sm_100a) GraphConv / GraphPool / GraphGather /
 * DMPNN message-passing path for DeepChem.
 *
 * The reference (pandegroup/deepchem) has no native code and therefore no FFI to mirror;
 * each entry point below replaces the Python/libtorch code cited next to it and is what a
 * maintainer would bind (ctypes stub in INTEGRATION.md).  Conventions:
 *   - every function returns an int status (DCGC_OK == 0, negative on error) and never
 *     throws or aborts; dcgc_last_error() returns a thread-local message.
 *   - "dev" pointers are device pointers owned by the caller (PyTorch caching allocator on
 *     the Python side); kernels only write caller-provided outputs / workspaces.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All device
 *     functions are asynchronous on that stream and re-entrant across streams.
 *   - matrices are row-major with an explicit leading dimension `ld*` (in elements).
 *   - integer layout arrays are int32 unless stated; deg_slice is int64 as in the reference.
 */
#ifndef DCGC_H_
#define DCGC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCGC_OK 0
#define DCGC_ERR_INVALID (-1) /* bad argument (null pointer, negative size, unsupported width) */
#define DCGC_ERR_DEGREE (-2)  /* an atom has more than max_deg neighbours (feat/graph_features.py:34-36) */
#define DCGC_ERR_INDEX (-3)   /* neighbour index outside its molecule */
#define DCGC_ERR_CUDA (-4)    /* CUDA runtime error; message has the cudaError string */
#define DCGC_ERR_NOMEM (-5)   /* workspace too small */

#define DCGC_MAX_DEG 10
#define DCGC_N_DEG 11

/* activation codes for fused epilogues */
#define DCGC_ACT_NONE 0
#define DCGC_ACT_RELU 1
#define DCGC_ACT_TANH 2

/* GEMM arithmetic modes */
#define DCGC_GEMM_FP32 0 /* SIMT FFMA, fp32 in / fp32 accumulate: the 1e-5 parity mode        */
#define DCGC_GEMM_BF16 1 /* tcgen05, operands rounded to bf16 / fp32 accumulate: the 2e-2 mode     */
#define DCGC_GEMM_TF32X3 2 /* tcgen05 kind::tf32 with 3-term split: fp32-grade accuracy       */
#define DCGC_GEMM_F16X3 3 /* forward entry points (dcgc_group_gemm_fwd*, dcgc_linear_fwd): the same 3-term split with
                             fp16 halves (kind::f16, K 16 per instruction: half the tensor-core instructions, the same
                             11 + 11 significand bits) for operands inside fp16's range, |x| < 65 504 — activations,
                             not gradients; everywhere else it means DCGC_GEMM_TF32X3 */

const char* dcgc_last_error(void);
int dcgc_version(void);
/* 1 if a CUDA device with compute capability 10.x is visible, 0 if not, <0 on error */
int dcgc_device_ok(void);

/* Launch accounting and live per-entry-point timing (used by bench.py): dcgc_launch_count()
 * returns the number of kernels this library has launched in the process; between
 * dcgc_profile_begin("dcgc_gather_sum") and dcgc_profile_end() every launch made by that entry
 * point is bracketed by CUDA events on its own stream; _end synchronises them and returns the
 * summed duration and the number of bracketed calls. */
long long dcgc_launch_count(void);
int dcgc_profile_begin(const char* entry_name);
int dcgc_profile_end(double* total_ms, long long* launches);
/* dcgc_profile_begin("*") brackets every entry point and every internal stage of the model engine;
 * dcgc_profile_report ends such a session and writes one line "scope total_ms calls" per scope. */
int dcgc_profile_report(char* out, int64_t cap);

/* --------------------------------------------------------------------------------------------
 * Host layout builder.  Replaces ConvMol.agglomerate_mols (deepchem/feat/mol_graphs.py:256-349)
 * and the per-molecule degree sort of ConvMol.__init__ (mol_graphs.py:113-185): the composite of
 * the two stable sorts is one stable counting sort of all atoms by degree.
 *
 * Input: a packed shard (no Python objects): atom_ptr[n_mols+1], adj_ptr[n_atoms+1],
 * adj_idx[n_edges] (molecule-local neighbour ids in adjacency-list order).
 * Output: one host slab (caller-allocated, ideally pinned) holding every array the device path
 * needs, at the byte offsets reported in dcgc_layout_info; a single H2D copy moves it.
 * Bit-exact with the reference for deg_slice (int64 [11,2], running starts), membership and
 * deg_adj_1..10 (col_idx is their concatenation, flattened row-major).
 * ------------------------------------------------------------------------------------------ */
typedef struct dcgc_layout_info {
  int64_t n_mols;     /* molecules in this batch                                    */
  int64_t n_segments; /* GraphGather segments (batch_size >= n_mols; extras are empty) */
  int64_t n_atoms;    /* N                                                          */
  int64_t n_edges;    /* E = sum_d d*N_d directed neighbour entries                 */
  int64_t n_tiles;    /* row tiles of at most tile_rows rows that never straddle a degree bucket */
  int32_t tile_rows;
  int32_t reserved;
  int64_t deg_count[DCGC_N_DEG]; /* N_d (host copy of deg_slice[:,1]) */
  /* byte offsets into the slab (each 256-byte aligned) */
  int64_t off_deg_slice;  /* int64 [11,2]                                              */
  int64_t off_membership; /* int32 [N]                                                 */
  int64_t off_perm;       /* int32 [N]   row i of the batch = atom perm[i] of the packed shard */
  int64_t off_row_ptr;    /* int32 [N+1]                                               */
  int64_t off_col_idx;    /* int32 [E]   == concat(deg_adj_1.flatten(), ..., deg_adj_10.flatten()) */
  int64_t off_t_row_ptr;  /* int32 [N+1] transposed CSR                                */
  int64_t off_t_src;      /* int32 [E]   row i referencing this row, ordered by (i, slot) */
  int64_t off_t_slot;     /* int32 [E]   slot k of that reference in row i's list       */
  int64_t off_mol_ptr;    /* int32 [n_segments+1]                                      */
  int64_t off_mol_atoms;  /* int32 [N]   rows of each molecule, ascending              */
  int64_t off_tiles;      /* int32 [n_tiles,4] = (row0, n_rows, degree, 0)             */
  int64_t slab_bytes;
  /* Molecule groups for the shared-memory staged kernels (dcgc_mg_*): consecutive molecules are packed
   * greedily into groups of at most group_rows atoms (a larger molecule is a group of its own).  Inside a degree
   * bucket the rows are ordered by molecule, so the rows of a group are ONE contiguous range per bucket: a
   * CTA stages them with one bulk copy per bucket and finds every neighbour in shared memory (no edge leaves
   * a molecule).  int32 [DCGC_GROUP_STRIDE * (1 + n_groups_alloc + 1)]:
   *   header  {n_groups, max rows of a group, max neighbour entries of a group, valid, group_rows, 0...}
   *   row g   {first row of group g in bucket 0..10, first molecule of group g}; row n_groups closes the last
   * `valid` is 0 when the layout does not have the property (rows of a bucket not ordered by molecule, or a
   * neighbour in another molecule: possible only for hand-made dcgc_layout_build_from_deg inputs); the staged
   * kernels then must not be used. */
  int64_t off_groups;
  int32_t group_rows;
  int32_t n_groups_alloc;
} dcgc_layout_info;
#define DCGC_GROUP_STRIDE 12
#define DCGC_GROUP_ROWS_DEFAULT 128

/* Pass 1: validate degrees, count, and compute sizes / offsets.  O(N). */
int dcgc_layout_plan(int64_t n_mols, const int32_t* atom_ptr, const int32_t* adj_ptr,
                     int64_t n_segments, int32_t tile_rows, dcgc_layout_info* info);
/* Pass 2: fill the slab.  `info` must come from dcgc_layout_plan on the same inputs. */
int dcgc_layout_build(int64_t n_mols, const int32_t* atom_ptr, const int32_t* adj_ptr,
                      const int32_t* adj_idx, const dcgc_layout_info* info, void* slab);
/* Host feature permutation into degree-major order (the numpy-facing MultiConvMol path):
 * dst[i, 0:n_feat] = src[perm[i], 0:n_feat]; columns n_feat..ld_dst-1 are zeroed. */
int dcgc_layout_permute_features_host(const float* src, int64_t ld_src, const int32_t* perm,
                                      int64_t n_atoms, int32_t n_feat, float* dst, int64_t ld_dst,
                                      int32_t n_threads);
/* Gather of molecules out of a packed shard (what a shuffled epoch does for every batch: DiskDataset.iterbatches
 * with deterministic=False permutes the sample indices, deepchem/data/datasets.py:1598-1730).  idx[n_take] are molecule
 * numbers of the source shard (repeats allowed: pad_batch tiles them, datasets.py:142-218).  _plan returns the sizes
 * of the outputs; features / features2 are optional per-atom row-major matrices (row_bytes bytes per atom each: the
 * fp32 feature matrix and its exact int8 copy) gathered with one memcpy per molecule, into pinned staging memory when
 * the caller provides it.  Adjacency entries are molecule-local and are copied unchanged. */
int dcgc_packed_take_plan(int64_t n_take, const int64_t* idx, int64_t n_src_mols, const int32_t* atom_ptr,
                          const int32_t* adj_ptr, int64_t* n_atoms_out, int64_t* n_entries_out);
int dcgc_packed_take(int64_t n_take, const int64_t* idx, int64_t n_src_mols, const int32_t* atom_ptr,
                     const int32_t* adj_ptr, const int32_t* adj_idx, const void* features, int64_t row_bytes,
                     const void* features2, int64_t row_bytes2, int32_t* out_atom_ptr, int32_t* out_adj_ptr,
                     int32_t* out_adj_idx, void* out_features, void* out_features2, int32_t n_threads);

/* Derive the same slab from an already-agglomerated reference layout (deg_slice, membership,
 * col_idx = concatenated deg_adj lists): used when a caller hands the layers plain tensors. */
int dcgc_layout_plan_from_deg(const int64_t* deg_slice, int64_t n_segments, int32_t tile_rows,
                              dcgc_layout_info* info);
int dcgc_layout_build_from_deg(const int64_t* deg_slice, const int32_t* membership,
                               const int32_t* col_idx, const dcgc_layout_info* info, void* slab);

/* --------------------------------------------------------------------------------------------
 * Device kernels.  fp32 activations, row-major, leading dimensions in floats.
 * ------------------------------------------------------------------------------------------ */

/* Host -> device upload of `bytes` bytes from PINNED host memory as a train of cudaMemcpyAsync calls of at most
 * `chunk_bytes` each on `stream` (chunk_bytes <= 0: one copy).  Replaces the per-array `torch.as_tensor(x, device)`
 * uploads of TorchModel._prepare_batch (torch_models/torch_model.py:923-952).  One 30 MB copy running beside the
 * training kernels slows them measurably on B200; moderate chunks do not (profiles/r2_interference.md). */
int dcgc_h2d_chunked(void* dst_dev, const void* src_host_pinned, int64_t bytes, int64_t chunk_bytes, void* stream);

/* dst[i, 0:n_feat] = src[perm[i], 0:n_feat], pad columns zeroed (device-side feature permute,
 * replaces `atoms_by_deg[order]`, mol_graphs.py:277). */
int dcgc_permute_rows(const float* src_dev, int64_t ld_src, const int32_t* perm_dev, int64_t n_rows,
                      int32_t n_feat, float* dst_dev, int64_t ld_dst, void* stream);

/* Same from an int8 feature matrix (exact conversion to fp32; dst rows 16-byte aligned, ld_dst % 4 == 0): the packed
 * shard format stores integer-valued feature matrices (every ConvMol feature of deepchem/feat/graph_features.py:
 * 282-391 is a one-hot, a formal charge or a radical count) as int8, a quarter of the upload. */
int dcgc_permute_rows_i8(const int8_t* src_dev, int64_t ld_src, const int32_t* perm_dev, int64_t n_rows,
                         int32_t n_feat, float* dst_dev, int64_t ld_dst, void* stream);

/* K1 / K5 / K8 — CSR gather-sum: out[i,:] = sum_{e in [row_ptr[i], row_ptr[i+1])} x[idx[e], :].
 * Forward of GraphConv.sum_neigh (torch_models/layers.py:6236-6246) with (row_ptr, col_idx);
 * its backward (the transposed scatter) with (t_row_ptr, t_src); DMPNN `message[mapping].sum(1)`
 * (layers.py:1629) with an ELL table turned into CSR.  Rows with no entries are written as zeros.
 * If addend_dev is non-null, out[i,:] = addend[i,:] + sum (out may alias addend): this fuses the
 * `self-path + neighbour-path` add of the GraphConv input gradient.
 * No atomics: one thread owns one 16-byte group of one output row; summation order = index order. */
int dcgc_gather_sum(const float* x_dev, int64_t ld_x, const int32_t* row_ptr_dev,
                    const int32_t* idx_dev, int64_t n_rows_out, int32_t width,
                    const float* addend_dev, int64_t ld_add, float* out_dev, int64_t ld_out,
                    void* stream);

/* Same gather for DEGREE-BUCKETED rows (the ConvMol layout: rows sorted by degree, row i of bucket d has
 * exactly d entries, idx = concat(deg_adj_1.flatten(), ...)): the CSR offsets are computed from the 11 bucket
 * sizes (host array) instead of being loaded, which removes one of the three dependent loads per thread.
 * Also valid for the transposed lists of a symmetric adjacency (in-degree == degree). */
int dcgc_gather_sum_bucketed(const float* x_dev, int64_t ld_x, const int64_t* deg_count_host,
                             const int32_t* idx_dev, int64_t n_rows_out, int32_t width, const float* addend_dev,
                             int64_t ld_add, float* out_dev, int64_t ld_out, void* stream);

/* K3 — GraphPool forward (layers.py:6342-6367): out[i,c] = max(x[i,c], max_k x[col[row_ptr[i]+k],c]).
 * If scale/shift are non-null the per-channel affine y = x*scale[c] + shift[c] (a folded
 * BatchNorm) is applied to every loaded element first.  arg_dev (uint8 [N, ld_arg], may be null
 * for inference) receives the FIRST slot attaining the max: 0 = self, k+1 = neighbour k. */
int dcgc_pool_fwd(const float* x_dev, int64_t ld_x, const float* scale_dev, const float* shift_dev,
                  const int32_t* row_ptr_dev, const int32_t* col_idx_dev, int64_t n_rows,
                  int32_t width, float* out_dev, int64_t ld_out, uint8_t* arg_dev, int64_t ld_arg,
                  void* stream);
/* K7 — GraphPool backward without scatter: dx[j,c] = [arg[j,c]==0]*dy[j,c]
 *      + sum_{e in T(j)} [arg[t_src[e],c] == t_slot[e]+1] * dy[t_src[e],c].
 * If scale is non-null the result is multiplied by scale[c] (chain rule of the folded affine). */
int dcgc_pool_bwd(const float* dy_dev, int64_t ld_dy, const uint8_t* arg_dev, int64_t ld_arg,
                  const float* scale_dev, const int32_t* t_row_ptr_dev, const int32_t* t_src_dev,
                  const int32_t* t_slot_dev, int64_t n_rows, int32_t width, float* dx_dev,
                  int64_t ld_dx, void* stream);

/* K4 — GraphGather forward (layers.py:6464-6479; pytorch_utils.py:20-74, 473-528):
 * out[g, 0:D] = act(sum_{i in mol g} x[i,:]), out[g, D:2D] = act(max_{i in mol g} x[i,:]),
 * empty segment -> sum 0, max -inf (tanh -> -1).  argrow_dev (int32 [n_seg, D], may be null)
 * receives the lowest row index attaining the max (-1 for empty segments).  A non-null
 * scale/shift applies the per-channel affine x*scale[c] + shift[c] (folded BatchNorm) on load. */
int dcgc_gather_fwd(const float* x_dev, int64_t ld_x, const float* scale_dev, const float* shift_dev,
                    const int32_t* mol_ptr_dev, const int32_t* mol_atoms_dev, int64_t n_segments,
                    int32_t width, int32_t act, float* out_dev, int64_t ld_out, int32_t* argrow_dev,
                    void* stream);
/* K7 — GraphGather backward: dx[i,c] = dsum[m,c] + [argrow[m,c]==i]*dmax[m,c] with m =
 * membership[i] and d* = dout * act'(out) (act' from the saved output). */
int dcgc_gather_bwd(const float* dout_dev, int64_t ld_dout, const float* out_dev, int64_t ld_out,
                    const int32_t* argrow_dev, const int32_t* membership_dev, int64_t n_rows,
                    int32_t width, int32_t act, float* dx_dev, int64_t ld_dx, void* stream);

/* K2 — degree-grouped affine map (GraphConv.forward, layers.py:6202-6229; also the atom-level
 * Dense and the DMPNN Linear layers with a single group):
 *   y[r,:] = act( a1[r,0:k1] . W[g][0:k1,:] + a2[r,0:k2] . W[g][k1:k1+k2,:] + bias[g,:] )
 * for every row r of every tile, g = tile degree (tiles from the layout slab; g = 0 for all rows
 * when tiles_dev is null and the whole [0,n_rows) range is one group).  W is [n_groups, k1+k2, n]
 * row-major; a2 may be null (k2 = 0).  bias may be null. */
int dcgc_group_gemm_fwd(int32_t mode, const float* a1_dev, int64_t ld_a1, int32_t k1,
                        const float* a2_dev, int64_t ld_a2, int32_t k2, const float* w_dev,
                        const float* bias_dev, int32_t n, const int32_t* tiles_dev, int64_t n_tiles,
                        int32_t tile_rows, int64_t n_rows, int32_t act, float* y_dev, int64_t ld_y,
                        void* stream);
/* Same, plus fused per-column statistics of the stored output for the BatchNorm that follows
 * (graphconvmodel.py:213-216): stats_part_dev [n_chunks][2][n] float64 receives, per chunk of rows, the
 * column sums of y and of y*y; *n_chunks_out (host) is the number of chunks written, at most
 * dcgc_gemm_stats_max_chunks().  Partials are combined in a fixed order (deterministic).
 * Tensor-core modes only (the statistics are accumulated in the tcgen05 kernel's epilogue). */
int dcgc_group_gemm_fwd_stats(int32_t mode, const float* a1_dev, int64_t ld_a1, int32_t k1,
                              const float* a2_dev, int64_t ld_a2, int32_t k2, const float* w_dev,
                              const float* bias_dev, int32_t n, const int32_t* tiles_dev, int64_t n_tiles,
                              int32_t tile_rows, int64_t n_rows, int32_t act, float* y_dev, int64_t ld_y,
                              double* stats_part_dev, int32_t* n_chunks_out, void* stream);
int32_t dcgc_gemm_stats_max_chunks(void);
/* K6 (dgrad): [d1 | d2][r,:] = g[r,0:n] . W[g]^T, columns 0:k1 to d1 and k1:k1+k2 to d2 (either
 * may be null to skip).  Takes the same W as the forward (reads it transposed). */
int dcgc_group_gemm_dgrad(int32_t mode, const float* g_dev, int64_t ld_g, int32_t n,
                          const float* w_dev, int32_t k1, int32_t k2, const int32_t* tiles_dev,
                          int64_t n_tiles, int32_t tile_rows, int64_t n_rows, float* d1_dev,
                          int64_t ld_d1, float* d2_dev, int64_t ld_d2, void* stream);
/* K6 (wgrad): dW[g] = [a1 | a2]_g^T . grad_g  ([k1+k2, n]) and dbias[g] = column sums of grad_g,
 * per group, deterministic (fixed split of the rows, partials reduced in a fixed order).
 * workspace: dcgc_group_gemm_wgrad_workspace() bytes.  Groups with no rows get zeros. */
int64_t dcgc_group_gemm_wgrad_workspace(int32_t k1, int32_t k2, int32_t n, int32_t n_groups);
int dcgc_group_gemm_wgrad(int32_t mode, const float* a1_dev, int64_t ld_a1, int32_t k1,
                          const float* a2_dev, int64_t ld_a2, int32_t k2, const float* g_dev,
                          int64_t ld_g, int32_t n, const int64_t* deg_count_host, int32_t n_groups,
                          float* dw_dev, float* dbias_dev, void* workspace_dev,
                          int64_t workspace_bytes, void* stream);

/* nn.Linear-layout variants (weight [n_out, k_in] as torch stores it): the atom-level Dense of the
 * model (graphconvmodel.py:172,222) and the DMPNN W_i / W_h / W_o (layers.py:1510-1517).
 *   fwd   y = act(x . w^T + bias)          dgrad  dx = g . w
 *   wgrad dw[n,k] = g^T . x, dbias[n] = column sums of g   (deterministic, same workspace rule) */
int dcgc_linear_fwd(int32_t mode, const float* x_dev, int64_t ld_x, int32_t k, const float* w_dev,
                    const float* bias_dev, int32_t n, int64_t n_rows, int32_t act, float* y_dev,
                    int64_t ld_y, void* stream);
int dcgc_linear_fwd_stats(int32_t mode, const float* x_dev, int64_t ld_x, int32_t k, const float* w_dev,
                          const float* bias_dev, int32_t n, int64_t n_rows, int32_t act, float* y_dev,
                          int64_t ld_y, double* stats_part_dev, int32_t* n_chunks_out, void* stream);
int dcgc_linear_dgrad(int32_t mode, const float* g_dev, int64_t ld_g, int32_t n, const float* w_dev,
                      int32_t k, int64_t n_rows, float* dx_dev, int64_t ld_dx, void* stream);
int64_t dcgc_linear_wgrad_workspace(int32_t k, int32_t n);
int dcgc_linear_wgrad(int32_t mode, const float* x_dev, int64_t ld_x, int32_t k, const float* g_dev,
                      int64_t ld_g, int32_t n, int64_t n_rows, float* dw_dev, float* dbias_dev,
                      void* workspace_dev, int64_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * D-MPNN host table builder.  Replaces _MapperDMPNN (deepchem/models/torch_models/dmpnn.py:123-243),
 * the batch-wide -1 padding of DMPNNModel.default_generator (dmpnn.py:741-753) and the PyG collation
 * with _ModData.__inc__ (dmpnn.py:17-35).
 *
 * Input: packed graphs — node_ptr[n_mols+1], edge_ptr[n_mols+1], edge_src / edge_dst[n_bonds]
 * (molecule-local atom ids; directed bonds in (i->j, j->i) pairs at positions (2k, 2k+1) as
 * DMPNNFeaturizer emits them).  Row space of the bond tensors: molecule k owns rows
 * [row_of_mol[k], row_of_mol[k] + E_k] = its E_k bonds + one zero pad row.
 * Output slab (int32 arrays at the byte offsets below, 256-byte aligned):
 *   a2b_ell [n_atoms, k], map_ell [n_rows, k]   == the reference's batched atom_to_incoming_bonds /
 *       mapping tensors (every entry shifted by the molecule's row offset, pads included; int64 there)
 *   bond_src [n_rows]   global atom row feeding f_ini (-1 on pad rows), bond_edge [n_rows] global bond id
 *   (a2b_ptr, a2b_idx) CSR atom <- bond rows, (map_ptr, map_idx) CSR bond row <- bond rows, and their
 *   transposes for the backward gathers.  keep_pads == 0 drops the pad entries (exact while the pad rows
 *   are zero, i.e. bias == False); keep_pads == 1 keeps them, resolved to the row torch would read.
 * ------------------------------------------------------------------------------------------ */
typedef struct dcgc_dmpnn_info {
  int64_t n_mols, n_atoms, n_bonds, n_rows; /* n_rows = n_bonds + n_mols */
  int64_t k;                                /* batch-wide max in-degree, at least 1 */
  int64_t n_a2b_entries, n_map_entries;     /* CSR entry counts */
  int32_t keep_pads, reserved;
  int64_t off_row_of_mol; /* int32 [n_mols+1] */
  int64_t off_mol_ptr;    /* int32 [n_mols+1] first atom of each molecule */
  int64_t off_bond_src, off_bond_edge;
  int64_t off_a2b_ell, off_map_ell;
  int64_t off_a2b_ptr, off_a2b_idx, off_a2b_t_ptr, off_a2b_t_idx;
  int64_t off_map_ptr, off_map_idx, off_map_t_ptr, off_map_t_idx;
  int64_t slab_bytes;
} dcgc_dmpnn_info;

int dcgc_dmpnn_plan(int64_t n_mols, const int32_t* node_ptr, const int32_t* edge_ptr, const int32_t* edge_src,
                    const int32_t* edge_dst, int32_t keep_pads, dcgc_dmpnn_info* info);
int dcgc_dmpnn_build(int64_t n_mols, const int32_t* node_ptr, const int32_t* edge_ptr, const int32_t* edge_src,
                     const int32_t* edge_dst, const dcgc_dmpnn_info* info, void* slab);

/* f_ini_atoms_bonds on the device (dmpnn.py:183-188): out[r, 0:fa] = atom_feat[bond_src[r], :],
 * out[r, fa:fa+fb] = bond_feat[bond_edge[r], :], zero rows where bond_src[r] < 0; columns
 * fa+fb..ld_out-1 are zeroed. */
int dcgc_dmpnn_concat_rows(const float* atom_feat_dev, int64_t ld_a, int32_t fa, const float* bond_feat_dev,
                           int64_t ld_b, int32_t fb, const int32_t* bond_src_dev, const int32_t* bond_edge_dev,
                           int64_t n_rows, float* out_dev, int64_t ld_out, void* stream);
/* Readout (torch_models/layers.py:1550-1583): out[m, :] = scale_m * sum_{a in [mol_ptr[m], mol_ptr[m+1])} x[a, :]
 * with scale_m = 1/n_atoms(m) (mode 0, 'mean'), 1 (mode 1, 'sum') or 1/norm (mode 2, 'norm'); rows summed in
 * ascending order.  The backward broadcasts: dx[a, :] = scale_m * dout[m(a), :]. */
int dcgc_segment_readout_fwd(const float* x_dev, int64_t ld_x, const int32_t* mol_ptr_dev, int64_t n_mols,
                             int32_t width, int32_t mode, float norm, float* out_dev, int64_t ld_out, void* stream);
int dcgc_segment_readout_bwd(const float* dout_dev, int64_t ld_dout, const int32_t* mol_ptr_dev, int64_t n_mols,
                             int64_t n_atoms, int32_t width, int32_t mode, float norm, float* dx_dev, int64_t ld_dx,
                             void* stream);

/* --------------------------------------------------------------------------------------------
 * Fused D-MPNN model: DMPNN.forward (deepchem/models/torch_models/dmpnn.py:246-449: DMPNNEncoderLayer,
 * torch_models/layers.py:1585-1649, + PositionwiseFeedForward, :795-910), the L2 loss under _StandardLoss
 * (models/losses.py:76-94, torch_model.py:1267-1294) and the backward of all of it as one call over flat
 * parameter / gradient slabs (the optimizer is dcgc_adam_step on the same slabs).  Covered configuration:
 * ReLU activations, dropout 0, encoder bias False, regression, no global features, widths that are
 * multiples of 4, >= 2 feed-forward linears, depth >= 2; anything else is served by the per-layer entry
 * points above.
 * ------------------------------------------------------------------------------------------ */
#define DCGC_DMPNN_MAX_FFN 8
typedef struct dcgc_dmpnn_model_config {
  int32_t atom_fdim, bond_fdim, hidden, depth;
  int32_t ffn_layers, ffn_hidden, n_out; /* ffn_layers linears: hidden -> ffn_hidden ... -> n_out */
  int32_t aggregation;                   /* 0 mean, 1 sum, 2 norm (layers.py:1550-1583) */
  float aggregation_norm;
  int32_t gemm_mode;
} dcgc_dmpnn_model_config;

/* Device pointers into the uploaded slab of dcgc_dmpnn_build (keep_pads == 0). */
typedef struct dcgc_dmpnn_tables {
  int64_t n_mols, n_atoms, n_rows;
  const int32_t* mol_ptr;                               /* [n_mols+1] first atom of each molecule */
  const int32_t *a2b_ptr, *a2b_idx, *a2b_t_ptr, *a2b_t_idx; /* atom <- bond rows and its transpose */
  const int32_t *map_ptr, *map_idx, *map_t_ptr, *map_t_idx; /* bond row <- bond rows and its transpose */
} dcgc_dmpnn_tables;

/* offsets (floats into the slabs): W_i [hidden, atom+bond], W_h [hidden, hidden], W_o [hidden, atom+hidden],
 * b_o [hidden], then (W [out, in], b [out]) per feed-forward linear: 4 + 2 * ffn_layers entries, nn.Linear layout. */
int dcgc_dmpnn_model_layout(const dcgc_dmpnn_model_config* cfg, int64_t* offsets, int64_t* n_params);
int64_t dcgc_dmpnn_model_workspace_bytes(const dcgc_dmpnn_model_config* cfg, int64_t n_rows, int64_t n_atoms,
                                         int64_t n_mols);
/* out [n_mols, n_out] dense; encoding_out (optional) [n_mols, hidden] = the encoder's molecule vectors. */
int dcgc_dmpnn_model_forward(const dcgc_dmpnn_model_config* cfg, const dcgc_dmpnn_tables* tables,
                             const float* atom_feat_dev, int64_t ld_af, const float* f_ini_dev, int64_t ld_fi,
                             const float* params_dev, void* workspace_dev, int64_t workspace_bytes, float* out_dev,
                             float* encoding_out_dev, void* stream);
/* y, w: [n_mols, n_out] (w may be NULL = ones); loss_dev receives mean(w * (out - y)^2); every gradient is
 * written (not accumulated) into grads_dev; out_dev (optional) receives the predictions [n_mols, n_out]. */
int dcgc_dmpnn_model_train_step(const dcgc_dmpnn_model_config* cfg, const dcgc_dmpnn_tables* tables,
                                const float* atom_feat_dev, int64_t ld_af, const float* f_ini_dev, int64_t ld_fi,
                                const float* y_dev, const float* w_dev, const float* params_dev, float* grads_dev,
                                void* workspace_dev, int64_t workspace_bytes, float* loss_dev, float* out_dev,
                                void* stream);

/* --------------------------------------------------------------------------------------------
 * Whole-model engine: GraphConvModel forward / loss / backward in one call over flat slabs.
 * Replaces the per-step Python of TorchModel.fit_generator (torch_model.py:428-443) +
 * _GraphConvTorchModel.forward (graphconvmodel.py:188-249) + autograd.  Layer widths and the
 * dense width must be multiples of 4; x must be zero-padded to a leading dimension that is a
 * multiple of 4.
 *
 * Parameter slab (fp32, offsets from dcgc_gcmodel_layout, every tensor 16-byte aligned):
 *   per conv layer l: W [11, 2*Fp_l, C_l]  (group d: rows 0:F self weight W_list[2(d-1)+1], rows
 *                       Fp:Fp+F neighbour weight W_list[2(d-1)]; group 0: W_list[20]; pad rows 0),
 *                     b [21, C_l] (reference order), BN gamma [C_l], beta [C_l]
 *   dense: W [D, C_L] (nn.Linear layout), b [D], BN gamma [D], beta [D]
 *   head:  W [n_out, 2D], b [n_out]
 * The gradient slab has the same layout.  bn_running holds (running_mean, running_var) per BN.
 * ------------------------------------------------------------------------------------------ */
#define DCGC_MODEL_MAX_LAYERS 8

typedef struct dcgc_topology {
  int64_t n_atoms, n_edges, n_segments, n_tiles;
  int64_t deg_count[DCGC_N_DEG]; /* host copy of N_d */
  const int32_t* row_ptr;
  const int32_t* col_idx;
  const int32_t* t_row_ptr;
  const int32_t* t_src;
  const int32_t* t_slot;
  const int32_t* mol_ptr;
  const int32_t* mol_atoms;
  const int32_t* membership;
  const int32_t* tiles; /* device pointers into the uploaded layout slab */
  int32_t symmetric;    /* 1 if t_row_ptr == row_ptr (every atom's in-degree equals its degree: true for
                           molecular graphs, false e.g. with a master atom, feat/graph_features.py:906-909) */
  int32_t reserved;
  /* molecule-group table (device pointer to the first table row, i.e. past the header) and the header values
   * read back on the host; n_groups == 0 disables the staged kernels */
  const int32_t* groups;
  int32_t n_groups, group_max_rows, group_max_entries, reserved2;
  /* per-row records of the staged kernels: dcgc_mg_record_bytes(n_atoms) device bytes filled by dcgc_mg_prepare
   * (NULL: staged kernels off) */
  const void* mg_records;
} dcgc_topology;

typedef struct dcgc_gcmodel_config {
  int32_t n_layers;
  int32_t n_feat;                        /* atom feature width (75) */
  int32_t widths[DCGC_MODEL_MAX_LAYERS]; /* graph_conv_layers */
  int32_t dense;                         /* dense_layer_size */
  int32_t n_out;                         /* n_tasks (regression) or n_tasks * n_classes */
  int32_t n_classes;
  int32_t mode;                          /* 0 regression (L2), 1 classification (softmax CE) */
  int32_t batch_norm;
  int32_t gemm_mode;                     /* DCGC_GEMM_* */
  float bn_eps;                          /* 1e-3 */
  float bn_momentum;                     /* 0.99: weight of the NEW statistic (torch convention) */
  int32_t input_exact;                   /* set per call: every entry of x is an integer with |neighbour sum| < 2048 (true of
                                            every ConvMol feature: one-hots, formal charge, radical electrons), so the first
                                            layer's operands are exact in tf32 and DCGC_GEMM_TF32X3 skips the identically-zero
                                            lo(A) * hi(W) term in its forward and weight-gradient GEMMs (same results) */
  int32_t reserved;
} dcgc_gcmodel_config;

/* --------------------------------------------------------------------------------------------
 * Molecule-group staged variants of K1/K5, K3 and K7 (csrc/molgroup_kernels.cu): one CTA per molecule group
 * of the layout slab stages the group's atom rows in shared memory with one bulk (TMA) copy per degree bucket
 * and gathers from there, so every activation row is read from HBM/L2 once, as a contiguous burst.  Same
 * arithmetic and summation / tie order as dcgc_gather_sum_bucketed / dcgc_pool_fwd / dcgc_pool_bwd (bit
 * identical results).  dcgc_mg_supported: 1 if the topology carries a valid group table and rows of
 * ld_floats floats (plus, for the pool backward, ld_arg_bytes of argmax bytes) of its largest group fit in
 * shared memory (at least two pipeline stages); leading dimensions must be multiples of 4 floats / 16 bytes.  transposed != 0 walks the
 * transposed lists (t_src) and requires a symmetric adjacency.
 * ------------------------------------------------------------------------------------------ */
int dcgc_mg_supported(const dcgc_topology* topo, int64_t ld_floats, int64_t ld_arg_bytes);
/* Once per batch, after the slab upload: fill topo-independent-of-width per-row records (row id, degree, the
 * shared-memory slot of every neighbour inside the row's group; for the transposed lists also the slot of the
 * row in each neighbour's list) into records_dev (dcgc_mg_record_bytes(n_atoms) bytes, 16-byte aligned).  The
 * caller then stores the pointer in topo->mg_records.  topo->mg_records is not read by this call. */
int64_t dcgc_mg_record_bytes(int64_t n_atoms);
int dcgc_mg_prepare(const dcgc_topology* topo, void* records_dev, void* stream);
int dcgc_mg_gather_sum(const float* x_dev, int64_t ld_x, const dcgc_topology* topo, int32_t transposed,
                       int32_t width, const float* addend_dev, int64_t ld_add, float* out_dev, int64_t ld_out,
                       void* stream);
int dcgc_mg_pool_fwd(const float* x_dev, int64_t ld_x, const float* scale_dev, const float* shift_dev,
                     const dcgc_topology* topo, int32_t width, float* out_dev, int64_t ld_out, uint8_t* arg_dev,
                     int64_t ld_arg, void* stream);
int dcgc_mg_pool_bwd(const float* dy_dev, int64_t ld_dy, const uint8_t* arg_dev, int64_t ld_arg,
                     const float* scale_dev, const dcgc_topology* topo, int32_t width, float* dx_dev,
                     int64_t ld_dx, void* stream);

/* dcgc_mg_pool_bwd that also emits the stage-1 column sums of the BatchNorm backward over the rows it writes:
 * part[cta][0][c] = sum_r dx[r,c], part[cta][1][c] = sum_r dx[r,c] * y[r,c] (float64, *n_chunks rows = the kernel's
 * grid, <= number of SMs), y = the BatchNorm input of the layer, mean = its batch column means (the products are
 * accumulated centred on them).  Replaces a separate pass over dx and y (Keras BatchNormalization backward inside
 * GraphConvModel, graph_models.py:862-902). */
int dcgc_mg_pool_bwd_stats(const float* dy_dev, int64_t ld_dy, const uint8_t* arg_dev, int64_t ld_arg,
                           const dcgc_topology* topo, int32_t width, float* dx_dev, int64_t ld_dx, const float* y_dev,
                           int64_t ld_y, const float* mean_dev, double* part_dev, int32_t* n_chunks, void* stream);

/* --------------------------------------------------------------------------------------------
 * MPNN edge-network message passing, forward and backward (the reference's torch port of these layers is forward-only;
 * the Keras originals, models/layers.py:3648-3887, train: the backward entry points are the gradients of the same
 * formulas).  The dense contractions go through dcgc_group_gemm_*; these are the parts that are not GEMMs.
 * ------------------------------------------------------------------------------------------ */

/* EdgeNetwork (torch_models/layers.py:4006-4088), first half of its bilinear factorisation: for destination atom
 * i with pairs q in [pair_ptr[i], pair_ptr[i+1])  (pairs grouped by atom_to_pair[:, 0], in pair order):
 *   z[i, f*h + b] = sum_q pf[pair_id[q], f] * x[pair_src[q], b]   (f < n_pf),    z[i, n_pf*h + b] = sum_q x[pair_src[q], b]
 * The message is then  z . W_ext  with  W_ext[f*h + b, a] = W[f, a*h + b],  W_ext[n_pf*h + b, a] = bias[a*h + b]
 * (one dense [n, (n_pf+1) h] x [(n_pf+1) h, h] GEMM instead of an h x h matrix per pair).  n_pf <= 32. */
int dcgc_pair_contract_fwd(const float* x_dev, int64_t ld_x, const float* pair_feat_dev, int64_t ld_pf,
                           const int32_t* pair_ptr_dev, const int32_t* pair_id_dev, const int32_t* pair_src_dev,
                           int64_t n_dst, int32_t n_pf, int32_t h, float* z_dev, int64_t ld_z, void* stream);
/* GatedRecurrentUnit (layers.py:2884-2919) around two GEMMs: g = [x | h_prev] . [[Wz Wr Wh]; [Uz Ur 0]] ([n, 3h]);
 *   gates:  z = sigmoid(g[:, 0:h] + bz),  r = sigmoid(g[:, h:2h] + br),  hr = h_prev * r
 *   out:    (1 - z) * tanh(g[:, 2h:3h] + u + bh) + z * x      with u = hr . Uh  (x is the message, as the reference writes it) */
int dcgc_gru_gates_fwd(const float* g_dev, int64_t ld_g, const float* bz_dev, const float* br_dev,
                       const float* hprev_dev, int64_t ld_h, int64_t n, int32_t h, float* z_dev, int64_t ld_z,
                       float* hr_dev, int64_t ld_hr, void* stream);
int dcgc_gru_out_fwd(const float* g_dev, int64_t ld_g, const float* u_dev, int64_t ld_u, const float* bh_dev,
                     const float* z_dev, int64_t ld_z, const float* x_dev, int64_t ld_x, int64_t n, int32_t h,
                     float* out_dev, int64_t ld_out, void* stream);
/* SetGather (layers.py:2976-3138), one set2set step: per molecule g with atoms mol_atoms[mol_ptr[g] .. mol_ptr[g+1])
 * (ascending): e_i = <x_i, q_g>, a = softmax(e), qstar[g] = [q_g | sum_i a_i x_i]  (zeros for an empty molecule);
 * then z = qstar . U + b (a GEMM) and the LSTM cell  i|f|o = sigmoid(z[0:3h]), c' = f c + i tanh(z[3h:4h]),
 * h' = o tanh(c').  max_atoms = the largest molecule of the batch (sizes the shared-memory scratch). */
int dcgc_setgather_attend_fwd(const float* x_dev, int64_t ld_x, const float* q_dev, int64_t ld_q,
                              const int32_t* mol_ptr_dev, const int32_t* mol_atoms_dev, int64_t n_mols, int32_t h,
                              int32_t max_atoms, float* qstar_dev, int64_t ld_qs, void* stream);
int dcgc_lstm_step_fwd(const float* z_dev, int64_t ld_z, const float* c_in_dev, int64_t n, int32_t h,
                       float* h_out_dev, float* c_out_dev, void* stream);

/* Backward of dcgc_pair_contract_fwd with respect to x (the pair features are data): for source atom j with the pairs
 * src_pair[src_ptr[j] .. src_ptr[j+1]) that read it (pair ids in ascending order: deterministic, no atomics),
 *   dx[j, b] = sum_q ( sum_f pf[p, f] * dz[pair_dst[p], f*h + b] + dz[pair_dst[p], n_pf*h + b] ),   p = src_pair[q]. */
int dcgc_pair_contract_bwd_x(const float* dz_dev, int64_t ld_z, const float* pair_feat_dev, int64_t ld_pf,
                             const int32_t* src_ptr_dev, const int32_t* src_pair_dev, const int32_t* pair_dst_dev,
                             int64_t n_src, int32_t n_pf, int32_t h, float* dx_dev, int64_t ld_dx, void* stream);
/* Backward of dcgc_gru_out_fwd: dzg = d out (x - cand), dx = d out z, dpre = d out (1 - z)(1 - cand^2) with
 * cand = tanh(g[:, 2h:3h] + u + bh) recomputed (dpre is the gradient of g[:, 2h:3h], of u and, summed over rows, of bh). */
int dcgc_gru_out_bwd(const float* dout_dev, int64_t ld_do, const float* g_dev, int64_t ld_g, const float* u_dev,
                     int64_t ld_u, const float* bh_dev, const float* z_dev, int64_t ld_z, const float* x_dev,
                     int64_t ld_x, int64_t n, int32_t h, float* dzg_dev, int64_t ld_dz, float* dx_dev, int64_t ld_dx,
                     float* dpre_dev, int64_t ld_dp, void* stream);
/* Backward of dcgc_gru_gates_fwd: dg = [dzg z (1 - z) | dhr h_prev r (1 - r) | dpre] ([n, 3h], the gradient of g and,
 * summed over rows, of bz / br / -), dh = dhr * r (the direct part of d h_prev). */
int dcgc_gru_gates_bwd(const float* dzg_dev, int64_t ld_dz, const float* dhr_dev, int64_t ld_dhr, const float* dpre_dev,
                       int64_t ld_dp, const float* g_dev, int64_t ld_g, const float* bz_dev, const float* br_dev,
                       const float* hprev_dev, int64_t ld_h, int64_t n, int32_t h, float* dg_dev, int64_t ld_dg,
                       float* dh_dev, int64_t ld_dh, void* stream);
/* Backward of dcgc_setgather_attend_fwd: given d qstar [n_mols, 2h] writes dx (every atom row once) and dq. */
int dcgc_setgather_attend_bwd(const float* x_dev, int64_t ld_x, const float* q_dev, int64_t ld_q, const float* dqs_dev,
                              int64_t ld_dqs, const int32_t* mol_ptr_dev, const int32_t* mol_atoms_dev, int64_t n_mols,
                              int32_t h, int32_t max_atoms, float* dx_dev, int64_t ld_dx, float* dq_dev, int64_t ld_dq,
                              void* stream);
/* Backward of dcgc_lstm_step_fwd: dh / dc_out may be null (zero); writes dz [n, 4h] and dc_in [n, h]. */
int dcgc_lstm_step_bwd(const float* z_dev, int64_t ld_z, const float* c_in_dev, const float* dh_dev,
                       const float* dc_out_dev, int64_t n, int32_t h, float* dz_dev, int64_t ld_dz, float* dc_in_dev,
                       void* stream);

/* param_offsets: 4 per conv layer (W, b, gamma, beta), then dense (W, b, gamma, beta), then head
 * (W, b); -1 where batch_norm is off.  bn_offsets: (mean, var) per BN, conv layers then dense. */
/* The fused engine runs the FORWARD GEMMs of the TF32x3 mode with fp16 operand halves (same 11 + 11 significand bits,
 * half the tensor-core instructions).  Both operands are multiplied by 16 before the split (undone exactly in the
 * epilogue) so that the low halves of values down to 2^-7 stay normal fp16 numbers; fp16 overflows above 65 504, i.e.
 * for operands above 4094.  Returns 1 if a forward GEMM has seen an activation or weight above 3750 since the library
 * was loaded (its results are then not to be trusted: set DCGC_FWD_F16X3=0), 0 if not,
 * -1 on a CUDA error.  Synchronises the device. */
int dcgc_tc_f16_overflow(void);

int dcgc_gcmodel_layout(const dcgc_gcmodel_config* cfg, int64_t* param_offsets, int64_t* bn_offsets,
                        int64_t* n_params, int64_t* n_bn);
int64_t dcgc_gcmodel_workspace_bytes(const dcgc_gcmodel_config* cfg, int64_t n_atoms,
                                     int64_t n_segments);
/* Inference / plain forward.  training != 0: batch statistics (running stats updated if
 * bn_running != NULL); training == 0: running statistics.  out [n_samples, n_out] (logits in
 * classification mode), probs (optional, classification), fingerprint [n_segments, 2D] (optional,
 * untrimmed as in the reference). */
int dcgc_gcmodel_forward(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x_dev,
                         int64_t ld_x, int64_t n_samples, const float* params_dev, float* bn_running_dev,
                         int32_t training, void* workspace_dev, int64_t workspace_bytes, float* out_dev,
                         float* probs_dev, float* fingerprint_dev, void* stream);
/* One training forward + loss + backward.  y: [n_samples, n_tasks] (regression) or one-hot
 * [n_samples, n_tasks, n_classes]; w: [n_samples, n_tasks] (may be NULL = ones).  loss_dev receives
 * the scalar loss (mean over n_samples * n_tasks elements); every gradient is written (not
 * accumulated) into grads_dev. */
int dcgc_gcmodel_train_step(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x_dev,
                            int64_t ld_x, const float* y_dev, const float* w_dev, int64_t n_samples,
                            const float* params_dev, float* grads_dev, float* bn_running_dev,
                            void* workspace_dev, int64_t workspace_bytes, float* loss_dev,
                            float* out_dev, void* stream);
/* The same step with explicit event hooks (cudaEvent_t handles, every one optional) recorded on `stream`:
 *   forward_event      between the forward and the backward pass (the host pipeline can start the upload of a later
 *                      batch beside the GEMM-heavy backward instead of beside the HBM-bound forward kernels);
 *   grad_events[i]     when gradient slice i of grads_dev is final, i < n_grad_events <= n_layers + 1:
 *                      slice 0 = [dense, its BatchNorm, head] = offsets [dense_w, n_params),
 *                      slice 1 + k = graph-conv layer n_layers - 1 - k with its BatchNorm = [conv_w[l], conv_w[l+1])
 *                      (conv_w[n_layers] := dense_w; offsets from dcgc_gcmodel_layout).
 * A data-parallel caller waits for each event on a communication stream and all-reduces that slice while the rest of
 * the backward pass runs ("issue per-layer as soon as that layer's wgrad is written"). */
int dcgc_gcmodel_train_step_ev(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x_dev,
                               int64_t ld_x, const float* y_dev, const float* w_dev, int64_t n_samples,
                               const float* params_dev, float* grads_dev, float* bn_running_dev,
                               void* workspace_dev, int64_t workspace_bytes, float* loss_dev, float* out_dev,
                               void* forward_event, void* const* grad_events, int32_t n_grad_events, void* stream);

/* ---- synchronised BatchNorm inside the fused step (SURVEY 8e opt-in): statistics over the atoms of every rank ----------
 * The ranks of one node exchange their per-column sums through PEER MEMORY: every rank owns a small mailbox
 * (dcgc_p2p_alloc: device memory with a CUDA IPC handle) that its peers map (dcgc_p2p_open) and write with plain stores
 * over NVLink from inside the BatchNorm finalize kernel — sums + row count, a system-scope fence, then a sequence flag;
 * the kernel then waits for the flags of all ranks in ITS OWN mailbox and adds the contributions in rank order, so
 * every rank obtains bit-identical global statistics without a library collective or a host round trip (two kernels
 * of one block per BatchNorm and direction, a few microseconds each).  Mailbox layout (doubles):
 *   [sender rank][slot = use & 1][2 * cap sums | row count | flag]   with cap >= the widest BatchNorm of the model;
 * bytes = dcgc_bn_sync_mailbox_bytes(world, cap).  `seq` numbers the BatchNorm uses since the mailboxes were zeroed:
 * the step consumes 2 * (n_layers + 1) numbers starting at seq0 (the caller advances it), every rank must pass the same. */
#define DCGC_SYNC_MAX_RANKS 8
typedef struct dcgc_bn_sync {
  int32_t world, rank;
  int32_t cap;                 /* columns a mailbox row holds */
  int32_t reserved;
  uint64_t seq0;               /* first sequence number of this step (> 0) */
  void* mailbox[DCGC_SYNC_MAX_RANKS];   /* mailbox of rank r as mapped in THIS process (own: the dcgc_p2p_alloc pointer) */
} dcgc_bn_sync;
int64_t dcgc_bn_sync_mailbox_bytes(int32_t world, int32_t cap);
/* device memory shareable between the processes of one node: handle_out = 64 bytes (cudaIpcMemHandle_t), zero-filled */
int dcgc_p2p_alloc(int64_t bytes, void** ptr_out, void* handle_out);
int dcgc_p2p_open(const void* handle, void** ptr_out);
int dcgc_p2p_close(void* ptr);
int dcgc_p2p_free(void* ptr);
/* dcgc_gcmodel_train_step_ev with synchronised BatchNorm (sync may be NULL or world == 1: identical to _ev).  The
 * gradients of gamma / beta stay LOCAL sums (the caller's gradient exchange adds them over ranks, as for every other
 * parameter); tensor-core GEMM modes only. */
int dcgc_gcmodel_train_step_sync(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x_dev,
                                 int64_t ld_x, const float* y_dev, const float* w_dev, int64_t n_samples,
                                 const float* params_dev, float* grads_dev, float* bn_running_dev,
                                 void* workspace_dev, int64_t workspace_bytes, float* loss_dev, float* out_dev,
                                 void* forward_event, void* const* grad_events, int32_t n_grad_events,
                                 const dcgc_bn_sync* sync, void* stream);
/* Fused Adam over a flat slab, torch.optim.Adam semantics (models/optimizers.py:190-241);
 * grads are multiplied by grad_scale first (1/world_size after a summing all-reduce). */
int dcgc_adam_step(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                   int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                   float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCGC_H_ */

// Device helpers of the D-MPNN path that are not gathers or GEMMs (those reuse dcgc_gather_sum and the
// dcgc_linear_* / dcgc_group_gemm_* entry points): the f_ini row assembly and the per-molecule readout.
#include "common.h"

namespace {

constexpr int kT = 256;

// out[r] = [atom_feat[bond_src[r]] | bond_feat[bond_edge[r]] | 0 pad]; one thread per output element,
// consecutive threads on consecutive columns of a row (coalesced stores, coalesced row reads)
__global__ void __launch_bounds__(kT)
concat_rows_kernel(const float* __restrict__ af, int64_t ld_a, int fa, const float* __restrict__ bf, int64_t ld_b,
                   int fb, const int32_t* __restrict__ bond_src, const int32_t* __restrict__ bond_edge,
                   int64_t n_rows, float* __restrict__ out, int64_t ld_out) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kT + threadIdx.x;
  const int64_t r = t / ld_out;
  const int c = (int)(t - r * ld_out);
  if (r >= n_rows) return;
  const int s = __ldg(bond_src + r);
  float v = 0.f;
  if (s >= 0) {
    if (c < fa) v = __ldg(af + (int64_t)s * ld_a + c);
    else if (c < fa + fb) v = __ldg(bf + (int64_t)__ldg(bond_edge + r) * ld_b + (c - fa));
  }
  out[r * ld_out + c] = v;
}

// one thread = one column of one molecule, rows added in ascending order (torch's sum over dim 0 of a
// contiguous [n, w] block adds rows in order for the small n of a molecule)
__global__ void __launch_bounds__(kT)
readout_fwd_kernel(const float* __restrict__ x, int64_t ld_x, const int32_t* __restrict__ mol_ptr, int64_t n_mols,
                   int width, int mode, float norm, float* __restrict__ out, int64_t ld_out) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kT + threadIdx.x;
  const int64_t m = t / width;
  const int c = (int)(t - m * width);
  if (m >= n_mols) return;
  const int a0 = __ldg(mol_ptr + m), a1 = __ldg(mol_ptr + m + 1);
  float s = 0.f;
  for (int a = a0; a < a1; ++a) s += __ldg(x + (int64_t)a * ld_x + c);
  // mean is a true division in the reference (sum / len), not a multiplication by the reciprocal
  if (mode == 0) s = s / (float)(a1 - a0);
  else if (mode == 2) s = s / norm;
  out[m * ld_out + c] = s;
}

__global__ void __launch_bounds__(kT)
readout_bwd_kernel(const float* __restrict__ dout, int64_t ld_dout, const int32_t* __restrict__ mol_ptr,
                   int64_t n_mols, int width, int mode, float norm, float* __restrict__ dx, int64_t ld_dx) {
  dcgc_griddep_wait();
  // one block row per molecule: blockIdx.x = molecule, threads stride over (atom, column)
  const int64_t m = blockIdx.x;
  const int a0 = __ldg(mol_ptr + m), a1 = __ldg(mol_ptr + m + 1);
  const int n = a1 - a0;
  for (int i = threadIdx.x; i < n * width; i += kT) {
    const int a = i / width, c = i - a * width;
    float g = __ldg(dout + m * ld_dout + c);
    if (mode == 0) g = g / (float)n;
    else if (mode == 2) g = g / norm;
    dx[(int64_t)(a0 + a) * ld_dx + c] = g;
  }
}

}  // namespace

extern "C" int dcgc_dmpnn_concat_rows(const float* af, int64_t ld_a, int32_t fa, const float* bf, int64_t ld_b,
                                      int32_t fb, const int32_t* bond_src, const int32_t* bond_edge, int64_t n_rows,
                                      float* out, int64_t ld_out, void* stream) {
  DCGC_CHECK_ARG(n_rows >= 0 && fa >= 0 && fb >= 0 && ld_a >= fa && ld_b >= fb && ld_out >= fa + fb,
                 "dcgc_dmpnn_concat_rows: bad sizes");
  if (n_rows == 0 || ld_out == 0) return DCGC_OK;
  DCGC_CHECK_ARG(bond_src && bond_edge && out && (af || fa == 0) && (bf || fb == 0),
                 "dcgc_dmpnn_concat_rows: null pointer");
  DcgcProfScope prof_scope("dcgc_dmpnn_concat_rows", (cudaStream_t)stream);
  const int64_t work = n_rows * ld_out;
  dcgc_launch(concat_rows_kernel, (unsigned)((work + kT - 1) / kT), kT, 0, (cudaStream_t)stream, 
      af, ld_a, fa, bf, ld_b, fb, bond_src, bond_edge, n_rows, out, ld_out);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_dmpnn_concat_rows");
  return DCGC_OK;
}

extern "C" int dcgc_segment_readout_fwd(const float* x, int64_t ld_x, const int32_t* mol_ptr, int64_t n_mols,
                                        int32_t width, int32_t mode, float norm, float* out, int64_t ld_out,
                                        void* stream) {
  DCGC_CHECK_ARG(n_mols >= 0 && width >= 0 && ld_x >= width && ld_out >= width, "dcgc_segment_readout_fwd: bad sizes");
  DCGC_CHECK_ARG(mode >= 0 && mode <= 2, "dcgc_segment_readout_fwd: Invalid aggregation");
  if (n_mols == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && mol_ptr && out, "dcgc_segment_readout_fwd: null pointer");
  DcgcProfScope prof_scope("dcgc_segment_readout_fwd", (cudaStream_t)stream);
  const int64_t work = n_mols * width;
  dcgc_launch(readout_fwd_kernel, (unsigned)((work + kT - 1) / kT), kT, 0, (cudaStream_t)stream, x, ld_x, mol_ptr, n_mols,
                                                                                       width, mode, norm, out, ld_out);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_segment_readout_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_segment_readout_bwd(const float* dout, int64_t ld_dout, const int32_t* mol_ptr, int64_t n_mols,
                                        int64_t n_atoms, int32_t width, int32_t mode, float norm, float* dx,
                                        int64_t ld_dx, void* stream) {
  DCGC_CHECK_ARG(n_mols >= 0 && n_atoms >= 0 && width >= 0 && ld_dout >= width && ld_dx >= width,
                 "dcgc_segment_readout_bwd: bad sizes");
  DCGC_CHECK_ARG(mode >= 0 && mode <= 2, "dcgc_segment_readout_bwd: Invalid aggregation");
  if (n_mols == 0 || width == 0 || n_atoms == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dout && mol_ptr && dx, "dcgc_segment_readout_bwd: null pointer");
  DcgcProfScope prof_scope("dcgc_segment_readout_bwd", (cudaStream_t)stream);
  dcgc_launch(readout_bwd_kernel, (unsigned)n_mols, kT, 0, (cudaStream_t)stream, dout, ld_dout, mol_ptr, n_mols, width, mode,
                                                                        norm, dx, ld_dx);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_segment_readout_bwd");
  return DCGC_OK;
}

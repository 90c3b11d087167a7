// Device kernels of the MPNN edge-network message passing, forward and backward.  (The reference's torch port of these
// layers is forward-only — plain-tensor weights, torch_models/layers.py:2884-3138, 4006-4088 — the Keras originals,
// models/layers.py:3648-3887, train; the backward kernels below are the gradients of the SAME forward formulas.)  The
// dense contractions run through dcgc_group_gemm_* (tcgen05); what is here is the part that is not a GEMM.
//
// EdgeNetwork.  The reference maps every atom PAIR p = (i, j) to an h x h matrix A_p = reshape(pf_p . W + b) and
// sums A_p . x_j over the pairs of destination atom i: 2 (P + 1) h^2 flops and h^2 floats of intermediate per pair
// (25 MB per 25-atom molecule at h = 100).  The map is bilinear in (pf_p, x_j), so
//     m_i[a] = sum_{f,b} W[f, a h + b] * Z_i[f, b]  +  sum_b bias[a h + b] * Z_i[P, b],
//     Z_i[f, b] = sum_{p -> i} pf_p[f] * x_{j(p)}[b],     Z_i[P, b] = sum_{p -> i} x_{j(p)}[b]:
// a small per-destination contraction (pair_contract_kernel, 2 (P + 1) h flops per pair) followed by ONE dense GEMM
// [n_atoms, (P + 1) h] x [(P + 1) h, h] on the tensor cores — ~h / 2 times fewer flops and no per-pair matrix.
#include "common.h"

namespace {

constexpr int kT = 128;

// grid = destination atoms, block = 128 threads over the hidden columns; pairs of a destination are visited in
// pair order (the order of the reference's sorted segment sum)
template <int PMAX>
__global__ void __launch_bounds__(kT)
pair_contract_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ pf, int64_t ld_pf,
                     const int32_t* __restrict__ pair_ptr, const int32_t* __restrict__ pair_id,
                     const int32_t* __restrict__ pair_src, int n_pf, int h, float* __restrict__ z, int64_t ld_z) {
  dcgc_griddep_wait();
  const int i = blockIdx.x;
  const int q0 = __ldg(pair_ptr + i), q1 = __ldg(pair_ptr + i + 1);
  for (int b = threadIdx.x; b < h; b += kT) {
    float acc[PMAX + 1];
#pragma unroll
    for (int f = 0; f <= PMAX; ++f) acc[f] = 0.f;
    for (int q = q0; q < q1; ++q) {
      const float xv = __ldg(x + (int64_t)__ldg(pair_src + q) * ld_x + b);
      const float* pr = pf + (int64_t)__ldg(pair_id + q) * ld_pf;
#pragma unroll
      for (int f = 0; f < PMAX; ++f)
        if (f < n_pf) acc[f] = fmaf(__ldg(pr + f), xv, acc[f]);
      acc[PMAX] += xv;
    }
    float* zi = z + (int64_t)i * ld_z + b;
#pragma unroll
    for (int f = 0; f < PMAX; ++f)
      if (f < n_pf) zi[(int64_t)f * h] = acc[f];
    zi[(int64_t)n_pf * h] = acc[PMAX];
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// g = [x | h_prev] . [[Wz Wr Wh]; [Uz Ur 0]]  (one two-operand GEMM, [n, 3h]):
//   z = sigmoid(g[:, 0:h] + bz), r = sigmoid(g[:, h:2h] + br), hr = h_prev * r
__global__ void __launch_bounds__(256)
gru_gates_kernel(const float* __restrict__ g, int64_t ld_g, const float* __restrict__ bz, const float* __restrict__ br,
                 const float* __restrict__ hprev, int64_t ld_h, int64_t n, int h, float* __restrict__ z, int64_t ld_z,
                 float* __restrict__ hr, int64_t ld_hr) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r_ = t / h;
  const int c = (int)(t - r_ * h);
  if (r_ >= n) return;
  const float zv = sigmoidf_(__ldg(g + r_ * ld_g + c) + __ldg(bz + c));
  const float rv = sigmoidf_(__ldg(g + r_ * ld_g + h + c) + __ldg(br + c));
  z[r_ * ld_z + c] = zv;
  hr[r_ * ld_hr + c] = __ldg(hprev + r_ * ld_h + c) * rv;
}

// out = (1 - z) * tanh(g[:, 2h:3h] + u + bh) + z * x        (u = hr . Uh; x = the message, layers.py:2916-2918)
__global__ void __launch_bounds__(256)
gru_out_kernel(const float* __restrict__ g, int64_t ld_g, const float* __restrict__ u, int64_t ld_u,
               const float* __restrict__ bh, const float* __restrict__ z, int64_t ld_z, const float* __restrict__ x,
               int64_t ld_x, int64_t n, int h, float* __restrict__ out, int64_t ld_out) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r_ = t / h;
  const int c = (int)(t - r_ * h);
  if (r_ >= n) return;
  const float zv = __ldg(z + r_ * ld_z + c);
  const float cand = tanhf(__ldg(g + r_ * ld_g + 2 * h + c) + __ldg(u + r_ * ld_u + c) + __ldg(bh + c));
  out[r_ * ld_out + c] = (1.f - zv) * cand + zv * __ldg(x + r_ * ld_x + c);
}

// set2set attention of one molecule per block (layers.py:3056-3076): e_i = <x_i, q_g>, a = softmax over the atoms
// of the molecule, r_g = sum_i a_i x_i (atoms in ascending order), q_star[g] = [q_g | r_g].
// dynamic shared memory: max_atoms floats
__global__ void __launch_bounds__(kT)
setgather_attend_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ q, int64_t ld_q,
                        const int32_t* __restrict__ mol_ptr, const int32_t* __restrict__ mol_atoms, int h,
                        float* __restrict__ qstar, int64_t ld_qs) {
  dcgc_griddep_wait();
  extern __shared__ float e_sh[];
  __shared__ float red[2];
  const int g = blockIdx.x;
  const int t0 = __ldg(mol_ptr + g), t1 = __ldg(mol_ptr + g + 1);
  const int n = t1 - t0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qg = q + (int64_t)g * ld_q;
  for (int a = warp; a < n; a += kT / 32) {
    const float* xr = x + (int64_t)__ldg(mol_atoms + t0 + a) * ld_x;
    float s = 0.f;
    for (int c = lane; c < h; c += 32) s = fmaf(__ldg(xr + c), __ldg(qg + c), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) e_sh[a] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int a = lane; a < n; a += 32) m = fmaxf(m, e_sh[a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int a = lane; a < n; a += 32) {
      const float ev = expf(e_sh[a] - m);
      e_sh[a] = ev;
      s += ev;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[0] = s;
  }
  __syncthreads();
  const float inv_den = n > 0 ? red[0] : 1.f;
  for (int c = threadIdx.x; c < h; c += kT) {
    float r = 0.f;
    for (int a = 0; a < n; ++a)
      r = fmaf(e_sh[a] / inv_den, __ldg(x + (int64_t)__ldg(mol_atoms + t0 + a) * ld_x + c), r);
    qstar[(int64_t)g * ld_qs + c] = __ldg(qg + c);
    qstar[(int64_t)g * ld_qs + h + c] = r;
  }
}

// LSTM cell on z = q_star . U + b  ([n, 4h]: i | f | o | candidate), layers.py:3100-3108
__global__ void __launch_bounds__(256)
lstm_step_kernel(const float* __restrict__ zg, int64_t ld_z, const float* __restrict__ c_in, int64_t n, int h,
                 float* __restrict__ h_out, float* __restrict__ c_out) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r_ = t / h;
  const int c = (int)(t - r_ * h);
  if (r_ >= n) return;
  const float* zr = zg + r_ * ld_z;
  const float iv = sigmoidf_(__ldg(zr + c)), fv = sigmoidf_(__ldg(zr + h + c)), ov = sigmoidf_(__ldg(zr + 2 * h + c));
  const float cn = fv * __ldg(c_in + r_ * h + c) + iv * tanhf(__ldg(zr + 3 * h + c));
  c_out[r_ * h + c] = cn;
  h_out[r_ * h + c] = ov * tanhf(cn);
}

// ------------------------------------------------------------------------------------------
// backward kernels
// ------------------------------------------------------------------------------------------
// d x_j[b] = sum over the pairs p with source j (in pair order: deterministic, no atomics) of
//            sum_f pf_p[f] * dZ_{dst(p)}[f, b] + dZ_{dst(p)}[P, b]
// grid = source atoms (pairs grouped by source: src_ptr / src_pair), block over the hidden columns.
template <int PMAX>
__global__ void __launch_bounds__(kT)
pair_contract_bwd_x_kernel(const float* __restrict__ dz, int64_t ld_z, const float* __restrict__ pf, int64_t ld_pf,
                           const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_pair,
                           const int32_t* __restrict__ pair_dst, int n_pf, int h, float* __restrict__ dx, int64_t ld_dx) {
  dcgc_griddep_wait();
  const int j = blockIdx.x;
  const int q0 = __ldg(src_ptr + j), q1 = __ldg(src_ptr + j + 1);
  for (int b = threadIdx.x; b < h; b += kT) {
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) {
      const int pid = __ldg(src_pair + q);
      const float* zr = dz + (int64_t)__ldg(pair_dst + pid) * ld_z + b;
      const float* pr = pf + (int64_t)pid * ld_pf;
      float v[PMAX + 1];
#pragma unroll
      for (int f = 0; f < PMAX; ++f) v[f] = f < n_pf ? __ldg(zr + (int64_t)f * h) : 0.f;    // independent loads
      v[PMAX] = __ldg(zr + (int64_t)n_pf * h);
#pragma unroll
      for (int f = 0; f < PMAX; ++f)
        if (f < n_pf) acc = fmaf(__ldg(pr + f), v[f], acc);
      acc += v[PMAX];
    }
    dx[(int64_t)j * ld_dx + b] = acc;
  }
}

// GRU backward, elementwise part 2 (the output equation): out = (1 - z) cand + z x, cand = tanh(g2 + u + bh)
//   d z = d out (x - cand);  d x = d out z;  d pre = d out (1 - z)(1 - cand^2)  (= d g2 = d u = d bh rows)
__global__ void __launch_bounds__(256)
gru_out_bwd_kernel(const float* __restrict__ dout, int64_t ld_do, const float* __restrict__ g, int64_t ld_g,
                   const float* __restrict__ u, int64_t ld_u, const float* __restrict__ bh, const float* __restrict__ z,
                   int64_t ld_z, const float* __restrict__ x, int64_t ld_x, int64_t n, int h, float* __restrict__ dzg,
                   int64_t ld_dz, float* __restrict__ dx, int64_t ld_dx, float* __restrict__ dpre, int64_t ld_dp) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r_ = t / h;
  const int c = (int)(t - r_ * h);
  if (r_ >= n) return;
  const float d = __ldg(dout + r_ * ld_do + c), zv = __ldg(z + r_ * ld_z + c), xv = __ldg(x + r_ * ld_x + c);
  const float cand = tanhf(__ldg(g + r_ * ld_g + 2 * h + c) + __ldg(u + r_ * ld_u + c) + __ldg(bh + c));
  dzg[r_ * ld_dz + c] = d * (xv - cand);
  dx[r_ * ld_dx + c] = d * zv;
  dpre[r_ * ld_dp + c] = d * (1.f - zv) * (1.f - cand * cand);
}

// GRU backward, elementwise part 1 (the gates): z = sigmoid(g0 + bz), r = sigmoid(g1 + br), hr = h_prev r.
//   dg[:, 0:h] = d z * z (1 - z);  dg[:, h:2h] = d hr * h_prev * r (1 - r);  dg[:, 2h:3h] = d pre;
//   d h_prev (direct part) = d hr * r
__global__ void __launch_bounds__(256)
gru_gates_bwd_kernel(const float* __restrict__ dzg, int64_t ld_dz, const float* __restrict__ dhr, int64_t ld_dhr,
                     const float* __restrict__ dpre, int64_t ld_dp, const float* __restrict__ g, int64_t ld_g,
                     const float* __restrict__ bz, const float* __restrict__ br, const float* __restrict__ hprev,
                     int64_t ld_h, int64_t n, int h, float* __restrict__ dg, int64_t ld_dg, float* __restrict__ dh,
                     int64_t ld_dh) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r_ = t / h;
  const int c = (int)(t - r_ * h);
  if (r_ >= n) return;
  const float zv = sigmoidf_(__ldg(g + r_ * ld_g + c) + __ldg(bz + c));
  const float rv = sigmoidf_(__ldg(g + r_ * ld_g + h + c) + __ldg(br + c));
  const float dh_r = __ldg(dhr + r_ * ld_dhr + c);
  dg[r_ * ld_dg + c] = __ldg(dzg + r_ * ld_dz + c) * zv * (1.f - zv);
  dg[r_ * ld_dg + h + c] = dh_r * __ldg(hprev + r_ * ld_h + c) * rv * (1.f - rv);
  dg[r_ * ld_dg + 2 * h + c] = __ldg(dpre + r_ * ld_dp + c);
  dh[r_ * ld_dh + c] = dh_r * rv;
}

// set2set attention backward of one molecule per block.  Forward: e_i = <x_i, q>, a = softmax(e), r = sum a_i x_i,
// q_star = [q | r].  Given d q_star = [dq0 | dr]:
//   da_i = <dr, x_i>;  de_i = a_i (da_i - sum_j a_j da_j);  dx_i = a_i dr + de_i q;  dq = dq0 + sum_i de_i x_i
// (atoms in ascending order; dx rows are written, not accumulated: every atom belongs to one molecule).
// dynamic shared memory: 2 * max_atoms floats (a, de)
__global__ void __launch_bounds__(kT)
setgather_attend_bwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ q, int64_t ld_q,
                            const float* __restrict__ dqs, int64_t ld_dqs, const int32_t* __restrict__ mol_ptr,
                            const int32_t* __restrict__ mol_atoms, int h, int max_atoms, float* __restrict__ dx,
                            int64_t ld_dx, float* __restrict__ dq, int64_t ld_dq) {
  dcgc_griddep_wait();
  extern __shared__ float sh[];
  float* a_sh = sh;
  float* de_sh = sh + max_atoms;
  const int g = blockIdx.x;
  const int t0 = __ldg(mol_ptr + g), t1 = __ldg(mol_ptr + g + 1);
  const int n = t1 - t0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qg = q + (int64_t)g * ld_q;
  const float* dr = dqs + (int64_t)g * ld_dqs + h;
  // e_i and da_i, one warp per atom
  for (int a = warp; a < n; a += kT / 32) {
    const float* xr = x + (int64_t)__ldg(mol_atoms + t0 + a) * ld_x;
    float s = 0.f, d = 0.f;
    for (int c = lane; c < h; c += 32) {
      const float xv = __ldg(xr + c);
      s = fmaf(xv, __ldg(qg + c), s);
      d = fmaf(xv, __ldg(dr + c), d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
    if (lane == 0) { a_sh[a] = s; de_sh[a] = d; }
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int a = lane; a < n; a += 32) m = fmaxf(m, a_sh[a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int a = lane; a < n; a += 32) { const float ev = expf(a_sh[a] - m); a_sh[a] = ev; s += ev; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float dot = 0.f;                       // sum_j a_j da_j
    for (int a = lane; a < n; a += 32) { const float av = a_sh[a] / s; a_sh[a] = av; dot = fmaf(av, de_sh[a], dot); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int a = lane; a < n; a += 32) de_sh[a] = a_sh[a] * (de_sh[a] - dot);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += kT) {
    const float qv = __ldg(qg + c), drv = __ldg(dr + c);
    float acc = __ldg(dqs + (int64_t)g * ld_dqs + c);
    for (int a = 0; a < n; ++a) {
      const int64_t row = __ldg(mol_atoms + t0 + a);
      dx[row * ld_dx + c] = fmaf(a_sh[a], drv, de_sh[a] * qv);
      acc = fmaf(de_sh[a], __ldg(x + row * ld_x + c), acc);
    }
    dq[(int64_t)g * ld_dq + c] = acc;
  }
}

// LSTM cell backward: h = o tanh(c'), c' = f c + i tanh(z3); gates from z = [i | f | o | z3] pre-activations
__global__ void __launch_bounds__(256)
lstm_step_bwd_kernel(const float* __restrict__ zg, int64_t ld_z, const float* __restrict__ c_in,
                     const float* __restrict__ dh, const float* __restrict__ dc_out, int64_t n, int h,
                     float* __restrict__ dz, int64_t ld_dz, float* __restrict__ dc_in) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r_ = t / h;
  const int c = (int)(t - r_ * h);
  if (r_ >= n) return;
  const float* zr = zg + r_ * ld_z;
  const float iv = sigmoidf_(__ldg(zr + c)), fv = sigmoidf_(__ldg(zr + h + c)), ov = sigmoidf_(__ldg(zr + 2 * h + c));
  const float gv = tanhf(__ldg(zr + 3 * h + c));
  const float cp = __ldg(c_in + r_ * h + c);
  const float cn = fv * cp + iv * gv;
  const float tc = tanhf(cn);
  const float dhv = dh ? __ldg(dh + r_ * h + c) : 0.f;
  const float dcn = (dc_out ? __ldg(dc_out + r_ * h + c) : 0.f) + dhv * ov * (1.f - tc * tc);
  float* dzr = dz + r_ * ld_dz;
  dzr[c] = dcn * gv * iv * (1.f - iv);
  dzr[h + c] = dcn * cp * fv * (1.f - fv);
  dzr[2 * h + c] = dhv * tc * ov * (1.f - ov);
  dzr[3 * h + c] = dcn * iv * (1.f - gv * gv);
  dc_in[r_ * h + c] = dcn * fv;
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int dcgc_pair_contract_fwd(const float* x, int64_t ld_x, const float* pf, int64_t ld_pf,
                                      const int32_t* pair_ptr, const int32_t* pair_id, const int32_t* pair_src,
                                      int64_t n_dst, int32_t n_pf, int32_t h, float* z, int64_t ld_z, void* stream) {
  DCGC_CHECK_ARG(n_dst >= 0 && n_pf >= 0 && h > 0 && ld_x >= h && ld_pf >= n_pf && ld_z >= (int64_t)(n_pf + 1) * h,
                 "dcgc_pair_contract_fwd: bad sizes");
  DCGC_CHECK_ARG(n_pf <= 32, "dcgc_pair_contract_fwd: at most 32 pair features are supported (got %d)", n_pf);
  if (n_dst == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && pair_ptr && z && (n_pf == 0 || pf), "dcgc_pair_contract_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DcgcProfScope prof_scope("dcgc_pair_contract_fwd", st);
  if (n_pf <= 8)
    dcgc_launch(pair_contract_kernel<8>, (unsigned)n_dst, kT, 0, st, x, ld_x, pf, ld_pf, pair_ptr, pair_id, pair_src, n_pf, h, z, ld_z);
  else if (n_pf <= 16)
    dcgc_launch(pair_contract_kernel<16>, (unsigned)n_dst, kT, 0, st, x, ld_x, pf, ld_pf, pair_ptr, pair_id, pair_src, n_pf, h, z, ld_z);
  else
    dcgc_launch(pair_contract_kernel<32>, (unsigned)n_dst, kT, 0, st, x, ld_x, pf, ld_pf, pair_ptr, pair_id, pair_src, n_pf, h, z, ld_z);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_pair_contract_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_gru_gates_fwd(const float* g, int64_t ld_g, const float* bz, const float* br, const float* hprev,
                                  int64_t ld_h, int64_t n, int32_t h, float* z, int64_t ld_z, float* hr, int64_t ld_hr,
                                  void* stream) {
  DCGC_CHECK_ARG(n >= 0 && h > 0 && ld_g >= 3 * (int64_t)h && ld_h >= h && ld_z >= h && ld_hr >= h,
                 "dcgc_gru_gates_fwd: bad sizes");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(g && bz && br && hprev && z && hr, "dcgc_gru_gates_fwd: null pointer");
  dcgc_launch(gru_gates_kernel, blocks_for(n * h, 256), 256, 0, (cudaStream_t)stream, g, ld_g, bz, br, hprev, ld_h, n, h, z, ld_z,
                                                                             hr, ld_hr);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gru_gates_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_gru_out_fwd(const float* g, int64_t ld_g, const float* u, int64_t ld_u, const float* bh,
                                const float* z, int64_t ld_z, const float* x, int64_t ld_x, int64_t n, int32_t h,
                                float* out, int64_t ld_out, void* stream) {
  DCGC_CHECK_ARG(n >= 0 && h > 0 && ld_g >= 3 * (int64_t)h && ld_u >= h && ld_z >= h && ld_x >= h && ld_out >= h,
                 "dcgc_gru_out_fwd: bad sizes");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(g && u && bh && z && x && out, "dcgc_gru_out_fwd: null pointer");
  dcgc_launch(gru_out_kernel, blocks_for(n * h, 256), 256, 0, (cudaStream_t)stream, g, ld_g, u, ld_u, bh, z, ld_z, x, ld_x, n, h,
                                                                           out, ld_out);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gru_out_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_setgather_attend_fwd(const float* x, int64_t ld_x, const float* q, int64_t ld_q,
                                         const int32_t* mol_ptr, const int32_t* mol_atoms, int64_t n_mols, int32_t h,
                                         int32_t max_atoms, float* qstar, int64_t ld_qs, void* stream) {
  DCGC_CHECK_ARG(n_mols >= 0 && h > 0 && max_atoms >= 0 && ld_x >= h && ld_q >= h && ld_qs >= 2 * (int64_t)h,
                 "dcgc_setgather_attend_fwd: bad sizes");
  DCGC_CHECK_ARG(max_atoms <= 12000, "dcgc_setgather_attend_fwd: molecules of more than 12000 atoms are not supported");
  if (n_mols == 0) return DCGC_OK;
  DCGC_CHECK_ARG(q && mol_ptr && qstar && (max_atoms == 0 || (x && mol_atoms)), "dcgc_setgather_attend_fwd: null pointer");
  dcgc_launch(setgather_attend_kernel, (unsigned)n_mols, kT, (size_t)(max_atoms > 0 ? max_atoms : 1) * 4, (cudaStream_t)stream, 
      x, ld_x, q, ld_q, mol_ptr, mol_atoms, h, qstar, ld_qs);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_setgather_attend_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_lstm_step_fwd(const float* zg, int64_t ld_z, const float* c_in, int64_t n, int32_t h, float* h_out,
                                  float* c_out, void* stream) {
  DCGC_CHECK_ARG(n >= 0 && h > 0 && ld_z >= 4 * (int64_t)h, "dcgc_lstm_step_fwd: bad sizes");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(zg && c_in && h_out && c_out, "dcgc_lstm_step_fwd: null pointer");
  dcgc_launch(lstm_step_kernel, blocks_for(n * h, 256), 256, 0, (cudaStream_t)stream, zg, ld_z, c_in, n, h, h_out, c_out);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_lstm_step_fwd");
  return DCGC_OK;
}

// ---- backward entry points -----------------------------------------------------------------------------------------
extern "C" int dcgc_pair_contract_bwd_x(const float* dz, int64_t ld_z, const float* pf, int64_t ld_pf,
                                        const int32_t* src_ptr, const int32_t* src_pair, const int32_t* pair_dst,
                                        int64_t n_src, int32_t n_pf, int32_t h, float* dx, int64_t ld_dx, void* stream) {
  DCGC_CHECK_ARG(n_src >= 0 && n_pf >= 0 && h > 0 && ld_dx >= h && ld_pf >= n_pf && ld_z >= (int64_t)(n_pf + 1) * h,
                 "dcgc_pair_contract_bwd_x: bad sizes");
  DCGC_CHECK_ARG(n_pf <= 32, "dcgc_pair_contract_bwd_x: at most 32 pair features are supported (got %d)", n_pf);
  if (n_src == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dz && src_ptr && dx && (n_pf == 0 || pf), "dcgc_pair_contract_bwd_x: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DcgcProfScope prof_scope("dcgc_pair_contract_bwd_x", st);
  if (n_pf <= 8)
    dcgc_launch(pair_contract_bwd_x_kernel<8>, (unsigned)n_src, kT, 0, st, dz, ld_z, pf, ld_pf, src_ptr, src_pair, pair_dst, n_pf, h, dx, ld_dx);
  else if (n_pf <= 16)
    dcgc_launch(pair_contract_bwd_x_kernel<16>, (unsigned)n_src, kT, 0, st, dz, ld_z, pf, ld_pf, src_ptr, src_pair, pair_dst, n_pf, h, dx, ld_dx);
  else
    dcgc_launch(pair_contract_bwd_x_kernel<32>, (unsigned)n_src, kT, 0, st, dz, ld_z, pf, ld_pf, src_ptr, src_pair, pair_dst, n_pf, h, dx, ld_dx);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_pair_contract_bwd_x");
  return DCGC_OK;
}

extern "C" int dcgc_gru_out_bwd(const float* dout, int64_t ld_do, const float* g, int64_t ld_g, const float* u,
                                int64_t ld_u, const float* bh, const float* z, int64_t ld_z, const float* x,
                                int64_t ld_x, int64_t n, int32_t h, float* dzg, int64_t ld_dz, float* dx,
                                int64_t ld_dx, float* dpre, int64_t ld_dp, void* stream) {
  DCGC_CHECK_ARG(n >= 0 && h > 0 && ld_do >= h && ld_g >= 3 * (int64_t)h && ld_u >= h && ld_z >= h && ld_x >= h &&
                     ld_dz >= h && ld_dx >= h && ld_dp >= h, "dcgc_gru_out_bwd: bad sizes");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dout && g && u && bh && z && x && dzg && dx && dpre, "dcgc_gru_out_bwd: null pointer");
  dcgc_launch(gru_out_bwd_kernel, blocks_for(n * h, 256), 256, 0, (cudaStream_t)stream, dout, ld_do, g, ld_g, u, ld_u, bh, z, ld_z,
                                                                               x, ld_x, n, h, dzg, ld_dz, dx, ld_dx, dpre,
                                                                               ld_dp);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gru_out_bwd");
  return DCGC_OK;
}

extern "C" int dcgc_gru_gates_bwd(const float* dzg, int64_t ld_dz, const float* dhr, int64_t ld_dhr, const float* dpre,
                                  int64_t ld_dp, const float* g, int64_t ld_g, const float* bz, const float* br,
                                  const float* hprev, int64_t ld_h, int64_t n, int32_t h, float* dg, int64_t ld_dg,
                                  float* dh, int64_t ld_dh, void* stream) {
  DCGC_CHECK_ARG(n >= 0 && h > 0 && ld_dz >= h && ld_dhr >= h && ld_dp >= h && ld_g >= 3 * (int64_t)h && ld_h >= h &&
                     ld_dg >= 3 * (int64_t)h && ld_dh >= h, "dcgc_gru_gates_bwd: bad sizes");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dzg && dhr && dpre && g && bz && br && hprev && dg && dh, "dcgc_gru_gates_bwd: null pointer");
  dcgc_launch(gru_gates_bwd_kernel, blocks_for(n * h, 256), 256, 0, (cudaStream_t)stream, dzg, ld_dz, dhr, ld_dhr, dpre, ld_dp, g,
                                                                                 ld_g, bz, br, hprev, ld_h, n, h, dg,
                                                                                 ld_dg, dh, ld_dh);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gru_gates_bwd");
  return DCGC_OK;
}

extern "C" int dcgc_setgather_attend_bwd(const float* x, int64_t ld_x, const float* q, int64_t ld_q, const float* dqs,
                                         int64_t ld_dqs, const int32_t* mol_ptr, const int32_t* mol_atoms,
                                         int64_t n_mols, int32_t h, int32_t max_atoms, float* dx, int64_t ld_dx,
                                         float* dq, int64_t ld_dq, void* stream) {
  DCGC_CHECK_ARG(n_mols >= 0 && h > 0 && max_atoms >= 0 && ld_x >= h && ld_q >= h && ld_dqs >= 2 * (int64_t)h &&
                     ld_dx >= h && ld_dq >= h, "dcgc_setgather_attend_bwd: bad sizes");
  DCGC_CHECK_ARG(max_atoms <= 6000, "dcgc_setgather_attend_bwd: molecules of more than 6000 atoms are not supported");
  if (n_mols == 0) return DCGC_OK;
  DCGC_CHECK_ARG(q && dqs && mol_ptr && dq && (max_atoms == 0 || (x && mol_atoms && dx)),
                 "dcgc_setgather_attend_bwd: null pointer");
  const int ma = max_atoms > 0 ? max_atoms : 1;
  dcgc_launch(setgather_attend_bwd_kernel, (unsigned)n_mols, kT, (size_t)ma * 8, (cudaStream_t)stream, 
      x, ld_x, q, ld_q, dqs, ld_dqs, mol_ptr, mol_atoms, h, ma, dx, ld_dx, dq, ld_dq);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_setgather_attend_bwd");
  return DCGC_OK;
}

extern "C" int dcgc_lstm_step_bwd(const float* zg, int64_t ld_z, const float* c_in, const float* dh, const float* dc_out,
                                  int64_t n, int32_t h, float* dz, int64_t ld_dz, float* dc_in, void* stream) {
  DCGC_CHECK_ARG(n >= 0 && h > 0 && ld_z >= 4 * (int64_t)h && ld_dz >= 4 * (int64_t)h, "dcgc_lstm_step_bwd: bad sizes");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(zg && c_in && dz && dc_in, "dcgc_lstm_step_bwd: null pointer");
  dcgc_launch(lstm_step_bwd_kernel, blocks_for(n * h, 256), 256, 0, (cudaStream_t)stream, zg, ld_z, c_in, dh, dc_out, n, h, dz,
                                                                                 ld_dz, dc_in);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_lstm_step_bwd");
  return DCGC_OK;
}

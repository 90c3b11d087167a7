// HBM-bound kernels of the GraphConv path for sm_100a: neighbour gather-sum (K1/K5/K8),
// GraphPool max (+argslot) and its scatter-free backward (K3/K7), GraphGather segment sum/max
// and its backward (K4/K7), row permutation.
//
// Common shape: one thread owns one 16-byte column group (float4) of one OUTPUT row, so every
// output element is written exactly once by exactly one thread (no float atomics, fixed
// summation order = index order), consecutive threads touch consecutive 16-byte groups of a
// row (128-bit coalesced loads of whole feature rows) and consecutive rows of the output.
// A scalar (VEC=1) instantiation handles widths / strides that are not 16-byte aligned
// (the reference's raw [N,75] feature matrix).
#include "common.h"

namespace {

constexpr int kThreads = 256;

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
  static __device__ __forceinline__ T load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void store(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <>
struct Vec<1> {
  using T = float;
  static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, T v) { *p = v; }
  static __device__ __forceinline__ T zero() { return 0.f; }
};

__device__ __forceinline__ void vadd(float4& a, const float4 b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void vadd(float& a, const float b) { a += b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline unsigned grid_for(int64_t work) { return (unsigned)((work + kThreads - 1) / kThreads); }

// ------------------------------------------------------------------------------------------
// permute rows (+ zero the pad columns)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
permute_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ perm,
                    int64_t n_rows, int n_feat, float* __restrict__ dst, int64_t ld_dst) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = t / ld_dst;
  const int c = (int)(t - row * ld_dst);
  if (row >= n_rows) return;
  dst[row * ld_dst + c] = c < n_feat ? __ldg(src + (int64_t)__ldg(perm + row) * ld_src + c) : 0.f;
}

// same, from an int8 feature matrix (PackedMols stores integer-valued feature matrices — one-hots, formal charge,
// radical electrons: every ConvMol feature — as int8; the conversion to fp32 is exact)
__global__ void __launch_bounds__(kThreads)
permute_rows_i8_kernel(const int8_t* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ perm,
                       int64_t n_rows, int n_feat, float* __restrict__ dst, int64_t ld_dst) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int groups = (int)(ld_dst >> 2);
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) * 4;
  if (row >= n_rows) return;
  const int8_t* s = src + (int64_t)__ldg(perm + row) * ld_src;
  float4 v;
  v.x = c + 0 < n_feat ? (float)__ldg(s + c + 0) : 0.f;
  v.y = c + 1 < n_feat ? (float)__ldg(s + c + 1) : 0.f;
  v.z = c + 2 < n_feat ? (float)__ldg(s + c + 2) : 0.f;
  v.w = c + 3 < n_feat ? (float)__ldg(s + c + 3) : 0.f;
  *reinterpret_cast<float4*>(dst + row * ld_dst + c) = v;
}

// ------------------------------------------------------------------------------------------
// K1/K5/K8: CSR gather-sum
// ------------------------------------------------------------------------------------------
// Degree-bucketed rows (the ConvMol layout): row i of bucket d has exactly d entries, so the CSR offsets are
// arithmetic and the row_ptr load — one of the three dependent loads row_ptr -> idx -> row that bound this
// kernel (profiles/r1c_ncu_gather_sum.md) — disappears.
struct DegBuckets {
  int row0[DCGC_N_DEG + 1];   // first row of bucket d; row0[11] = number of rows
  int e0[DCGC_N_DEG];         // first entry of bucket d
};

template <int VEC, bool BUCKET>
__global__ void __launch_bounds__(kThreads)
gather_sum_kernel(const float* __restrict__ x, int64_t ld_x, const int32_t* __restrict__ row_ptr,
                  const DegBuckets bk, const int32_t* __restrict__ idx, int64_t n_rows, int groups, int width,
                  const float* addend, int64_t ld_add, float* out, int64_t ld_out) {
  dcgc_griddep_wait();
  using V = Vec<VEC>;
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) * VEC;
  if (row >= n_rows || c >= width) return;
  int e, e1;
  if (BUCKET) {
    // static indices only: a run-time index into the by-value parameter struct makes the compiler copy it to
    // local memory (measured: 52 us instead of 32 us per launch)
    int d = 0, r0 = bk.row0[0], eb = bk.e0[0];
#pragma unroll
    for (int k = 1; k < DCGC_N_DEG; ++k) {
      const bool ge = (int)row >= bk.row0[k];
      d = ge ? k : d;
      r0 = ge ? bk.row0[k] : r0;
      eb = ge ? bk.e0[k] : eb;
    }
    e = eb + ((int)row - r0) * d;
    e1 = e + d;
  } else {
    e = __ldg(row_ptr + row);
    e1 = __ldg(row_ptr + row + 1);
  }
  typename V::T acc = V::zero();
  if (addend) acc = *reinterpret_cast<const typename V::T*>(addend + row * ld_add + c);  // may alias out
  const float* xc = x + c;
  // four independent row loads in flight per thread (degrees 1..4 carry almost all rows)
  for (; e + 4 <= e1; e += 4) {
    const int j0 = __ldg(idx + e), j1 = __ldg(idx + e + 1), j2 = __ldg(idx + e + 2), j3 = __ldg(idx + e + 3);
    const typename V::T v0 = V::load(xc + (int64_t)j0 * ld_x);
    const typename V::T v1 = V::load(xc + (int64_t)j1 * ld_x);
    const typename V::T v2 = V::load(xc + (int64_t)j2 * ld_x);
    const typename V::T v3 = V::load(xc + (int64_t)j3 * ld_x);
    vadd(acc, v0); vadd(acc, v1); vadd(acc, v2); vadd(acc, v3);
  }
  // remainder of 1..3 rows: all of them in flight at once (degree-3 atoms, ~27 % of a molecule, used to pay two
  // dependent round trips: a pair, then the single)
  const int rem = e1 - e;
  if (rem == 3) {
    const int j0 = __ldg(idx + e), j1 = __ldg(idx + e + 1), j2 = __ldg(idx + e + 2);
    const typename V::T v0 = V::load(xc + (int64_t)j0 * ld_x);
    const typename V::T v1 = V::load(xc + (int64_t)j1 * ld_x);
    const typename V::T v2 = V::load(xc + (int64_t)j2 * ld_x);
    vadd(acc, v0); vadd(acc, v1); vadd(acc, v2);
  } else if (rem == 2) {
    const int j0 = __ldg(idx + e), j1 = __ldg(idx + e + 1);
    const typename V::T v0 = V::load(xc + (int64_t)j0 * ld_x);
    const typename V::T v1 = V::load(xc + (int64_t)j1 * ld_x);
    vadd(acc, v0); vadd(acc, v1);
  } else if (rem == 1) {
    vadd(acc, V::load(xc + (int64_t)__ldg(idx + e) * ld_x));
  }
  V::store(out + row * ld_out + c, acc);
}

// ------------------------------------------------------------------------------------------
// K3: GraphPool forward.  First slot attaining the max wins (strict > when scanning
// self, nbr_0, nbr_1, ...), which is torch.max's tie rule on the reference's
// concat([self, gathered]) (torch_models/layers.py:6358-6361).
// ------------------------------------------------------------------------------------------
template <int VEC, bool AFFINE, bool ARG>
__global__ void __launch_bounds__(kThreads)
pool_fwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ scale,
                const float* __restrict__ shift, const int32_t* __restrict__ row_ptr,
                const int32_t* __restrict__ col_idx, int64_t n_rows, int groups, int width,
                float* __restrict__ out, int64_t ld_out, uint8_t* __restrict__ arg, int64_t ld_arg) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) * VEC;
  if (row >= n_rows || c >= width) return;
  float sc[VEC], sh[VEC], m[VEC];
  uint8_t a[VEC];
  if (AFFINE) {
    if (VEC == 4) {   // c is a multiple of 4 and the statistics buffers are 16-byte aligned
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c));
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(shift + c));
      sc[0] = s4.x; sc[1 % VEC] = s4.y; sc[2 % VEC] = s4.z; sc[3 % VEC] = s4.w;
      sh[0] = h4.x; sh[1 % VEC] = h4.y; sh[2 % VEC] = h4.z; sh[3 % VEC] = h4.w;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) { sc[v] = __ldg(scale + c + v); sh[v] = __ldg(shift + c + v); }
    }
  }
  auto load = [&](int64_t r, float* dst) {
    if (VEC == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + c));
      dst[0] = q.x; dst[1 % VEC] = q.y; dst[2 % VEC] = q.z; dst[3 % VEC] = q.w;
    } else {
      dst[0] = __ldg(x + r * ld_x + c);
    }
    if (AFFINE) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) dst[v] = fmaf(dst[v], sc[v], sh[v]);
    }
  };
  load(row, m);
#pragma unroll
  for (int v = 0; v < VEC; ++v) a[v] = 0;
  const int e0 = __ldg(row_ptr + row), e1 = __ldg(row_ptr + row + 1);
  int e = e0;
  auto cmp = [&](const float* u, int slot) {
#pragma unroll
    for (int v = 0; v < VEC; ++v)
      if (u[v] > m[v]) { m[v] = u[v]; a[v] = (uint8_t)slot; }
  };
  // up to four neighbour rows in flight at once (degree 3 and 4 used to pay two dependent round trips);
  // the comparisons stay in slot order, so the first-slot tie rule is unchanged
  for (; e + 4 <= e1; e += 4) {
    float u0[VEC], u1[VEC], u2[VEC], u3[VEC];
    const int j0 = __ldg(col_idx + e), j1 = __ldg(col_idx + e + 1), j2 = __ldg(col_idx + e + 2),
              j3 = __ldg(col_idx + e + 3);
    load(j0, u0); load(j1, u1); load(j2, u2); load(j3, u3);
    cmp(u0, e - e0 + 1); cmp(u1, e - e0 + 2); cmp(u2, e - e0 + 3); cmp(u3, e - e0 + 4);
  }
  const int rem = e1 - e;
  if (rem == 3) {
    float u0[VEC], u1[VEC], u2[VEC];
    const int j0 = __ldg(col_idx + e), j1 = __ldg(col_idx + e + 1), j2 = __ldg(col_idx + e + 2);
    load(j0, u0); load(j1, u1); load(j2, u2);
    cmp(u0, e - e0 + 1); cmp(u1, e - e0 + 2); cmp(u2, e - e0 + 3);
  } else if (rem == 2) {
    float u0[VEC], u1[VEC];
    const int j0 = __ldg(col_idx + e), j1 = __ldg(col_idx + e + 1);
    load(j0, u0); load(j1, u1);
    cmp(u0, e - e0 + 1); cmp(u1, e - e0 + 2);
  } else if (rem == 1) {
    float u0[VEC];
    load(__ldg(col_idx + e), u0);
    cmp(u0, e - e0 + 1);
  }
  if (VEC == 4) {
    *reinterpret_cast<float4*>(out + row * ld_out + c) = make_float4(m[0], m[1 % VEC], m[2 % VEC], m[3 % VEC]);
    if (ARG) *reinterpret_cast<uchar4*>(arg + row * ld_arg + c) = make_uchar4(a[0], a[1 % VEC], a[2 % VEC], a[3 % VEC]);
  } else {
    out[row * ld_out + c] = m[0];
    if (ARG) arg[row * ld_arg + c] = a[0];
  }
}

// K7: GraphPool backward over the transposed CSR (each source row gathers from the rows that
// selected it; nothing is scattered).
template <int VEC, bool AFFINE>
__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy, const uint8_t* __restrict__ arg,
                int64_t ld_arg, const float* __restrict__ scale, const int32_t* __restrict__ t_row_ptr,
                const int32_t* __restrict__ t_src, const int32_t* __restrict__ t_slot, int64_t n_rows,
                int groups, int width, float* __restrict__ dx, int64_t ld_dx) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) * VEC;
  if (row >= n_rows || c >= width) return;
  float acc[VEC];
  auto take = [&](int64_t r, int slot, bool first) {
    float d[VEC];
    uint8_t a[VEC];
    if (VEC == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(dy + r * ld_dy + c));
      const uchar4 b = __ldg(reinterpret_cast<const uchar4*>(arg + r * ld_arg + c));
      d[0] = q.x; d[1 % VEC] = q.y; d[2 % VEC] = q.z; d[3 % VEC] = q.w;
      a[0] = b.x; a[1 % VEC] = b.y; a[2 % VEC] = b.z; a[3 % VEC] = b.w;
    } else {
      d[0] = __ldg(dy + r * ld_dy + c);
      a[0] = __ldg(arg + r * ld_arg + c);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float g = (a[v] == slot) ? d[v] : 0.f;
      acc[v] = first ? g : acc[v] + g;
    }
  };
  take(row, 0, true);
  const int e1 = __ldg(t_row_ptr + row + 1);
  for (int e = __ldg(t_row_ptr + row); e < e1; ++e) take(__ldg(t_src + e), __ldg(t_slot + e) + 1, false);
  if (AFFINE) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] *= __ldg(scale + c + v);
  }
  if (VEC == 4) {
    *reinterpret_cast<float4*>(dx + row * ld_dx + c) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
  } else {
    dx[row * ld_dx + c] = acc[0];
  }
}


__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------
// K4: GraphGather.  One thread owns one 16-byte column group of one molecule and walks the
// molecule's rows in ascending order (the order CPU scatter_add accumulates in,
// deepchem/utils/pytorch_utils.py:70-72).  Max ties keep the lowest row (strict >).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DCGC_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == DCGC_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float act_grad_from_out(float o, int act) {
  if (act == DCGC_ACT_RELU) return o > 0.f ? 1.f : 0.f;
  if (act == DCGC_ACT_TANH) return 1.f - o * o;
  return 1.f;
}

template <int VEC, bool ARG>
__global__ void __launch_bounds__(kThreads)
gather_fwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ scale,
                  const float* __restrict__ shift, const int32_t* __restrict__ mol_ptr,
                  const int32_t* __restrict__ mol_atoms, int64_t n_seg, int groups, int width, int act,
                  float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argrow,
                  const float* __restrict__ mean, float* __restrict__ zc_sum, float* __restrict__ zc_arg) {
  dcgc_griddep_wait();
  // zc_sum / zc_arg (training with BatchNorm folded in, optional): per molecule the sum over its rows of the RAW input
  // centred on the batch mean, and the centred raw value of the row that attains the max — what the BatchNorm backward
  // of the layer in front needs to form its column sums at MOLECULE level (dense_bn_sums_kernel, model.cu)
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t g = t / groups;
  const int c = (int)(t - g * groups) * VEC;
  if (g >= n_seg || c >= width) return;
  float s[VEC], m[VEC], sc[VEC], sh[VEC], mu[VEC], zs[VEC], za[VEC];
  int32_t a[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    s[v] = 0.f; m[v] = -INFINITY; a[v] = -1;
    sc[v] = scale ? __ldg(scale + c + v) : 1.f;
    sh[v] = scale ? __ldg(shift + c + v) : 0.f;
    mu[v] = zc_sum ? __ldg(mean + c + v) : 0.f;
    zs[v] = 0.f; za[v] = 0.f;
  }
  // Eight atoms at a time: their row indices, then their rows, are loaded as independent requests and only then added
  // in ascending row order (sum order and the lowest-row tie rule of the max are unchanged).  One atom per iteration
  // was one dependent index -> row round trip after the other: 0.49 of the HBM peak in situ.
  constexpr int U = 8;
  const int e1 = __ldg(mol_ptr + g + 1);
  for (int e = __ldg(mol_ptr + g); e < e1; e += U) {
    int32_t r[U];
#pragma unroll
    for (int i = 0; i < U; ++i) r[i] = e + i < e1 ? __ldg(mol_atoms + e + i) : -1;
    float u[U][VEC];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (r[i] < 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) u[i][v] = 0.f;
      } else if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + (int64_t)r[i] * ld_x + c));
        u[i][0] = q.x; u[i][1 % VEC] = q.y; u[i][2 % VEC] = q.z; u[i][3 % VEC] = q.w;
      } else {
        u[i][0] = __ldg(x + (int64_t)r[i] * ld_x + c);
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (r[i] < 0) continue;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float t2 = scale ? fmaf(u[i][v], sc[v], sh[v]) : u[i][v];
        s[v] += t2;
        zs[v] += u[i][v] - mu[v];
        if (t2 > m[v] || a[v] < 0) { m[v] = t2; a[v] = r[i]; za[v] = u[i][v] - mu[v]; }
      }
    }
  }
  if (zc_sum) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      zc_sum[g * (int64_t)width + c + v] = zs[v];
      zc_arg[g * (int64_t)width + c + v] = za[v];
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    out[g * ld_out + c + v] = apply_act(s[v], act);
    out[g * ld_out + width + c + v] = apply_act(m[v], act);
    if (ARG) argrow[g * (int64_t)width + c + v] = a[v];
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
gather_bwd_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ out,
                  int64_t ld_out, const int32_t* __restrict__ argrow, const int32_t* __restrict__ membership,
                  int64_t n_rows, int groups, int width, int act, float* __restrict__ dx, int64_t ld_dx) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) * VEC;
  if (row >= n_rows || c >= width) return;
  const int64_t g = __ldg(membership + row);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int cc = c + v;
    if (cc >= width) break;
    float d = __ldg(dout + g * ld_dout + cc) * act_grad_from_out(__ldg(out + g * ld_out + cc), act);
    if (__ldg(argrow + g * (int64_t)width + cc) == (int32_t)row)
      d += __ldg(dout + g * ld_dout + width + cc) * act_grad_from_out(__ldg(out + g * ld_out + width + cc), act);
    dx[row * ld_dx + cc] = d;
  }
}


// vectorised GraphGather backward: one thread = 4 columns of one atom row; the per-molecule
// operands (dout, out, argrow: shared by the ~25 atoms of a molecule) come as 128-bit loads
__global__ void __launch_bounds__(kThreads)
gather_bwd_vec_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ out,
                      int64_t ld_out, const int32_t* __restrict__ argrow, const int32_t* __restrict__ membership,
                      int64_t n_rows, int groups, int width, int act, float* __restrict__ dx, int64_t ld_dx) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) * 4;
  if (row >= n_rows) return;
  const int64_t g = __ldg(membership + row);
  const float4 ds = ldg4(dout + g * ld_dout + c), os = ldg4(out + g * ld_out + c);
  const float4 dm = ldg4(dout + g * ld_dout + width + c), om = ldg4(out + g * ld_out + width + c);
  const int4 ar = __ldg(reinterpret_cast<const int4*>(argrow + g * (int64_t)width + c));
  const int r32 = (int)row;
  float4 o;
  o.x = ds.x * act_grad_from_out(os.x, act) + (ar.x == r32 ? dm.x * act_grad_from_out(om.x, act) : 0.f);
  o.y = ds.y * act_grad_from_out(os.y, act) + (ar.y == r32 ? dm.y * act_grad_from_out(om.y, act) : 0.f);
  o.z = ds.z * act_grad_from_out(os.z, act) + (ar.z == r32 ? dm.z * act_grad_from_out(om.z, act) : 0.f);
  o.w = ds.w * act_grad_from_out(os.w, act) + (ar.w == r32 ? dm.w * act_grad_from_out(om.w, act) : 0.f);
  *reinterpret_cast<float4*>(dx + row * ld_dx + c) = o;
}

constexpr int kGbRowLanes = 16;

// GraphGather backward that writes the gradient of the layer BEHIND the BatchNorm + ReLU in front of it directly:
//   dA = d sum + [row is the arg-max] d max      (as gather_bwd_vec_kernel),
//   G  = relu'(z) * c1 * (dA - c2 - (z - mean) * invstd * c3)      (as bn_relu_bwd_apply)
// with the per-column coefficients already known (dense_bn_sums_kernel + finalize): the separate apply pass over dA and
// z (3 x N x width x 4 bytes) disappears.  Same walk as gather_bwd_stats_kernel: four rows in flight per thread.
__global__ void __launch_bounds__(32 * kGbRowLanes, 2)
gather_bwd_apply_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ out, int64_t ld_out,
                        const int32_t* __restrict__ argrow, const int32_t* __restrict__ membership, int64_t n_rows,
                        int width, int act, const float* __restrict__ z, int64_t ld_z, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ coef, float* __restrict__ g,
                        int64_t ld_g, int64_t rows_per_chunk) {
  dcgc_griddep_wait();
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + 4 * cx;
  if (c >= width) return;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = min(n_rows, r0 + rows_per_chunk);
  const float4 m = ldg4(mean + c), is = ldg4(invstd + c);
  const float4 c1 = ldg4(coef + c), c2 = ldg4(coef + width + c), c3 = ldg4(coef + 2 * width + c);
  for (int64_t rb = r0 + ry; rb < r1; rb += 4 * kGbRowLanes) {
    int64_t mol[4];
    float4 zv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + (int64_t)u * kGbRowLanes;
      mol[u] = r < r1 ? (int64_t)__ldg(membership + r) : -1;
      zv[u] = r < r1 ? ldg4(z + r * ld_z + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (mol[u] < 0) continue;
      const int r32 = (int)(rb + (int64_t)u * kGbRowLanes);
      const float4 ds = ldg4(dout + mol[u] * ld_dout + c), os = ldg4(out + mol[u] * ld_out + c);
      const float4 dm = ldg4(dout + mol[u] * ld_dout + width + c), om = ldg4(out + mol[u] * ld_out + width + c);
      const int4 ar = __ldg(reinterpret_cast<const int4*>(argrow + mol[u] * (int64_t)width + c));
      float4 o;
      o.x = ds.x * act_grad_from_out(os.x, act) + (ar.x == r32 ? dm.x * act_grad_from_out(om.x, act) : 0.f);
      o.y = ds.y * act_grad_from_out(os.y, act) + (ar.y == r32 ? dm.y * act_grad_from_out(om.y, act) : 0.f);
      o.z = ds.z * act_grad_from_out(os.z, act) + (ar.z == r32 ? dm.z * act_grad_from_out(om.z, act) : 0.f);
      o.w = ds.w * act_grad_from_out(os.w, act) + (ar.w == r32 ? dm.w * act_grad_from_out(om.w, act) : 0.f);
      const float4 v = zv[u];
      float4 q;
      q.x = v.x > 0.f ? c1.x * (o.x - c2.x - (v.x - m.x) * is.x * c3.x) : 0.f;
      q.y = v.y > 0.f ? c1.y * (o.y - c2.y - (v.y - m.y) * is.y * c3.y) : 0.f;
      q.z = v.z > 0.f ? c1.z * (o.z - c2.z - (v.z - m.z) * is.z * c3.z) : 0.f;
      q.w = v.w > 0.f ? c1.w * (o.w - c2.w - (v.w - m.w) * is.w * c3.w) : 0.f;
      *reinterpret_cast<float4*>(g + (int64_t)r32 * ld_g + c) = q;
    }
  }
}

// Column sums of dA = GraphGather-backward(d fingerprint) WITHOUT forming dA: every atom of a molecule receives the
// molecule's d sum and one atom per column its d max, so
//   sum_r dA[r, c]             = sum_mol ( atoms(mol) * ds[mol, c] + dm[mol, c] )
//   sum_r dA[r, c] (z - mean)  = sum_mol ( ds[mol, c] * zc_sum[mol, c] + dm[mol, c] * zc_arg[mol, c] )
// (ds / dm = d fingerprint times the activation derivative; dm only where the molecule has atoms).  4096 molecules
// instead of 102 k atoms.  Block = 32 columns x 8 molecule lanes over a chunk of 128 molecules; float64 sums, lanes
// combined in order; one row of partials per chunk in the layout bn_bwd_finalize reads (part[1] = uncentred).
constexpr int kDbsMols = 128;
constexpr int kDbsLanes = 32;      // molecule lanes per block: 32 columns x 32 lanes = 1 024 threads, 4 molecules each
__global__ void __launch_bounds__(32 * kDbsLanes)
dense_bn_sums_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ out, int64_t ld_out,
                     const int32_t* __restrict__ argrow, const int32_t* __restrict__ mol_ptr, int64_t n_seg, int width,
                     int act, const float* __restrict__ zc_sum, const float* __restrict__ zc_arg,
                     const float* __restrict__ mean, double* __restrict__ part) {
  dcgc_griddep_wait();
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const int64_t m0 = (int64_t)blockIdx.x * kDbsMols, m1 = min(n_seg, m0 + kDbsMols);
  double a = 0.0, b = 0.0;
  if (c < width) {
    // every load of the thread's four molecules is issued before the first use (the first version walked 16 molecules
    // per thread with two dependent round trips each: 31 us for a 20 MB problem)
    constexpr int U = kDbsMols / kDbsLanes;
    int cnt[U], ar[U];
    float d0[U], o0[U], d1[U], o1[U], zs[U], za[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t mol = m0 + ly + (int64_t)u * kDbsLanes;
      const bool ok = mol < m1;
      const int64_t mm = ok ? mol : m0;
      cnt[u] = ok ? __ldg(mol_ptr + mm + 1) - __ldg(mol_ptr + mm) : 0;
      ar[u] = __ldg(argrow + mm * (int64_t)width + c);
      d0[u] = __ldg(dout + mm * ld_dout + c); o0[u] = __ldg(out + mm * ld_out + c);
      d1[u] = __ldg(dout + mm * ld_dout + width + c); o1[u] = __ldg(out + mm * ld_out + width + c);
      zs[u] = __ldg(zc_sum + mm * (int64_t)width + c); za[u] = __ldg(zc_arg + mm * (int64_t)width + c);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (cnt[u] <= 0) continue;
      const float ds = d0[u] * act_grad_from_out(o0[u], act);
      const float dm = ar[u] >= 0 ? d1[u] * act_grad_from_out(o1[u], act) : 0.f;
      a += (double)cnt[u] * (double)ds + (double)dm;
      b += (double)ds * (double)zs[u] + (double)dm * (double)za[u];
    }
  }
  __shared__ double sh[2][kDbsLanes][32];
  sh[0][ly][cx] = a; sh[1][ly][cx] = b;
  __syncthreads();
  if (ly == 0 && c < width) {
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int l = 0; l < kDbsLanes; ++l) { sa += sh[0][l][cx]; sb += sh[1][l][cx]; }
    part[((int64_t)blockIdx.x * 2) * width + c] = sa;
    part[((int64_t)blockIdx.x * 2 + 1) * width + c] = sb + (double)__ldg(mean + c) * sa;
  }
}

// GraphGather backward with the BatchNorm-backward column sums of its result fused in: besides dx it emits, per block,
// sum_r dx[r, c] and sum_r dx[r, c] * z[r, c] (z = the BatchNorm input of the same rows) — what a separate
// two-tensor pass over dx and z (col_moments_partial: 2 x N x width x 4 bytes, 25 us at the bench shape) computed.
// Block = 32 column lanes (float4 -> 128 columns) x 16 row lanes; a block walks a contiguous row range with four rows
// in flight per thread (the one-row-per-thread kernel above sits at 0.35 of the HBM peak: five dependent loads and
// no second row to overlap them with).  fp32 partial sums over four rows, float64 from there on, the 16 row lanes
// combined in lane order through shared memory: deterministic.  part: [blocks][2][width] doubles.
__global__ void __launch_bounds__(32 * kGbRowLanes, 2)
gather_bwd_stats_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ out, int64_t ld_out,
                        const int32_t* __restrict__ argrow, const int32_t* __restrict__ membership, int64_t n_rows,
                        int width, int act, float* __restrict__ dx, int64_t ld_dx, const float* __restrict__ z,
                        int64_t ld_z, int64_t rows_per_chunk, double* __restrict__ part, const DcgcBnFin bnfin) {
  dcgc_griddep_wait();
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + 4 * cx;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = min(n_rows, r0 + rows_per_chunk);
  double da[4] = {0, 0, 0, 0}, dab[4] = {0, 0, 0, 0};
  if (c < width) {
    for (int64_t rb = r0 + ry; rb < r1; rb += 4 * kGbRowLanes) {
      int64_t g[4];
      float4 zv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = rb + (int64_t)u * kGbRowLanes;
        g[u] = r < r1 ? (int64_t)__ldg(membership + r) : -1;
        zv[u] = r < r1 ? ldg4(z + r * ld_z + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        o[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g[u] < 0) continue;
        const int r32 = (int)(rb + (int64_t)u * kGbRowLanes);
        const float4 ds = ldg4(dout + g[u] * ld_dout + c), os = ldg4(out + g[u] * ld_out + c);
        const float4 dm = ldg4(dout + g[u] * ld_dout + width + c), om = ldg4(out + g[u] * ld_out + width + c);
        const int4 ar = __ldg(reinterpret_cast<const int4*>(argrow + g[u] * (int64_t)width + c));
        o[u].x = ds.x * act_grad_from_out(os.x, act) + (ar.x == r32 ? dm.x * act_grad_from_out(om.x, act) : 0.f);
        o[u].y = ds.y * act_grad_from_out(os.y, act) + (ar.y == r32 ? dm.y * act_grad_from_out(om.y, act) : 0.f);
        o[u].z = ds.z * act_grad_from_out(os.z, act) + (ar.z == r32 ? dm.z * act_grad_from_out(om.z, act) : 0.f);
        o[u].w = ds.w * act_grad_from_out(os.w, act) + (ar.w == r32 ? dm.w * act_grad_from_out(om.w, act) : 0.f);
        *reinterpret_cast<float4*>(dx + (int64_t)r32 * ld_dx + c) = o[u];
      }
      float sa[4] = {0, 0, 0, 0}, sab[4] = {0, 0, 0, 0};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sa[0] += o[u].x; sa[1] += o[u].y; sa[2] += o[u].z; sa[3] += o[u].w;
        sab[0] = fmaf(o[u].x, zv[u].x, sab[0]); sab[1] = fmaf(o[u].y, zv[u].y, sab[1]);
        sab[2] = fmaf(o[u].z, zv[u].z, sab[2]); sab[3] = fmaf(o[u].w, zv[u].w, sab[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { da[e] += sa[e]; dab[e] += sab[e]; }
    }
  }
  __shared__ double sh[2][kGbRowLanes][128];
#pragma unroll
  for (int e = 0; e < 4; ++e) { sh[0][ry][4 * cx + e] = da[e]; sh[1][ry][4 * cx + e] = dab[e]; }
  __syncthreads();
  if (threadIdx.x < 256) {
    const int q = threadIdx.x >> 7, col = threadIdx.x & 127;
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < kGbRowLanes; ++l) s += sh[q][l][col];
    const int cc = blockIdx.y * 128 + col;
    if (cc < width) part[((int64_t)blockIdx.x * 2 + q) * width + cc] = s;
  }
  __syncthreads();                                                     // sh is free again
  dcgc_bn_fin_last_cta(bnfin, (int)gridDim.x, gridDim.x * gridDim.y, 0, 32 * kGbRowLanes, (int)threadIdx.x, &sh[0][0][0]);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int dcgc_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  DCGC_CUDA_CALL(cudaGetDevice(&dev));
  DCGC_CUDA_CALL(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}

extern "C" int dcgc_h2d_chunked(void* dst, const void* src, int64_t bytes, int64_t chunk_bytes, void* stream) {
  DCGC_CHECK_ARG(bytes >= 0, "dcgc_h2d_chunked: negative size");
  if (bytes == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dst && src, "dcgc_h2d_chunked: null pointer");
  if (chunk_bytes <= 0 || chunk_bytes > bytes) chunk_bytes = bytes;
  chunk_bytes = (chunk_bytes + 255) / 256 * 256;
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t o = 0; o < bytes; o += chunk_bytes) {
    const int64_t n = bytes - o < chunk_bytes ? bytes - o : chunk_bytes;
    DCGC_CUDA_CALL(cudaMemcpyAsync(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, (size_t)n,
                                   cudaMemcpyHostToDevice, st));
  }
  return DCGC_OK;
}

extern "C" int dcgc_permute_rows(const float* src, int64_t ld_src, const int32_t* perm, int64_t n_rows,
                                 int32_t n_feat, float* dst, int64_t ld_dst, void* stream) {
  DCGC_CHECK_ARG(n_rows >= 0 && n_feat >= 0 && ld_src >= n_feat && ld_dst >= n_feat,
                 "dcgc_permute_rows: bad sizes");
  if (n_rows == 0 || ld_dst == 0) return DCGC_OK;
  DCGC_CHECK_ARG(src && perm && dst, "dcgc_permute_rows: null pointer");
  dcgc_launch(permute_rows_kernel, grid_for(n_rows * ld_dst), kThreads, 0, (cudaStream_t)stream, 
      src, ld_src, perm, n_rows, n_feat, dst, ld_dst);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_permute_rows");
  return DCGC_OK;
}

extern "C" int dcgc_permute_rows_i8(const int8_t* src, int64_t ld_src, const int32_t* perm, int64_t n_rows,
                                    int32_t n_feat, float* dst, int64_t ld_dst, void* stream) {
  DCGC_CHECK_ARG(n_rows >= 0 && n_feat >= 0 && ld_src >= n_feat && ld_dst >= n_feat, "dcgc_permute_rows_i8: bad sizes");
  if (n_rows == 0 || ld_dst == 0) return DCGC_OK;
  DCGC_CHECK_ARG(src && perm && dst, "dcgc_permute_rows_i8: null pointer");
  DCGC_CHECK_ARG(ld_dst % 4 == 0 && aligned16(dst), "dcgc_permute_rows_i8: dst rows must be 16-byte aligned");
  dcgc_launch(permute_rows_i8_kernel, grid_for(n_rows * (ld_dst / 4)), kThreads, 0, (cudaStream_t)stream, 
      src, ld_src, perm, n_rows, n_feat, dst, ld_dst);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_permute_rows_i8");
  return DCGC_OK;
}

static int gather_sum_impl(const float* x, int64_t ld_x, const int32_t* row_ptr, const int64_t* deg_count,
                           const int32_t* idx, int64_t n_rows_out, int32_t width, const float* addend, int64_t ld_add,
                           float* out, int64_t ld_out, void* stream) {
  DCGC_CHECK_ARG(n_rows_out >= 0 && width >= 0 && ld_x >= width && ld_out >= width &&
                     (addend == nullptr || ld_add >= width), "dcgc_gather_sum: bad sizes");
  if (n_rows_out == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && (row_ptr || deg_count) && out, "dcgc_gather_sum: null pointer");
  DegBuckets bk{};
  if (deg_count) {
    int64_t r = 0, e = 0;
    for (int d = 0; d < DCGC_N_DEG; ++d) {
      DCGC_CHECK_ARG(deg_count[d] >= 0, "dcgc_gather_sum_bucketed: negative bucket size");
      bk.row0[d] = (int)r;
      bk.e0[d] = (int)e;
      r += deg_count[d];
      e += (int64_t)d * deg_count[d];
    }
    bk.row0[DCGC_N_DEG] = (int)r;
    DCGC_CHECK_ARG(r == n_rows_out && e < ((int64_t)1 << 31), "dcgc_gather_sum_bucketed: bucket sizes do not add up to n_rows_out");
  }
  DcgcProfScope prof_scope("dcgc_gather_sum", (cudaStream_t)stream);
  const bool v4 = width % 4 == 0 && ld_x % 4 == 0 && ld_out % 4 == 0 && aligned16(x) && aligned16(out) &&
                  (addend == nullptr || (ld_add % 4 == 0 && aligned16(addend)));
  cudaStream_t st = (cudaStream_t)stream;
  const int groups = v4 ? width / 4 : width;
  const unsigned grid = grid_for(n_rows_out * groups);
#define DCGC_GS_LAUNCH(V, B) \
  dcgc_launch(gather_sum_kernel<V, B>, grid, kThreads, 0, st, x, ld_x, row_ptr, bk, idx, n_rows_out, groups, width, addend, ld_add, out, ld_out)
  if (v4) { if (deg_count) DCGC_GS_LAUNCH(4, true); else DCGC_GS_LAUNCH(4, false); }
  else { if (deg_count) DCGC_GS_LAUNCH(1, true); else DCGC_GS_LAUNCH(1, false); }
#undef DCGC_GS_LAUNCH
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gather_sum");
  return DCGC_OK;
}

extern "C" int dcgc_gather_sum(const float* x, int64_t ld_x, const int32_t* row_ptr, const int32_t* idx,
                               int64_t n_rows_out, int32_t width, const float* addend, int64_t ld_add,
                               float* out, int64_t ld_out, void* stream) {
  DCGC_CHECK_ARG(row_ptr || n_rows_out == 0, "dcgc_gather_sum: null row_ptr");
  return gather_sum_impl(x, ld_x, row_ptr, nullptr, idx, n_rows_out, width, addend, ld_add, out, ld_out, stream);
}

extern "C" int dcgc_gather_sum_bucketed(const float* x, int64_t ld_x, const int64_t* deg_count_host, const int32_t* idx,
                                        int64_t n_rows_out, int32_t width, const float* addend, int64_t ld_add,
                                        float* out, int64_t ld_out, void* stream) {
  DCGC_CHECK_ARG(deg_count_host, "dcgc_gather_sum_bucketed: null deg_count");
  return gather_sum_impl(x, ld_x, nullptr, deg_count_host, idx, n_rows_out, width, addend, ld_add, out, ld_out, stream);
}

extern "C" int dcgc_pool_fwd(const float* x, int64_t ld_x, const float* scale, const float* shift,
                             const int32_t* row_ptr, const int32_t* col_idx, int64_t n_rows,
                             int32_t width, float* out, int64_t ld_out, uint8_t* arg, int64_t ld_arg,
                             void* stream) {
  DCGC_CHECK_ARG(n_rows >= 0 && width >= 0 && ld_x >= width && ld_out >= width, "dcgc_pool_fwd: bad sizes");
  DCGC_CHECK_ARG((scale == nullptr) == (shift == nullptr), "dcgc_pool_fwd: scale and shift go together");
  DCGC_CHECK_ARG(arg == nullptr || ld_arg >= width, "dcgc_pool_fwd: ld_arg smaller than width");
  if (n_rows == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && row_ptr && out, "dcgc_pool_fwd: null pointer");
  DcgcProfScope prof_scope("dcgc_pool_fwd", (cudaStream_t)stream);
  const bool v4 = width % 4 == 0 && ld_x % 4 == 0 && ld_out % 4 == 0 && aligned16(x) && aligned16(out) &&
                  (arg == nullptr || (ld_arg % 4 == 0 && (reinterpret_cast<uintptr_t>(arg) & 3) == 0)) &&
                  (scale == nullptr || (aligned16(scale) && aligned16(shift)));
  const int groups = v4 ? width / 4 : width;
  const unsigned grid = grid_for(n_rows * groups);
  cudaStream_t st = (cudaStream_t)stream;
#define DCGC_POOL_LAUNCH(V, A, R)                                                              \
  dcgc_launch(pool_fwd_kernel<V, A, R>, grid, kThreads, 0, st, x, ld_x, scale, shift, row_ptr, col_idx, \
                                                      n_rows, groups, width, out, ld_out, arg, ld_arg)
  if (v4) {
    if (scale) { if (arg) DCGC_POOL_LAUNCH(4, true, true); else DCGC_POOL_LAUNCH(4, true, false); }
    else { if (arg) DCGC_POOL_LAUNCH(4, false, true); else DCGC_POOL_LAUNCH(4, false, false); }
  } else {
    if (scale) { if (arg) DCGC_POOL_LAUNCH(1, true, true); else DCGC_POOL_LAUNCH(1, true, false); }
    else { if (arg) DCGC_POOL_LAUNCH(1, false, true); else DCGC_POOL_LAUNCH(1, false, false); }
  }
#undef DCGC_POOL_LAUNCH
  DCGC_CUDA_LAUNCH_CHECK("dcgc_pool_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_pool_bwd(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t ld_arg,
                             const float* scale, const int32_t* t_row_ptr, const int32_t* t_src,
                             const int32_t* t_slot, int64_t n_rows, int32_t width, float* dx,
                             int64_t ld_dx, void* stream) {
  DCGC_CHECK_ARG(n_rows >= 0 && width >= 0 && ld_dy >= width && ld_dx >= width && ld_arg >= width,
                 "dcgc_pool_bwd: bad sizes");
  if (n_rows == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dy && arg && t_row_ptr && dx, "dcgc_pool_bwd: null pointer");
  DcgcProfScope prof_scope("dcgc_pool_bwd", (cudaStream_t)stream);
  const bool v4 = width % 4 == 0 && ld_dy % 4 == 0 && ld_dx % 4 == 0 && ld_arg % 4 == 0 && aligned16(dy) &&
                  aligned16(dx) && (reinterpret_cast<uintptr_t>(arg) & 3) == 0;
  const int groups = v4 ? width / 4 : width;
  const unsigned grid = grid_for(n_rows * groups);
  cudaStream_t st = (cudaStream_t)stream;
#define DCGC_POOLB_LAUNCH(V, A)                                                                   \
  dcgc_launch(pool_bwd_kernel<V, A>, grid, kThreads, 0, st, dy, ld_dy, arg, ld_arg, scale, t_row_ptr, t_src, \
                                                   t_slot, n_rows, groups, width, dx, ld_dx)
  if (v4) { if (scale) DCGC_POOLB_LAUNCH(4, true); else DCGC_POOLB_LAUNCH(4, false); }
  else { if (scale) DCGC_POOLB_LAUNCH(1, true); else DCGC_POOLB_LAUNCH(1, false); }
#undef DCGC_POOLB_LAUNCH
  DCGC_CUDA_LAUNCH_CHECK("dcgc_pool_bwd");
  return DCGC_OK;
}

extern "C" int dcgc_gather_fwd(const float* x, int64_t ld_x, const float* scale, const float* shift,
                               const int32_t* mol_ptr, const int32_t* mol_atoms, int64_t n_segments,
                               int32_t width, int32_t act, float* out, int64_t ld_out, int32_t* argrow,
                               void* stream) {
  return dcgc_gather_fwd_train(x, ld_x, scale, shift, mol_ptr, mol_atoms, n_segments, width, act, out, ld_out, argrow,
                               nullptr, nullptr, nullptr, stream);
}

// dcgc_gather_fwd that also writes the centred raw per-molecule sum / arg-max value (see gather_fwd_kernel); fused
// engine only, declared in common.h
int dcgc_gather_fwd_train(const float* x, int64_t ld_x, const float* scale, const float* shift, const int32_t* mol_ptr,
                          const int32_t* mol_atoms, int64_t n_segments, int32_t width, int32_t act, float* out,
                          int64_t ld_out, int32_t* argrow, const float* mean, float* zc_sum, float* zc_arg,
                          void* stream) {
  DCGC_CHECK_ARG((zc_sum == nullptr) == (zc_arg == nullptr) && (zc_sum == nullptr || (mean && argrow)),
                 "dcgc_gather_fwd_train: zc_sum, zc_arg, mean and argrow go together");
  DCGC_CHECK_ARG((scale == nullptr) == (shift == nullptr), "dcgc_gather_fwd: scale and shift go together");
  DCGC_CHECK_ARG(n_segments >= 0 && width >= 0 && ld_x >= width && ld_out >= 2 * (int64_t)width,
                 "dcgc_gather_fwd: bad sizes");
  DCGC_CHECK_ARG(act >= DCGC_ACT_NONE && act <= DCGC_ACT_TANH, "dcgc_gather_fwd: unknown activation %d", act);
  if (n_segments == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && mol_ptr && out, "dcgc_gather_fwd: null pointer");
  DcgcProfScope prof_scope("dcgc_gather_fwd", (cudaStream_t)stream);
  const bool v4 = width % 4 == 0 && ld_x % 4 == 0 && aligned16(x);
  const int groups = v4 ? width / 4 : width;
  const unsigned grid = grid_for(n_segments * groups);
  cudaStream_t st = (cudaStream_t)stream;
#define DCGC_GATHER_LAUNCH(V, R)                                                                    \
  dcgc_launch(gather_fwd_kernel<V, R>, grid, kThreads, 0, st, x, ld_x, scale, shift, mol_ptr, mol_atoms,      \
                                                     n_segments, groups, width, act, out, ld_out, argrow, mean,  \
                                                     zc_sum, zc_arg)
  if (v4) { if (argrow) DCGC_GATHER_LAUNCH(4, true); else DCGC_GATHER_LAUNCH(4, false); }
  else { if (argrow) DCGC_GATHER_LAUNCH(1, true); else DCGC_GATHER_LAUNCH(1, false); }
#undef DCGC_GATHER_LAUNCH
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gather_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_gather_bwd(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out,
                               const int32_t* argrow, const int32_t* membership, int64_t n_rows,
                               int32_t width, int32_t act, float* dx, int64_t ld_dx, void* stream) {
  DCGC_CHECK_ARG(n_rows >= 0 && width >= 0 && ld_dout >= 2 * (int64_t)width && ld_out >= 2 * (int64_t)width &&
                     ld_dx >= width, "dcgc_gather_bwd: bad sizes");
  DCGC_CHECK_ARG(act >= DCGC_ACT_NONE && act <= DCGC_ACT_TANH, "dcgc_gather_bwd: unknown activation %d", act);
  if (n_rows == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dout && out && argrow && membership && dx, "dcgc_gather_bwd: null pointer");
  DcgcProfScope prof_scope("dcgc_gather_bwd", (cudaStream_t)stream);
  const int groups = (width + 3) / 4;
  const bool v4 = width % 4 == 0 && ld_dout % 4 == 0 && ld_out % 4 == 0 && ld_dx % 4 == 0 && aligned16(dout) &&
                  aligned16(out) && aligned16(dx) && aligned16(argrow);
  if (v4)
    dcgc_launch(gather_bwd_vec_kernel, grid_for(n_rows * groups), kThreads, 0, (cudaStream_t)stream, 
        dout, ld_dout, out, ld_out, argrow, membership, n_rows, groups, width, act, dx, ld_dx);
  else
    dcgc_launch(gather_bwd_kernel<4>, grid_for(n_rows * groups), kThreads, 0, (cudaStream_t)stream, 
        dout, ld_dout, out, ld_out, argrow, membership, n_rows, groups, width, act, dx, ld_dx);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gather_bwd");
  return DCGC_OK;
}

// Fused-engine variant of dcgc_gather_bwd (declared in common.h, not part of the ABI): also writes the
// BatchNorm-backward column sums of dx against z, one row of partials per block; *n_chunks_out = rows written.
// Needs 16-byte aligned rows; returns DCGC_ERR_INVALID otherwise (the caller then takes the two-pass route).
int dcgc_gather_bwd_stats(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const int32_t* argrow,
                          const int32_t* membership, int64_t n_rows, int32_t width, int32_t act, float* dx,
                          int64_t ld_dx, const float* z, int64_t ld_z, int max_chunks, double* part,
                          int32_t* n_chunks_out, const DcgcBnFin* fin, void* stream) {
  DCGC_CHECK_ARG(n_rows > 0 && width > 0 && dout && out && argrow && membership && dx && z && part && n_chunks_out,
                 "dcgc_gather_bwd_stats: bad arguments");
  DCGC_CHECK_ARG(width % 4 == 0 && ld_dout % 4 == 0 && ld_out % 4 == 0 && ld_dx % 4 == 0 && ld_z % 4 == 0 &&
                     aligned16(dout) && aligned16(out) && aligned16(dx) && aligned16(argrow) && aligned16(z),
                 "dcgc_gather_bwd_stats: rows must be 16-byte aligned");
  DcgcProfScope prof_scope("dcgc_gather_bwd", (cudaStream_t)stream);
  int64_t chunks = (n_rows + 127) / 128;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rows = (n_rows + chunks - 1) / chunks;
  rows = (rows + kGbRowLanes - 1) / kGbRowLanes * kGbRowLanes;
  chunks = (n_rows + rows - 1) / rows;
  dim3 grid((unsigned)chunks, (unsigned)((width + 127) / 128));
  dcgc_launch(gather_bwd_stats_kernel, grid, 32 * kGbRowLanes, 0, (cudaStream_t)stream, 
      dout, ld_dout, out, ld_out, argrow, membership, n_rows, width, act, dx, ld_dx, z, ld_z, rows, part,
      fin ? *fin : DcgcBnFin{});
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gather_bwd_stats");
  *n_chunks_out = (int32_t)chunks;
  return DCGC_OK;
}

// Fused-engine pair for the dense layer's BatchNorm + ReLU backward (declared in common.h, not part of the ABI):
// dcgc_dense_bn_sums writes the column-sum partials at molecule level (*n_chunks_out rows), dcgc_gather_bwd_apply writes
// G = relu'(z) * BatchNorm-backward(GraphGather-backward(d fingerprint)) given the finalized coefficients.
int dcgc_dense_bn_sums(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const int32_t* argrow,
                       const int32_t* mol_ptr, int64_t n_segments, int32_t width, int32_t act, const float* zc_sum,
                       const float* zc_arg, const float* mean, double* part, int32_t* n_chunks_out, void* stream) {
  DCGC_CHECK_ARG(n_segments > 0 && width > 0 && dout && out && argrow && mol_ptr && zc_sum && zc_arg && mean && part &&
                     n_chunks_out, "dcgc_dense_bn_sums: bad arguments");
  DcgcProfScope prof_scope("bn_stats_bwd", (cudaStream_t)stream);
  const int64_t chunks = (n_segments + kDbsMols - 1) / kDbsMols;
  dim3 grid((unsigned)chunks, (unsigned)((width + 31) / 32));
  dcgc_launch(dense_bn_sums_kernel, grid, 32 * kDbsLanes, 0, (cudaStream_t)stream, dout, ld_dout, out, ld_out, argrow, mol_ptr, n_segments, width,
                                                             act, zc_sum, zc_arg, mean, part);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_dense_bn_sums");
  *n_chunks_out = (int32_t)chunks;
  return DCGC_OK;
}

int dcgc_gather_bwd_apply(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const int32_t* argrow,
                          const int32_t* membership, int64_t n_rows, int32_t width, int32_t act, const float* z,
                          int64_t ld_z, const float* mean, const float* invstd, const float* coef, float* g,
                          int64_t ld_g, void* stream) {
  DCGC_CHECK_ARG(n_rows > 0 && width > 0 && dout && out && argrow && membership && z && mean && invstd && coef && g,
                 "dcgc_gather_bwd_apply: bad arguments");
  DCGC_CHECK_ARG(width % 4 == 0 && ld_dout % 4 == 0 && ld_out % 4 == 0 && ld_g % 4 == 0 && ld_z % 4 == 0 &&
                     aligned16(dout) && aligned16(out) && aligned16(g) && aligned16(argrow) && aligned16(z) &&
                     aligned16(mean) && aligned16(invstd) && aligned16(coef),
                 "dcgc_gather_bwd_apply: rows must be 16-byte aligned");
  DcgcProfScope prof_scope("dcgc_gather_bwd", (cudaStream_t)stream);
  int64_t chunks = (n_rows + 127) / 128;
  if (chunks > 592) chunks = 592;                  // four waves of one block per SM at most; each walks a row range
  int64_t rows = (n_rows + chunks - 1) / chunks;
  rows = (rows + kGbRowLanes - 1) / kGbRowLanes * kGbRowLanes;
  chunks = (n_rows + rows - 1) / rows;
  dim3 grid((unsigned)chunks, (unsigned)((width + 127) / 128));
  dcgc_launch(gather_bwd_apply_kernel, grid, 32 * kGbRowLanes, 0, (cudaStream_t)stream, 
      dout, ld_dout, out, ld_out, argrow, membership, n_rows, width, act, z, ld_z, mean, invstd, coef, g, ld_g, rows);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_gather_bwd_apply");
  return DCGC_OK;
}

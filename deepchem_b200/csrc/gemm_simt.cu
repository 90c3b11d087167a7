// fp32 (SIMT FFMA) degree-grouped GEMMs for the GraphConv path: the 1e-5 parity mode of
//   K2  y = act([a1|a2] . W[g] + bias[g])            (GraphConv.forward, layers.py:6202-6229)
//   K6  [d1|d2] = grad . W[g]^T                       (dgrad)
//   K6  dW[g] = [a1|a2]_g^T . grad_g, dbias[g]        (wgrad; deterministic two-stage split-K)
// Rows are degree-sorted, so every 128-row tile of the layout slab lies in exactly one degree
// bucket and uses one weight group; there is no per-degree launch and empty buckets cost nothing.
//
// Tiling: CTA tile 128 x BN (BN = 128 or 64) x 16, 256 threads, 8 x (BN/16) accumulators per
// thread, double-buffered shared memory with register prefetch (one barrier per K step).
#include "common.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;
constexpr int LDS_A = BM + 4;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct GemmArgs {
  const float* a1; int64_t ld_a1; int k1;
  const float* a2; int64_t ld_a2; int k2;
  const float* w; int64_t w_group_stride; int64_t ld_w;
  const float* bias; int64_t bias_group_stride;
  int n1, n2;
  float* c1; int64_t ld_c1;
  float* c2; int64_t ld_c2;
  const int32_t* tiles; int64_t n_rows;
  int act;
  int a1_vec, a2_vec, w_vec, c1_vec, c2_vec;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == DCGC_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == DCGC_ACT_TANH) return tanhf(v);
  return v;
}

// masked 4-wide load of p[0..3] where only the first `valid` elements may be touched
__device__ __forceinline__ float4 load4_masked(const float* p, int valid, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && vec) return __ldg(reinterpret_cast<const float4*>(p));
  if (valid > 0) v.x = __ldg(p);
  if (valid > 1) v.y = __ldg(p + 1);
  if (valid > 2) v.z = __ldg(p + 2);
  if (valid > 3) v.w = __ldg(p + 3);
  return v;
}

template <int BN>
__device__ __forceinline__ void mma_tile(const float (*As)[LDS_A], const float (*Bs)[BN + 4], int tx, int ty,
                                         float (&acc)[8][BN / 16]) {
  constexpr int TN = BN / 16;
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    float a[8], b[TN];
    const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
    a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
    if (TN == 8) {
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// y = act([a1|a2] . B + bias), B[k][j] = W[g][k][j] (TRANS_B = false) or W[g][j][k] (true)
template <int BN, bool TRANS_B>
__global__ void __launch_bounds__(NT, 2) gemm_kernel(const GemmArgs p) {
  dcgc_griddep_wait();
  constexpr int TN = BN / 16;
  constexpr int B_IT = BN / 64;  // float4 loads of the B tile per thread
  __shared__ __align__(16) float As[2][BK][LDS_A];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int row0, rows, g;
  if (p.tiles) {
    const int4 t = __ldg(reinterpret_cast<const int4*>(p.tiles) + blockIdx.x);
    row0 = t.x; rows = t.y; g = t.z;
  } else {
    row0 = blockIdx.x * BM;
    rows = (int)min((int64_t)BM, p.n_rows - row0);
    g = 0;
  }
  const int n0 = blockIdx.y * BN;
  const int N = p.n1 + p.n2;
  const float* W = p.w + (int64_t)g * p.w_group_stride;
  const int chunks1 = (p.k1 + BK - 1) / BK, chunks2 = (p.k2 + BK - 1) / BK;
  const int total = chunks1 + chunks2;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float4 ra[2], rb[B_IT];

  auto gload = [&](int ch) {
    const float* src; int64_t ld; int ksrc, kbase, wrow; bool vec;
    if (ch < chunks1) { src = p.a1; ld = p.ld_a1; ksrc = p.k1; kbase = ch * BK; wrow = kbase; vec = p.a1_vec; }
    else { src = p.a2; ld = p.ld_a2; ksrc = p.k2; kbase = (ch - chunks1) * BK; wrow = p.k1 + kbase; vec = p.a2_vec; }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int f = tid + it * NT, r = f >> 2, k = kbase + 4 * (f & 3);
      ra[it] = r < rows ? load4_masked(src + (int64_t)(row0 + r) * ld + k, ksrc - k, vec)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!TRANS_B) {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        const int f = tid + it * NT, kk = f / (BN / 4), n = n0 + 4 * (f % (BN / 4));
        rb[it] = (kbase + kk < ksrc) ? load4_masked(W + (int64_t)(wrow + kk) * p.ld_w + n, N - n, p.w_vec)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        const int f = tid + it * NT, j = n0 + (f >> 2), k = kbase + 4 * (f & 3);
        rb[it] = j < N ? load4_masked(W + (int64_t)j * p.ld_w + k, ksrc - k, p.w_vec)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int f = tid + it * NT, r = f >> 2, kq = 4 * (f & 3);
      As[buf][kq + 0][r] = ra[it].x; As[buf][kq + 1][r] = ra[it].y;
      As[buf][kq + 2][r] = ra[it].z; As[buf][kq + 3][r] = ra[it].w;
    }
    if (!TRANS_B) {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        const int f = tid + it * NT, kk = f / (BN / 4), nq = 4 * (f % (BN / 4));
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb[it];
      }
    } else {
#pragma unroll
      for (int it = 0; it < B_IT; ++it) {
        const int f = tid + it * NT, jj = f >> 2, kq = 4 * (f & 3);
        Bs[buf][kq + 0][jj] = rb[it].x; Bs[buf][kq + 1][jj] = rb[it].y;
        Bs[buf][kq + 2][jj] = rb[it].z; Bs[buf][kq + 3][jj] = rb[it].w;
      }
    }
  };

  if (total > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int ch = 0; ch < total; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < total) gload(ch + 1);
    mma_tile<BN>(As[buf], Bs[buf], tx, ty, acc);
    if (ch + 1 < total) sstore(buf ^ 1);
    __syncthreads();
  }

  const float* bias = p.bias ? p.bias + (int64_t)g * p.bias_group_stride : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    if (r >= rows) continue;
    const int64_t grow = row0 + r;
#pragma unroll
    for (int jb = 0; jb < TN / 4; ++jb) {
      const int cb = n0 + jb * 64 + tx * 4;
      if (cb >= N) continue;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = acc[i][jb * 4 + e];
        if (bias && cb + e < N) t += __ldg(bias + cb + e);
        v[e] = act_apply(t, p.act);
      }
      if (cb + 3 < p.n1 && p.c1_vec) {
        *reinterpret_cast<float4*>(p.c1 + grow * p.ld_c1 + cb) = make_float4(v[0], v[1], v[2], v[3]);
      } else if (cb >= p.n1 && cb + 3 < N && p.c2_vec && ((cb - p.n1) & 3) == 0) {
        *reinterpret_cast<float4*>(p.c2 + grow * p.ld_c2 + (cb - p.n1)) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = cb + e;
          if (c < p.n1) { if (p.c1) p.c1[grow * p.ld_c1 + c] = v[e]; }
          else if (c < N) { if (p.c2) p.c2[grow * p.ld_c2 + (c - p.n1)] = v[e]; }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// wgrad: stage 1 — every CTA reduces one chunk of rows of one group into a 128 x BN partial of
// dW (and of dbias); stage 2 — partials of a group are summed in chunk order.
// ------------------------------------------------------------------------------------------
using WgradArgs = DcgcWgradArgs;

template <int BN>
__global__ void __launch_bounds__(NT, 2) wgrad_kernel(const WgradArgs p) {
  dcgc_griddep_wait();
  constexpr int TN = BN / 16;
  constexpr int B_IT = BN / 64;
  __shared__ __align__(16) float As[2][BK][LDS_A];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int chunk = blockIdx.x;
  int g = 0;
  while (g + 1 < p.n_groups && chunk >= p.chunk_prefix[g + 1]) ++g;
  const int64_t r_begin = p.group_row0[g] + (int64_t)(chunk - p.chunk_prefix[g]) * p.chunk_rows;
  const int64_t r_end = min(p.group_row0[g + 1], r_begin + p.chunk_rows);
  const int mt = blockIdx.y / p.tiles_n, nt = blockIdx.y - mt * p.tiles_n;
  const int m0 = mt * BM, n0 = nt * BN;
  const int Kt = p.k1 + p.k2;
  const int steps = (int)((r_end - r_begin + BK - 1) / BK);

  float acc[8][TN], bsum[TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < TN; ++j) bsum[j] = 0.f;
  float4 ra[2], rb[B_IT];

  auto gload = [&](int s) {
    const int64_t rbase = r_begin + (int64_t)s * BK;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int f = tid + it * NT, kk = f >> 5, feat = m0 + 4 * (f & 31);
      const int64_t r = rbase + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < r_end && feat < Kt) {
        if (feat + 3 < p.k1) {
          v = load4_masked(p.a1 + r * p.ld_a1 + feat, 4, p.a1_vec);
        } else if (feat >= p.k1) {
          const int f2 = feat - p.k1;
          v = load4_masked(p.a2 + r * p.ld_a2 + f2, p.k2 - f2, p.a2_vec && (f2 & 3) == 0);
        } else {  // group straddles the a1 | a2 boundary
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int fe = feat + q;
            e[q] = fe < p.k1 ? __ldg(p.a1 + r * p.ld_a1 + fe)
                             : (fe < Kt ? __ldg(p.a2 + r * p.ld_a2 + (fe - p.k1)) : 0.f);
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      ra[it] = v;
    }
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
      const int f = tid + it * NT, kk = f / (BN / 4), n = n0 + 4 * (f % (BN / 4));
      const int64_t r = rbase + kk;
      rb[it] = r < r_end ? load4_masked(p.g + r * p.ld_g + n, p.n - n, p.g_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int f = tid + it * NT;
      *reinterpret_cast<float4*>(&As[buf][f >> 5][4 * (f & 31)]) = ra[it];
    }
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
      const int f = tid + it * NT;
      *reinterpret_cast<float4*>(&Bs[buf][f / (BN / 4)][4 * (f % (BN / 4))]) = rb[it];
    }
  };

  if (steps > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int s = 0; s < steps; ++s) {
    const int buf = s & 1;
    if (s + 1 < steps) gload(s + 1);
    mma_tile<BN>(As[buf], Bs[buf], tx, ty, acc);
    if (mt == 0 && ty == 0) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
#pragma unroll
        for (int jb = 0; jb < TN / 4; ++jb) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][jb * 64 + tx * 4]);
          bsum[jb * 4 + 0] += b.x; bsum[jb * 4 + 1] += b.y; bsum[jb * 4 + 2] += b.z; bsum[jb * 4 + 3] += b.w;
        }
      }
    }
    if (s + 1 < steps) sstore(buf ^ 1);
    __syncthreads();
  }

  float* ws = p.ws + (int64_t)chunk * Kt * p.n;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= Kt) continue;
#pragma unroll
    for (int jb = 0; jb < TN / 4; ++jb) {
      const int cb = n0 + jb * 64 + tx * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (cb + e < p.n) ws[(int64_t)m * p.n + cb + e] = acc[i][jb * 4 + e];
    }
  }
  if (mt == 0 && ty == 0) {
#pragma unroll
    for (int jb = 0; jb < TN / 4; ++jb)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = n0 + jb * 64 + tx * 4 + e;
        if (c < p.n) p.wsb[(int64_t)chunk * p.n + c] = bsum[jb * 4 + e];
      }
  }
}

struct ReduceArgs {
  const float* ws; const float* wsb;
  float* dw; float* dbias;
  float* dbias21;  // optional: the same sums scattered to GraphConv's 21 bias rows (row 20 <- group 0, rows 2(g-1), 2(g-1)+1 <- group g)
  int64_t kn;  // (k1+k2)*n
  int n, n_groups;
  int transpose;  // write dw[g][j][k] instead of dw[g][k][j] (nn.Linear weight layout)
  int chunk_prefix[DCGC_N_DEG + 1];
};

__global__ void __launch_bounds__(NT) wgrad_reduce_kernel(const ReduceArgs p) {
  dcgc_griddep_wait();
  const int g = blockIdx.y;
  const int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (idx >= p.kn + p.n) return;
  const int c0 = p.chunk_prefix[g], c1 = p.chunk_prefix[g + 1];
  float s = 0.f;
  if (idx < p.kn) {
    // chunk order is fixed (deterministic); 8 independent loads in flight per thread instead of one
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(p.ws + (int64_t)(c + u) * p.kn + idx);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; c < c1; ++c) s += __ldg(p.ws + (int64_t)c * p.kn + idx);
    int64_t o = idx;
    if (p.transpose) {
      const int64_t k = idx / p.n, j = idx - k * p.n;
      o = j * (p.kn / p.n) + k;
    }
    p.dw[(int64_t)g * p.kn + o] = s;
  } else if (p.dbias) {
    const int64_t j = idx - p.kn;
    for (int c = c0; c < c1; ++c) s += __ldg(p.wsb + (int64_t)c * p.n + j);
    p.dbias[(int64_t)g * p.n + j] = s;
    if (p.dbias21) {
      if (g == 0) {
        p.dbias21[(int64_t)20 * p.n + j] = s;
      } else {
        p.dbias21[(int64_t)(2 * (g - 1)) * p.n + j] = s;
        p.dbias21[(int64_t)(2 * (g - 1) + 1) * p.n + j] = s;
      }
    }
  }
}

constexpr int kTargetChunks = 148;

}  // namespace

static int group_gemm_fwd_impl(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2,
                               int64_t ld_a2, int32_t k2, const float* w, const float* bias, int32_t n,
                               const int32_t* tiles, int64_t n_tiles, int32_t tile_rows, int64_t n_rows,
                               int32_t act, float* y, int64_t ld_y, double* stats, int32_t* stats_chunks,
                               void* stream, const DcgcGemmOpts* opts = nullptr) {
  if (stats_chunks) *stats_chunks = 0;
  DCGC_CHECK_ARG(stats == nullptr || dcgc_tc_terms(mode) != 0,
                 "dcgc_group_gemm_fwd_stats: fused column statistics exist in DCGC_GEMM_TF32X3 mode only");
  DCGC_CHECK_ARG(mode == DCGC_GEMM_FP32 || dcgc_tc_terms(mode) != 0,
                 "dcgc_group_gemm_fwd: GEMM mode %d is not available in this build", mode);
  DCGC_CHECK_ARG(k1 >= 0 && k2 >= 0 && n >= 0 && n_rows >= 0 && ld_a1 >= k1 && ld_y >= n,
                 "dcgc_group_gemm_fwd: bad sizes");
  DCGC_CHECK_ARG(a2 != nullptr || k2 == 0, "dcgc_group_gemm_fwd: a2 is null but k2 > 0");
  DCGC_CHECK_ARG(a2 == nullptr || ld_a2 >= k2, "dcgc_group_gemm_fwd: ld_a2 smaller than k2");
  DCGC_CHECK_ARG(act >= DCGC_ACT_NONE && act <= DCGC_ACT_TANH, "dcgc_group_gemm_fwd: unknown activation %d", act);
  DCGC_CHECK_ARG(tiles == nullptr || tile_rows == BM, "dcgc_group_gemm_fwd: tile_rows must be %d", BM);
  if (n_rows == 0 || n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(a1 && w && y, "dcgc_group_gemm_fwd: null pointer");
  DcgcProfScope prof_scope("dcgc_group_gemm_fwd", (cudaStream_t)stream);
  if (dcgc_tc_terms(mode)) {
    DcgcGemmOpts o2 = opts ? *opts : DcgcGemmOpts{};
    if (mode == DCGC_GEMM_F16X3) o2.f16x3 = 1;
    return dcgc_tc_gemm(dcgc_tc_terms(mode), a1, ld_a1, k1, a2, ld_a2, a2 ? k2 : 0, w, tiles ? DCGC_N_DEG : 1, 1, bias, n, 0, tiles, n_tiles,
                        n_rows, act, y, ld_y, nullptr, 0, (cudaStream_t)stream, stats, stats_chunks, &o2);
  }
  GemmArgs p{};
  p.a1 = a1; p.ld_a1 = ld_a1; p.k1 = k1;
  p.a2 = a2; p.ld_a2 = ld_a2; p.k2 = a2 ? k2 : 0;
  p.w = w; p.w_group_stride = (int64_t)(k1 + k2) * n; p.ld_w = n;
  p.bias = bias; p.bias_group_stride = n;
  p.n1 = n; p.n2 = 0;
  p.c1 = y; p.ld_c1 = ld_y; p.c2 = nullptr; p.ld_c2 = 0;
  p.tiles = tiles; p.n_rows = n_rows; p.act = act;
  p.a1_vec = ld_a1 % 4 == 0 && aligned16(a1);
  p.a2_vec = a2 && ld_a2 % 4 == 0 && aligned16(a2);
  p.w_vec = n % 4 == 0 && aligned16(w);
  p.c1_vec = ld_y % 4 == 0 && aligned16(y);
  p.c2_vec = 0;
  const int64_t row_tiles = tiles ? n_tiles : (n_rows + BM - 1) / BM;
  if (row_tiles == 0) return DCGC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 64) {
    dim3 grid((unsigned)row_tiles, (unsigned)((n + 127) / 128));
    dcgc_launch(gemm_kernel<128, false>, grid, NT, 0, st, p);
  } else {
    dim3 grid((unsigned)row_tiles, 1);
    dcgc_launch(gemm_kernel<64, false>, grid, NT, 0, st, p);
  }
  DCGC_CUDA_LAUNCH_CHECK("dcgc_group_gemm_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_group_gemm_fwd(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2,
                                   int64_t ld_a2, int32_t k2, const float* w, const float* bias, int32_t n,
                                   const int32_t* tiles, int64_t n_tiles, int32_t tile_rows, int64_t n_rows,
                                   int32_t act, float* y, int64_t ld_y, void* stream) {
  return group_gemm_fwd_impl(mode, a1, ld_a1, k1, a2, ld_a2, k2, w, bias, n, tiles, n_tiles, tile_rows, n_rows, act, y,
                             ld_y, nullptr, nullptr, stream);
}

extern "C" int dcgc_group_gemm_fwd_stats(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2,
                                         int64_t ld_a2, int32_t k2, const float* w, const float* bias, int32_t n,
                                         const int32_t* tiles, int64_t n_tiles, int32_t tile_rows, int64_t n_rows,
                                         int32_t act, float* y, int64_t ld_y, double* stats_part,
                                         int32_t* n_chunks_out, void* stream) {
  DCGC_CHECK_ARG(stats_part && n_chunks_out, "dcgc_group_gemm_fwd_stats: null statistics buffer");
  return group_gemm_fwd_impl(mode, a1, ld_a1, k1, a2, ld_a2, k2, w, bias, n, tiles, n_tiles, tile_rows, n_rows, act, y,
                             ld_y, stats_part, n_chunks_out, stream);
}

extern "C" int32_t dcgc_gemm_stats_max_chunks(void) { return dcgc_tc_num_sms(); }

static int group_gemm_dgrad_impl(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w,
                                 int32_t k1, int32_t k2, const int32_t* tiles, int64_t n_tiles,
                                 int32_t tile_rows, int64_t n_rows, float* d1, int64_t ld_d1, float* d2,
                                 int64_t ld_d2, void* stream, const DcgcGemmOpts* opts) {
  DCGC_CHECK_ARG(mode == DCGC_GEMM_FP32 || dcgc_tc_terms(mode) != 0,
                 "dcgc_group_gemm_dgrad: GEMM mode %d is not available in this build", mode);
  DCGC_CHECK_ARG(k1 >= 0 && k2 >= 0 && n >= 0 && n_rows >= 0 && ld_g >= n, "dcgc_group_gemm_dgrad: bad sizes");
  DCGC_CHECK_ARG((d1 == nullptr || ld_d1 >= k1) && (d2 == nullptr || ld_d2 >= k2),
                 "dcgc_group_gemm_dgrad: output leading dimension too small");
  DCGC_CHECK_ARG(tiles == nullptr || tile_rows == BM, "dcgc_group_gemm_dgrad: tile_rows must be %d", BM);
  if (n_rows == 0 || k1 + k2 == 0 || (!d1 && !d2)) return DCGC_OK;
  DCGC_CHECK_ARG(g && w, "dcgc_group_gemm_dgrad: null pointer");
  DcgcProfScope prof_scope("dcgc_group_gemm_dgrad", (cudaStream_t)stream);
  if (dcgc_tc_terms(mode))
    return dcgc_tc_gemm(dcgc_tc_terms(mode), g, ld_g, n, nullptr, 0, 0, w, tiles ? DCGC_N_DEG : 1, 0, nullptr, k1, k2, tiles, n_tiles,
                        n_rows, DCGC_ACT_NONE, d1, ld_d1, d2, ld_d2, (cudaStream_t)stream, nullptr, nullptr, opts);
  GemmArgs p{};
  p.a1 = g; p.ld_a1 = ld_g; p.k1 = n;
  p.a2 = nullptr; p.ld_a2 = 0; p.k2 = 0;
  p.w = w; p.w_group_stride = (int64_t)(k1 + k2) * n; p.ld_w = n;
  p.bias = nullptr; p.bias_group_stride = 0;
  p.n1 = k1; p.n2 = k2;
  p.c1 = d1; p.ld_c1 = ld_d1; p.c2 = d2; p.ld_c2 = ld_d2;
  p.tiles = tiles; p.n_rows = n_rows; p.act = DCGC_ACT_NONE;
  p.a1_vec = ld_g % 4 == 0 && aligned16(g);
  p.a2_vec = 0;
  p.w_vec = n % 4 == 0 && aligned16(w);
  p.c1_vec = d1 && ld_d1 % 4 == 0 && aligned16(d1);
  p.c2_vec = d2 && ld_d2 % 4 == 0 && aligned16(d2);
  const int64_t row_tiles = tiles ? n_tiles : (n_rows + BM - 1) / BM;
  if (row_tiles == 0) return DCGC_OK;
  const int N = k1 + k2;
  cudaStream_t st = (cudaStream_t)stream;
  if (N > 64) {
    dim3 grid((unsigned)row_tiles, (unsigned)((N + 127) / 128));
    dcgc_launch(gemm_kernel<128, true>, grid, NT, 0, st, p);
  } else {
    dim3 grid((unsigned)row_tiles, 1);
    dcgc_launch(gemm_kernel<64, true>, grid, NT, 0, st, p);
  }
  DCGC_CUDA_LAUNCH_CHECK("dcgc_group_gemm_dgrad");
  return DCGC_OK;
}

extern "C" int dcgc_group_gemm_dgrad(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w,
                                     int32_t k1, int32_t k2, const int32_t* tiles, int64_t n_tiles,
                                     int32_t tile_rows, int64_t n_rows, float* d1, int64_t ld_d1, float* d2,
                                     int64_t ld_d2, void* stream) {
  return group_gemm_dgrad_impl(mode, g, ld_g, n, w, k1, k2, tiles, n_tiles, tile_rows, n_rows, d1, ld_d1, d2, ld_d2,
                               stream, nullptr);
}
int dcgc_group_gemm_dgrad_opts(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w, int32_t k1,
                               int32_t k2, const int32_t* tiles, int64_t n_tiles, int32_t tile_rows, int64_t n_rows,
                               float* d1, int64_t ld_d1, float* d2, int64_t ld_d2, const DcgcGemmOpts& opts,
                               void* stream) {
  return group_gemm_dgrad_impl(mode, g, ld_g, n, w, k1, k2, tiles, n_tiles, tile_rows, n_rows, d1, ld_d1, d2, ld_d2,
                               stream, &opts);
}
int dcgc_group_gemm_fwd_opts(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2, int64_t ld_a2,
                             int32_t k2, const float* w, const float* bias, int32_t n, const int32_t* tiles,
                             int64_t n_tiles, int32_t tile_rows, int64_t n_rows, int32_t act, float* y, int64_t ld_y,
                             double* stats_part, int32_t* n_chunks_out, const DcgcGemmOpts& opts, void* stream) {
  return group_gemm_fwd_impl(mode, a1, ld_a1, k1, a2, ld_a2, k2, w, bias, n, tiles, n_tiles, tile_rows, n_rows, act, y,
                             ld_y, stats_part, n_chunks_out, stream, &opts);
}

extern "C" int64_t dcgc_group_gemm_wgrad_workspace(int32_t k1, int32_t k2, int32_t n, int32_t n_groups) {
  if (k1 < 0 || k2 < 0 || n < 0 || n_groups < 1 || n_groups > DCGC_N_DEG) return 0;
  const int64_t chunks = kTargetChunks + n_groups;
  return chunks * ((int64_t)(k1 + k2) * n + n) * 4 + 256;
}

static int wgrad_impl(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2,
                      int64_t ld_a2, int32_t k2, const float* g, int64_t ld_g, int32_t n,
                      const int64_t* deg_count, int32_t n_groups, float* dw, float* dbias,
                      void* workspace, int64_t workspace_bytes, int transpose, void* stream, int a_exact = 0,
                      float* dbias21 = nullptr) {
  DCGC_CHECK_ARG(mode == DCGC_GEMM_FP32 || dcgc_tc_terms(mode) != 0,
                 "dcgc_group_gemm_wgrad: GEMM mode %d is not available in this build", mode);
  DCGC_CHECK_ARG(k1 >= 0 && k2 >= 0 && n >= 0 && ld_a1 >= k1 && ld_g >= n, "dcgc_group_gemm_wgrad: bad sizes");
  DCGC_CHECK_ARG(n_groups >= 1 && n_groups <= DCGC_N_DEG && deg_count, "dcgc_group_gemm_wgrad: bad groups");
  DCGC_CHECK_ARG(a2 != nullptr || k2 == 0, "dcgc_group_gemm_wgrad: a2 is null but k2 > 0");
  DCGC_CHECK_ARG(a2 == nullptr || ld_a2 >= k2, "dcgc_group_gemm_wgrad: ld_a2 smaller than k2");
  const int Kt = k1 + k2;
  if (Kt == 0 || n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dw, "dcgc_group_gemm_wgrad: null dw");
  int64_t total_rows = 0;
  for (int i = 0; i < n_groups; ++i) {
    DCGC_CHECK_ARG(deg_count[i] >= 0, "dcgc_group_gemm_wgrad: negative group size");
    total_rows += deg_count[i];
  }
  WgradArgs p{};
  p.a1 = a1; p.ld_a1 = ld_a1; p.k1 = k1;
  p.a2 = a2; p.ld_a2 = ld_a2; p.k2 = a2 ? k2 : 0;
  p.g = g; p.ld_g = ld_g; p.n = n;
  const bool tc = dcgc_tc_terms(mode) != 0;
  int64_t target = kTargetChunks;
  if (tc) {
    // one CTA per SM (the kernel takes the whole shared memory): keep the grid within one wave;
    // every non-empty group adds at most one ragged chunk
    int nonempty = 0;
    for (int i = 0; i < n_groups; ++i) nonempty += deg_count[i] > 0;
    target = dcgc_tc_num_sms() / dcgc_tc_wgrad_grid_y(Kt, n) - nonempty;
    if (target > kTargetChunks) target = kTargetChunks;
    if (target < 1) target = 1;
  }
  int64_t chunk_rows = (total_rows + target - 1) / target;
  const int64_t gran = tc ? 32 : BK;
  chunk_rows = (chunk_rows + gran - 1) / gran * gran;
  if (chunk_rows < 4 * BK) chunk_rows = 4 * BK;
  p.chunk_rows = (int)chunk_rows;
  p.n_groups = n_groups;
  ReduceArgs q{};
  int64_t row = 0;
  int chunks = 0;
  for (int i = 0; i < n_groups; ++i) {
    p.group_row0[i] = row;
    p.chunk_prefix[i] = q.chunk_prefix[i] = chunks;
    row += deg_count[i];
    chunks += (int)((deg_count[i] + chunk_rows - 1) / chunk_rows);
  }
  for (int i = n_groups; i <= DCGC_N_DEG; ++i) {
    p.group_row0[i] = row;
    p.chunk_prefix[i] = q.chunk_prefix[i] = chunks;
  }
  const int64_t kn = (int64_t)Kt * n;
  const int64_t need = (int64_t)chunks * (kn + n) * 4;
  if (need > workspace_bytes || (chunks > 0 && workspace == nullptr)) {
    dcgc_set_error("dcgc_group_gemm_wgrad: workspace of %lld bytes needed, %lld given", (long long)need,
                   (long long)workspace_bytes);
    return DCGC_ERR_NOMEM;
  }
  DCGC_CHECK_ARG(chunks == 0 || (a1 && g), "dcgc_group_gemm_wgrad: null pointer");
  p.ws = (float*)workspace;
  p.wsb = p.ws + (int64_t)chunks * kn;
  p.a1_vec = ld_a1 % 4 == 0 && aligned16(a1);
  p.a2_vec = a2 && ld_a2 % 4 == 0 && aligned16(a2);
  p.g_vec = ld_g % 4 == 0 && aligned16(g);
  p.a_exact = (a_exact && mode == DCGC_GEMM_TF32X3) ? 1 : 0;
  DcgcProfScope prof_scope("dcgc_group_gemm_wgrad", (cudaStream_t)stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (chunks > 0 && tc) {
    int st_ = dcgc_tc_wgrad_stage1(dcgc_tc_terms(mode), p, chunks, st);
    if (st_ != DCGC_OK) return st_;
  } else if (chunks > 0) {
    const int tiles_m = (Kt + BM - 1) / BM;
    if (n > 64) {
      p.tiles_n = (n + 127) / 128;
      dim3 grid((unsigned)chunks, (unsigned)(tiles_m * p.tiles_n));
      dcgc_launch(wgrad_kernel<128>, grid, NT, 0, st, p);
    } else {
      p.tiles_n = 1;
      dim3 grid((unsigned)chunks, (unsigned)tiles_m);
      dcgc_launch(wgrad_kernel<64>, grid, NT, 0, st, p);
    }
    DCGC_CUDA_LAUNCH_CHECK("dcgc_group_gemm_wgrad (stage 1)");
  }
  q.ws = p.ws; q.wsb = p.wsb; q.dw = dw; q.dbias = dbias; q.kn = kn; q.n = n; q.n_groups = n_groups;
  q.transpose = transpose;
  q.dbias21 = (dbias && n_groups == DCGC_N_DEG) ? dbias21 : nullptr;
  dim3 rgrid((unsigned)((kn + n + NT - 1) / NT), (unsigned)n_groups);
  dcgc_launch(wgrad_reduce_kernel, rgrid, NT, 0, st, q);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_group_gemm_wgrad (stage 2)");
  return DCGC_OK;
}

extern "C" int dcgc_group_gemm_wgrad(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2,
                                     int64_t ld_a2, int32_t k2, const float* g, int64_t ld_g, int32_t n,
                                     const int64_t* deg_count, int32_t n_groups, float* dw, float* dbias,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  return wgrad_impl(mode, a1, ld_a1, k1, a2, ld_a2, k2, g, ld_g, n, deg_count, n_groups, dw, dbias, workspace,
                    workspace_bytes, 0, stream);
}
int dcgc_group_gemm_wgrad_opts(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2, int64_t ld_a2,
                               int32_t k2, const float* g, int64_t ld_g, int32_t n, const int64_t* deg_count,
                               int32_t n_groups, float* dw, float* dbias, void* workspace, int64_t workspace_bytes,
                               int a_exact, void* stream, float* dbias21) {
  return wgrad_impl(mode, a1, ld_a1, k1, a2, ld_a2, k2, g, ld_g, n, deg_count, n_groups, dw, dbias, workspace,
                    workspace_bytes, 0, stream, a_exact, dbias21);
}

// ------------------------------------------------------------------------------------------
// nn.Linear-layout helpers (weight stored [n_out, k_in] as torch does): the atom-level Dense
// layer (graphconvmodel.py:172,222) and the DMPNN W_i / W_h / W_o (layers.py:1510-1517).
// ------------------------------------------------------------------------------------------
static int linear_fwd_impl(int32_t mode, const float* x, int64_t ld_x, int32_t k, const float* w, const float* bias,
                           int32_t n, int64_t n_rows, int32_t act, float* y, int64_t ld_y, double* stats,
                           int32_t* stats_chunks, void* stream, const DcgcGemmOpts* opts = nullptr) {
  if (stats_chunks) *stats_chunks = 0;
  DCGC_CHECK_ARG(stats == nullptr || dcgc_tc_terms(mode) != 0,
                 "dcgc_linear_fwd_stats: fused column statistics exist in DCGC_GEMM_TF32X3 mode only");
  DCGC_CHECK_ARG(mode == DCGC_GEMM_FP32 || dcgc_tc_terms(mode) != 0,
                 "dcgc_linear_fwd: GEMM mode %d is not available in this build", mode);
  DCGC_CHECK_ARG(k >= 0 && n >= 0 && n_rows >= 0 && ld_x >= k && ld_y >= n, "dcgc_linear_fwd: bad sizes");
  DCGC_CHECK_ARG(act >= DCGC_ACT_NONE && act <= DCGC_ACT_TANH, "dcgc_linear_fwd: unknown activation %d", act);
  if (n_rows == 0 || n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && w && y, "dcgc_linear_fwd: null pointer");
  DcgcProfScope prof_scope("dcgc_linear_fwd", (cudaStream_t)stream);
  if (dcgc_tc_terms(mode)) {
    DcgcGemmOpts o2 = opts ? *opts : DcgcGemmOpts{};
    if (mode == DCGC_GEMM_F16X3) o2.f16x3 = 1;
    return dcgc_tc_gemm(dcgc_tc_terms(mode), x, ld_x, k, nullptr, 0, 0, w, 1, 0, bias, n, 0, nullptr, 0, n_rows, act, y, ld_y, nullptr, 0,
                        (cudaStream_t)stream, stats, stats_chunks, &o2);
  }
  GemmArgs p{};
  p.a1 = x; p.ld_a1 = ld_x; p.k1 = k;
  p.w = w; p.w_group_stride = 0; p.ld_w = k;
  p.bias = bias; p.bias_group_stride = 0;
  p.n1 = n; p.n2 = 0;
  p.c1 = y; p.ld_c1 = ld_y;
  p.tiles = nullptr; p.n_rows = n_rows; p.act = act;
  p.a1_vec = ld_x % 4 == 0 && aligned16(x);
  p.w_vec = k % 4 == 0 && aligned16(w);
  p.c1_vec = ld_y % 4 == 0 && aligned16(y);
  const int64_t row_tiles = (n_rows + BM - 1) / BM;
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 64) {
    dim3 grid((unsigned)row_tiles, (unsigned)((n + 127) / 128));
    dcgc_launch(gemm_kernel<128, true>, grid, NT, 0, st, p);
  } else {
    dim3 grid((unsigned)row_tiles, 1);
    dcgc_launch(gemm_kernel<64, true>, grid, NT, 0, st, p);
  }
  DCGC_CUDA_LAUNCH_CHECK("dcgc_linear_fwd");
  return DCGC_OK;
}

extern "C" int dcgc_linear_fwd(int32_t mode, const float* x, int64_t ld_x, int32_t k, const float* w,
                               const float* bias, int32_t n, int64_t n_rows, int32_t act, float* y, int64_t ld_y,
                               void* stream) {
  return linear_fwd_impl(mode, x, ld_x, k, w, bias, n, n_rows, act, y, ld_y, nullptr, nullptr, stream);
}

extern "C" int dcgc_linear_fwd_stats(int32_t mode, const float* x, int64_t ld_x, int32_t k, const float* w,
                                     const float* bias, int32_t n, int64_t n_rows, int32_t act, float* y,
                                     int64_t ld_y, double* stats_part, int32_t* n_chunks_out, void* stream) {
  DCGC_CHECK_ARG(stats_part && n_chunks_out, "dcgc_linear_fwd_stats: null statistics buffer");
  return linear_fwd_impl(mode, x, ld_x, k, w, bias, n, n_rows, act, y, ld_y, stats_part, n_chunks_out, stream);
}

int dcgc_linear_fwd_opts(int32_t mode, const float* x, int64_t ld_x, int32_t k, const float* w, const float* bias,
                         int32_t n, int64_t n_rows, int32_t act, float* y, int64_t ld_y, double* stats_part,
                         int32_t* n_chunks_out, const DcgcGemmOpts& opts, void* stream) {
  return linear_fwd_impl(mode, x, ld_x, k, w, bias, n, n_rows, act, y, ld_y, stats_part, n_chunks_out, stream, &opts);
}
int dcgc_linear_dgrad_opts(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w, int32_t k,
                           int64_t n_rows, float* dx, int64_t ld_dx, const DcgcGemmOpts& opts, void* stream) {
  return group_gemm_fwd_impl(mode, g, ld_g, n, nullptr, 0, 0, w, nullptr, k, nullptr, 0, BM, n_rows, DCGC_ACT_NONE, dx,
                             ld_dx, nullptr, nullptr, stream, &opts);
}

extern "C" int dcgc_linear_dgrad(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w, int32_t k,
                                 int64_t n_rows, float* dx, int64_t ld_dx, void* stream) {
  // dx = g . w with w [n, k]: a plain (non-transposed) product with K = n, N = k
  return dcgc_group_gemm_fwd(mode, g, ld_g, n, nullptr, 0, 0, w, nullptr, k, nullptr, 0, BM, n_rows, DCGC_ACT_NONE,
                             dx, ld_dx, stream);
}

extern "C" int64_t dcgc_linear_wgrad_workspace(int32_t k, int32_t n) {
  return dcgc_group_gemm_wgrad_workspace(k, 0, n, 1);
}

extern "C" int dcgc_linear_wgrad(int32_t mode, const float* x, int64_t ld_x, int32_t k, const float* g, int64_t ld_g,
                                 int32_t n, int64_t n_rows, float* dw, float* dbias, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  // dw [n, k] = g^T . x ; dbias [n] = column sums of g
  const int64_t counts[1] = {n_rows};
  return wgrad_impl(mode, x, ld_x, k, nullptr, 0, 0, g, ld_g, n, counts, 1, dw, dbias, workspace, workspace_bytes, 1,
                    stream);
}

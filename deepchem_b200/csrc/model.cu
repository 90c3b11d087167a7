// Whole-model engine for GraphConvModel: one C call runs the forward, the loss and the full
// backward of
//   [GraphConv -> ReLU -> BatchNorm -> GraphPool] x L -> Dense -> ReLU -> BatchNorm ->
//   GraphGather(tanh) -> Linear head -> (softmax) -> weighted mean loss
// (deepchem/models/torch_models/graphconvmodel.py:188-249, torch_model.py:435-443,1275-1294)
// over a flat fp32 parameter slab, writing every gradient into a flat gradient slab of the same
// layout (one NCCL all-reduce per step in data parallel, one fused Adam launch).
//
// Fusions relative to the layer-by-layer graph:
//   * BatchNorm's normalise+affine is folded into the loads of the consumer (GraphPool max /
//     GraphGather) as a per-channel scale/shift: the normalised tensor is never materialised;
//   * BatchNorm backward + ReLU backward are one elementwise pass producing the GEMM gradient;
//   * per-channel statistics (forward: sum y, sum y^2; backward: sum dA, sum dA*y) are two-stage
//     deterministic column reductions with float64 accumulation (no float atomics);
//   * the self-path and neighbour-path input gradients are added inside the transposed gather.
#include <math.h>

#include <stdlib.h>

#include <map>
#include <mutex>

#include "common.h"

namespace {

constexpr int kT = 256;

inline unsigned blocks_for(int64_t n, int t = kT) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------
// column moments: out_a[c] = sum_r a[r,c], out_ab[c] = sum_r a[r,c]*b[r,c]   (float64, 2 stages)
// ------------------------------------------------------------------------------------------
constexpr int kMomLanes = 16;       // row lanes per block (x 32 float4 column lanes = 512 threads)
constexpr int kMomMaxChunks = 296;  // stage-1 blocks: two per SM, each walks a contiguous row range

inline int mom_chunks(int64_t n) {
  const int64_t c = (n + 127) / 128;
  return (int)(c < 1 ? 1 : (c > kMomMaxChunks ? kMomMaxChunks : c));
}
inline int64_t mom_rows(int64_t n, int chunks) {
  const int64_t r = (n + chunks - 1) / chunks;
  return (r + kMomLanes - 1) / kMomLanes * kMomLanes;
}

// rows of the statistics workspace: enough for our own stage 1 and for the GEMM-epilogue partials
inline int part_chunks(int64_t n) {
  const int a = mom_chunks(n), b = dcgc_gemm_stats_max_chunks();
  return a > b ? a : b;
}

// block: 32 column lanes (float4 -> 128 columns) x 16 row lanes walking rows r0+ry, r0+ry+16, ...;
// 4 independent 128-bit loads per tensor in flight per thread; fp32 partials over 4 rows, float64 from
// there on (fixed order => deterministic).
__global__ void __launch_bounds__(32 * kMomLanes, 2)
col_moments_partial(const float* __restrict__ a, int64_t ld_a, const float* __restrict__ b, int64_t ld_b,
                    int64_t n_rows, int width, int64_t rows_per_chunk, double* __restrict__ part) {
  dcgc_griddep_wait();
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + 4 * cx;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = min(n_rows, r0 + rows_per_chunk);
  double da[4] = {0, 0, 0, 0}, dab[4] = {0, 0, 0, 0};
  if (c < width) {
    for (int64_t rb = r0 + ry; rb < r1; rb += 4 * kMomLanes) {
      float4 va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = rb + (int64_t)u * kMomLanes;
        if (r < r1) {
          va[u] = __ldg(reinterpret_cast<const float4*>(a + r * ld_a + c));
          vb[u] = (b == a) ? va[u] : __ldg(reinterpret_cast<const float4*>(b + r * ld_b + c));
        } else {
          va[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          vb[u] = va[u];
        }
      }
      float sa[4] = {0, 0, 0, 0}, sab[4] = {0, 0, 0, 0};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sa[0] += va[u].x; sa[1] += va[u].y; sa[2] += va[u].z; sa[3] += va[u].w;
        sab[0] = fmaf(va[u].x, vb[u].x, sab[0]); sab[1] = fmaf(va[u].y, vb[u].y, sab[1]);
        sab[2] = fmaf(va[u].z, vb[u].z, sab[2]); sab[3] = fmaf(va[u].w, vb[u].w, sab[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { da[e] += sa[e]; dab[e] += sab[e]; }
    }
  }
  __shared__ double sh[2][kMomLanes][128];
#pragma unroll
  for (int e = 0; e < 4; ++e) { sh[0][ry][4 * cx + e] = da[e]; sh[1][ry][4 * cx + e] = dab[e]; }
  __syncthreads();
  // fixed-order reduction over the row lanes: threads 0..255 -> (quantity, column)
  if (threadIdx.x < 256) {
    const int q = threadIdx.x >> 7, col = threadIdx.x & 127;
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < kMomLanes; ++l) s += sh[q][l][col];
    const int cc = blockIdx.y * 128 + col;
    if (cc < width) part[((int64_t)blockIdx.x * 2 + q) * width + cc] = s;
  }
}

// Sum of the stage-1 partials.  One WARP per column (8 columns per 256-thread block, so 16 blocks at width 128 instead
// of the 8 blocks of 16 columns x 32 chunk lanes this replaced: the kernel is pure latency, 5.6 us in the launch list):
// lane l owns chunks l, l + 32, ... with all its loads issued before the first add (one memory latency), then the 32
// lane sums are combined by a fixed butterfly of double shuffles — deterministic, no shared memory, no block barrier.
constexpr int kFinCols = 8;        // columns (= warps) per block
constexpr int kPartPerLane = 10;   // covers 320 chunks; more are handled by the strided loop below
__device__ __forceinline__ bool reduce_partials(const double* __restrict__ part, int n_chunks, int width,
                                                double& s, double& ss, int& c_out) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * kFinCols + (threadIdx.x >> 5);
  if (c >= width) return false;              // (whole warps: the shuffles below stay converged)
  double va[kPartPerLane], vb[kPartPerLane];
#pragma unroll
  for (int u = 0; u < kPartPerLane; ++u) {
    const int k = lane + 32 * u;
    va[u] = k < n_chunks ? part[((int64_t)k * 2) * width + c] : 0.0;
    vb[u] = k < n_chunks ? part[((int64_t)k * 2 + 1) * width + c] : 0.0;
  }
  double a = 0.0, b = 0.0;
#pragma unroll
  for (int u = 0; u < kPartPerLane; ++u) { a += va[u]; b += vb[u]; }
  for (int k = lane + 32 * kPartPerLane; k < n_chunks; k += 32) {
    a += part[((int64_t)k * 2) * width + c];
    b += part[((int64_t)k * 2 + 1) * width + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane != 0) return false;
  s = a; ss = b; c_out = c;
  return true;
}

// BatchNorm forward finalize: batch statistics -> folded scale/shift, running-stat update.
// torch semantics (graphconvmodel.py:150-158: eps=1e-3, momentum=0.99 meaning new-stat weight):
//   normalise with the biased batch variance; running_var uses the unbiased one.
__global__ void __launch_bounds__(32 * kFinCols) bn_fwd_finalize(const double* __restrict__ part, int n_chunks, int width, int64_t n_rows,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                float* __restrict__ scale_out, float* __restrict__ shift_out) {
  dcgc_griddep_wait();
  double s, ss;
  int c;
  if (!reduce_partials(part, n_chunks, width, s, ss, c)) return;
  const double n = (double)n_rows;
  const double mean = n > 0 ? s / n : 0.0;
  double var = n > 0 ? ss / n - mean * mean : 0.0;
  if (var < 0) var = 0;
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = (float)(g * invstd);
  mean_out[c] = (float)mean;
  invstd_out[c] = (float)invstd;
  scale_out[c] = sc;
  shift_out[c] = (float)(b - mean * (double)sc);
  if (running_mean) {
    const double unbiased = n > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_fold(const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ running_mean, const float* __restrict__ running_var,
                             float eps, int width, float* __restrict__ scale_out, float* __restrict__ shift_out) {
  dcgc_griddep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= width) return;
  const float invstd = 1.f / sqrtf(running_var[c] + eps);
  const float sc = gamma[c] * invstd;
  scale_out[c] = sc;
  shift_out[c] = beta[c] - running_mean[c] * sc;
}

// BatchNorm backward finalize: dgamma, dbeta and the three per-channel coefficients of
//   dY = c1 * (dA - mean_dA - xhat * mean_dAx),  xhat = (Y - mean) * invstd
__global__ void __launch_bounds__(32 * kFinCols) bn_bwd_finalize(const double* __restrict__ part, int n_chunks, int width, int64_t n_rows,
                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ scale, float* __restrict__ dgamma,
                                float* __restrict__ dbeta, float* __restrict__ coef /* [3, width] */) {
  dcgc_griddep_wait();
  double sda, sday;
  int c;
  if (!reduce_partials(part, n_chunks, width, sda, sday, c)) return;
  const double n = (double)n_rows;
  const double dg = (double)invstd[c] * (sday - (double)mean[c] * sda);  // sum dA * xhat
  if (dgamma) dgamma[c] = (float)dg;
  if (dbeta) dbeta[c] = (float)sda;
  coef[c] = scale[c];
  coef[width + c] = n > 0 ? (float)(sda / n) : 0.f;
  coef[2 * width + c] = n > 0 ? (float)(dg / n) : 0.f;
}

// g = relu'(y) * bn_backward(dA)   (in place on dA allowed).  One thread owns one 4-column group and walks
// kApplyRows rows of a row block, so the five per-channel coefficient vectors are read once per thread
// instead of once per element (ncu r1j: LSU wavefronts 53 %, L1 hit rate 74 % from those re-reads).
constexpr int kApplyRows = 4;
__global__ void __launch_bounds__(kT)
bn_relu_bwd_apply(const float* da, int64_t ld_da, const float* __restrict__ y, int64_t ld_y,
                  const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ coef, int64_t n_rows, int width, int relu, float* g, int64_t ld_g) {
  dcgc_griddep_wait();
  const int groups = width >> 2;
  const int rows_per_block = (kT / groups) * kApplyRows;      // kT is a multiple of groups or groups > kT
  const int tg = threadIdx.x % groups, tr = threadIdx.x / groups;
  const int lanes = kT / groups;
  if (tr >= lanes) return;
  const int c = tg << 2;
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
  const float4 c1 = __ldg(reinterpret_cast<const float4*>(coef + c));
  const float4 c2 = __ldg(reinterpret_cast<const float4*>(coef + width + c));
  const float4 c3 = __ldg(reinterpret_cast<const float4*>(coef + 2 * width + c));
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block + tr;
  float4 d[kApplyRows], v[kApplyRows];
#pragma unroll
  for (int u = 0; u < kApplyRows; ++u) {
    const int64_t row = r0 + (int64_t)u * lanes;
    if (row < n_rows) {
      d[u] = *reinterpret_cast<const float4*>(da + row * ld_da + c);
      v[u] = __ldg(reinterpret_cast<const float4*>(y + row * ld_y + c));
    }
  }
#pragma unroll
  for (int u = 0; u < kApplyRows; ++u) {
    const int64_t row = r0 + (int64_t)u * lanes;
    if (row >= n_rows) continue;
    float4 o;
    o.x = c1.x * (d[u].x - c2.x - (v[u].x - m.x) * is.x * c3.x);
    o.y = c1.y * (d[u].y - c2.y - (v[u].y - m.y) * is.y * c3.y);
    o.z = c1.z * (d[u].z - c2.z - (v[u].z - m.z) * is.z * c3.z);
    o.w = c1.w * (d[u].w - c2.w - (v[u].w - m.w) * is.w * c3.w);
    if (relu) {
      o.x = v[u].x > 0.f ? o.x : 0.f; o.y = v[u].y > 0.f ? o.y : 0.f;
      o.z = v[u].z > 0.f ? o.z : 0.f; o.w = v[u].w > 0.f ? o.w : 0.f;
    }
    *reinterpret_cast<float4*>(g + row * ld_g + c) = o;
  }
}

// no BatchNorm: g = relu'(y) * dA
__global__ void __launch_bounds__(kT)
relu_bwd_apply(const float* da, int64_t ld_da, const float* __restrict__ y, int64_t ld_y, int64_t n_rows,
               int width, float* g, int64_t ld_g) {
  dcgc_griddep_wait();
  const int64_t t = (int64_t)blockIdx.x * kT + threadIdx.x;
  const int groups = width >> 2;
  const int64_t row = t / groups;
  const int c = (int)(t - row * groups) << 2;
  if (row >= n_rows) return;
  float4 d = *reinterpret_cast<const float4*>(da + row * ld_da + c);
  const float4 v = __ldg(reinterpret_cast<const float4*>(y + row * ld_y + c));
  d.x = v.x > 0.f ? d.x : 0.f; d.y = v.y > 0.f ? d.y : 0.f;
  d.z = v.z > 0.f ? d.z : 0.f; d.w = v.w > 0.f ? d.w : 0.f;
  *reinterpret_cast<float4*>(g + row * ld_g + c) = d;
}

// ------------------------------------------------------------------------------------------
// Synchronised BatchNorm: the finalize kernels with the cross-rank exchange inside (include/dcgc.h, dcgc_bn_sync).
// ONE block: thread c owns column c.  (1) local sums of the partial rows; (2) they are stored into the mailbox row
// [this rank][slot] of EVERY rank (peer stores over NVLink), system fence, then the sequence flag of that row;
// (3) wait until all `world` rows of the own mailbox carry the flag; (4) add the rows in rank order — every rank sees
// the same values in the same order: bit-identical statistics — and finalize with the global row count.
// KIND 1: forward (mean / invstd / scale / shift, running statistics); KIND 2: backward (dgamma / dbeta from the LOCAL
// sums: the gradient exchange adds them over ranks; the apply coefficients from the global ones).
// ------------------------------------------------------------------------------------------
struct SyncArgs {
  int world, rank, cap;
  unsigned long long seq;
  double* mailbox[DCGC_SYNC_MAX_RANKS];
};
template <int KIND>
__global__ void __launch_bounds__(512)
bn_sync_kernel(const SyncArgs sy, const double* __restrict__ part, int n_chunks, int width, long long n_rows_local,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
               float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean_io,
               float* __restrict__ invstd_io, float* __restrict__ scale_io, float* __restrict__ shift_out,
               float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  dcgc_griddep_wait_only();     // (no early launch of the kernel behind this one: it waits for its peers below)
  const int t = threadIdx.x;
  const int slot = (int)(sy.seq & 1ull);
  const size_t row = 2 * (size_t)sy.cap + 2;
  double a_loc[2] = {0.0, 0.0}, b_loc[2] = {0.0, 0.0};          // up to two columns per thread (width <= 1024)
  __shared__ double sh_a[512], sh_b[512];
  const bool sliced = width <= 256;        // every thread takes a slice of the partial rows of one column (16 loads deep)
  if (sliced) {
    const int slices = 512 / width, per = (n_chunks + slices - 1) / slices;
    const int c = t % width, sl = t / width;
    double a = 0.0, b = 0.0;
    if (sl < slices) {
      const int k1 = min(n_chunks, (sl + 1) * per);
      for (int k = sl * per; k < k1; k += 16) {
        double va[16], vb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          va[u] = k + u < k1 ? part[((size_t)(k + u) * 2) * width + c] : 0.0;
          vb[u] = k + u < k1 ? part[((size_t)(k + u) * 2 + 1) * width + c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) { a += va[u]; b += vb[u]; }
      }
    }
    sh_a[t] = a; sh_b[t] = b;
    __syncthreads();
    if (t < width)
      for (int q = 0; q < slices; ++q) { a_loc[0] += sh_a[q * width + t]; b_loc[0] += sh_b[q * width + t]; }
  }
  for (int i = 0, c = t; c < width; c += 512, ++i) {
    double a = a_loc[i], b = b_loc[i];
    if (!sliced) {
      int k = 0;
      for (; k + 8 <= n_chunks; k += 8) {
        double va[8], vb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          va[u] = part[((size_t)(k + u) * 2) * width + c];
          vb[u] = part[((size_t)(k + u) * 2 + 1) * width + c];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { a += va[u]; b += vb[u]; }
      }
      for (; k < n_chunks; ++k) { a += part[((size_t)k * 2) * width + c]; b += part[((size_t)k * 2 + 1) * width + c]; }
      a_loc[i] = a; b_loc[i] = b;
    }
    for (int p = 0; p < sy.world; ++p) {
      double* dst = sy.mailbox[p] + ((size_t)sy.rank * 2 + slot) * row;
      dst[c] = a;
      dst[sy.cap + c] = b;
    }
  }
  if (t == 0)
    for (int p = 0; p < sy.world; ++p) (sy.mailbox[p] + ((size_t)sy.rank * 2 + slot) * row)[2 * sy.cap] = (double)n_rows_local;
  __threadfence_system();
  __syncthreads();
  if (t < sy.world) {
    volatile unsigned long long* flag =
        reinterpret_cast<volatile unsigned long long*>(sy.mailbox[t] + ((size_t)sy.rank * 2 + slot) * row + 2 * sy.cap + 1);
    *flag = sy.seq;
  }
  double* mine = sy.mailbox[sy.rank];
  if (t < sy.world) {
    volatile unsigned long long* flag =
        reinterpret_cast<volatile unsigned long long*>(mine + ((size_t)t * 2 + slot) * row + 2 * sy.cap + 1);
    unsigned long long t0 = 0;
    for (unsigned spin = 0; *flag != sy.seq; ++spin) {
      if ((spin & 0x3ffu) == 0x3ffu) {         // wall-clock bound (60 s): a dead peer traps instead of hanging the GPU
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 60000000000ull) __trap();
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  double rows = 0.0;
  for (int r = 0; r < sy.world; ++r) rows += __ldcv(mine + ((size_t)r * 2 + slot) * row + 2 * sy.cap);
  for (int i = 0, c = t; c < width; c += 512, ++i) {
    double a = 0.0, b = 0.0;
    for (int r = 0; r < sy.world; ++r) {
      a += __ldcv(mine + ((size_t)r * 2 + slot) * row + c);
      b += __ldcv(mine + ((size_t)r * 2 + slot) * row + sy.cap + c);
    }
    if (KIND == 1) {
      const double mean = rows > 0 ? a / rows : 0.0;
      double var = rows > 0 ? b / rows - mean * mean : 0.0;
      if (var < 0) var = 0;
      const double invstd = 1.0 / sqrt(var + (double)eps);
      const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float sc = (float)(g * invstd);
      mean_io[c] = (float)mean;
      invstd_io[c] = (float)invstd;
      scale_io[c] = sc;
      shift_out[c] = (float)(be - mean * (double)sc);
      if (running_mean) {
        const double unbiased = rows > 1 ? var * rows / (rows - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
      }
    } else {
      const double is = (double)invstd_io[c], mu = (double)mean_io[c];
      const double dg_loc = is * (b_loc[i] - mu * a_loc[i]);
      const double dg = is * (b - mu * a);
      if (dgamma) dgamma[c] = (float)dg_loc;
      if (dbeta) dbeta[c] = (float)a_loc[i];
      coef[c] = scale_io[c];
      coef[width + c] = rows > 0 ? (float)(a / rows) : 0.f;
      coef[2 * width + c] = rows > 0 ? (float)(dg / rows) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// GraphConv bias packing: reference order b[0..20] (layers.py:6189-6226) <-> per-degree sums
// ------------------------------------------------------------------------------------------
struct BiasPackAll { const float* b21[DCGC_MODEL_MAX_LAYERS]; float* b11[DCGC_MODEL_MAX_LAYERS]; int c[DCGC_MODEL_MAX_LAYERS]; };
// every conv layer of the model in one launch (blockIdx.y = layer)
__global__ void conv_bias_pack_all(const BiasPackAll a) {
  dcgc_griddep_wait();
  const int l = blockIdx.y, c_out = a.c[l];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= DCGC_N_DEG * c_out) return;
  const int d = i / c_out, c = i - d * c_out;
  const float* b21 = a.b21[l];
  a.b11[l][i] = d == 0 ? b21[20 * c_out + c] : b21[(2 * (d - 1)) * c_out + c] + b21[(2 * (d - 1) + 1) * c_out + c];
}
// ------------------------------------------------------------------------------------------
// head: out[b,t] = fp[b,:] . Wh[t,:] + bh[t]   (one warp per output element)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT)
head_fwd(const float* __restrict__ fp, int64_t ld_fp, const float* __restrict__ wh, const float* __restrict__ bh,
         int64_t n_rows, int k, int n_out, float* __restrict__ out) {
  dcgc_griddep_wait();
  const int64_t warp = ((int64_t)blockIdx.x * kT + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_rows * n_out) return;
  const int64_t b = warp / n_out;
  const int t = (int)(warp - b * n_out);
  float s = 0.f;
  for (int j = lane; j < k; j += 32) s = fmaf(__ldg(fp + b * ld_fp + j), __ldg(wh + (int64_t)t * k + j), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[b * n_out + t] = s + (bh ? bh[t] : 0.f);
}

// per-(sample, task) loss and d(out); mean over all n_samples*n_tasks elements
// (torch_model.py:1275-1294).  mode 0: w*(out-y)^2 (losses.py:76-94); mode 1: -w*sum_c y_c
// log_softmax(logits)_c (losses.py:236-259).
__global__ void __launch_bounds__(kT)
loss_fwd_bwd(const float* __restrict__ out, const float* __restrict__ y, const float* __restrict__ w,
             int64_t n_elems, int n_classes, int mode, float* __restrict__ per_elem, float* __restrict__ dout) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n_elems) return;
  const float wi = w ? w[i] : 1.f;
  const float inv = 1.f / (float)n_elems;
  if (mode == 0) {
    const float d = out[i] - y[i];
    per_elem[i] = wi * d * d;
    dout[i] = 2.f * wi * d * inv;
  } else {
    const float* lg = out + i * n_classes;
    const float* yy = y + i * n_classes;
    float mx = -INFINITY;
    for (int c = 0; c < n_classes; ++c) mx = fmaxf(mx, lg[c]);
    float se = 0.f, sy = 0.f;
    for (int c = 0; c < n_classes; ++c) { se += expf(lg[c] - mx); sy += yy[c]; }
    const float lse = mx + logf(se);
    float l = 0.f;
    for (int c = 0; c < n_classes; ++c) {
      const float logp = lg[c] - lse;
      l -= yy[c] * logp;
      dout[i * n_classes + c] = wi * inv * (expf(logp) * sy - yy[c]);
    }
    per_elem[i] = wi * l;
  }
}

__global__ void __launch_bounds__(1024) loss_reduce(const float* __restrict__ per_elem, int64_t n_elems,
                                                    float* __restrict__ loss) {
  dcgc_griddep_wait();
  __shared__ double sh[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n_elems; i += 1024) s += per_elem[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = n_elems > 0 ? (float)(sh[0] / (double)n_elems) : 0.f;
}

__global__ void softmax_rows(const float* __restrict__ logits, int64_t n_rows, int n_classes,
                             float* __restrict__ probs) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const float* lg = logits + i * n_classes;
  float mx = -INFINITY;
  for (int c = 0; c < n_classes; ++c) mx = fmaxf(mx, lg[c]);
  float se = 0.f;
  for (int c = 0; c < n_classes; ++c) se += expf(lg[c] - mx);
  for (int c = 0; c < n_classes; ++c) probs[i * n_classes + c] = expf(lg[c] - mx) / se;
}

// head backward, stage 1: partial dWh[t,j] / dbh[t] over a chunk of samples
constexpr int kHeadChunk = 128;
__global__ void __launch_bounds__(kT)
head_bwd_partial(const float* __restrict__ dout, const float* __restrict__ fp, int64_t ld_fp, int64_t n_rows,
                 int k, int n_out, float* __restrict__ part /* [chunks, n_out, k+1] */) {
  dcgc_griddep_wait();
  const int t = blockIdx.y;
  const int64_t b0 = (int64_t)blockIdx.x * kHeadChunk, b1 = min(n_rows, b0 + kHeadChunk);
  for (int j = threadIdx.x; j <= k; j += kT) {
    float s = 0.f;
    if (j < k) {
      for (int64_t b = b0; b < b1; ++b) s = fmaf(__ldg(dout + b * n_out + t), __ldg(fp + b * ld_fp + j), s);
    } else {
      for (int64_t b = b0; b < b1; ++b) s += __ldg(dout + b * n_out + t);
    }
    part[((int64_t)blockIdx.x * n_out + t) * (k + 1) + j] = s;
  }
}
__global__ void head_bwd_final(const float* __restrict__ part, int n_chunks, int k, int n_out,
                               float* __restrict__ dwh, float* __restrict__ dbh) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_out * (k + 1)) return;
  const int t = (int)(i / (k + 1)), j = (int)(i - (int64_t)t * (k + 1));
  float s = 0.f;
  for (int c = 0; c < n_chunks; ++c) s += part[((int64_t)c * n_out + t) * (k + 1) + j];
  if (j < k) dwh[(int64_t)t * k + j] = s;
  else if (dbh) dbh[t] = s;
}
// dfp[b,j] = sum_t dout[b,t] * Wh[t,j] for b < n_rows, 0 for the padded segments
__global__ void __launch_bounds__(kT)
head_bwd_input(const float* __restrict__ dout, const float* __restrict__ wh, int64_t n_rows, int64_t n_seg, int k,
               int n_out, float* __restrict__ dfp) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n_seg * k) return;
  const int64_t b = i / k;
  const int j = (int)(i - b * k);
  float s = 0.f;
  if (b < n_rows)
    for (int t = 0; t < n_out; ++t) s = fmaf(__ldg(dout + b * n_out + t), __ldg(wh + (int64_t)t * k + j), s);
  dfp[i] = s;
}

// ------------------------------------------------------------------------------------------
// Fused head for small output widths (n_out <= 32): output layer, loss, d(out), d(fingerprint) and the per-block
// partials of dWh / dbh / loss in ONE kernel over blocks of 32 samples, then one small kernel that adds the partials
// in block order.  Replaces head_fwd + loss_fwd_bwd + loss_reduce + head_bwd_partial + head_bwd_final +
// head_bwd_input (six dependent launches, 38 us of a 1.2 ms step in the ncu launch list, none of them with more than
// a few hundred microseconds of bandwidth-bound work).  The block keeps its 32 fingerprint rows and the whole Wh in
// shared memory; rows are padded by one float so that threads walking different samples hit different banks.
// ------------------------------------------------------------------------------------------
constexpr int kHfRows = 32;
__global__ void __launch_bounds__(kT)
head_fused_kernel(const float* __restrict__ fp, int64_t ld_fp, const float* __restrict__ wh, const float* __restrict__ bh,
                  const float* __restrict__ y, const float* __restrict__ w, int64_t n_samples, int64_t n_seg, int k,
                  int n_out, int n_classes, int mode, int64_t n_elems, float* __restrict__ out, float* __restrict__ dfp,
                  float* __restrict__ part /* [blocks][n_out][k + 1] */, double* __restrict__ loss_part /* [blocks] */) {
  dcgc_griddep_wait();
  extern __shared__ float hsm[];
  const int kp = k + 1;
  float* fp_s = hsm;                          // [32][k + 1]
  float* wh_s = fp_s + kHfRows * kp;          // [n_out][k]
  float* out_s = wh_s + n_out * k;            // [32][n_out]
  float* dout_s = out_s + kHfRows * n_out;    // [32][n_out]
  float* le_s = dout_s + kHfRows * n_out;     // [32][per_row] per-element losses
  const int per_row = mode == 1 ? n_out / n_classes : n_out;
  const int64_t b0 = (int64_t)blockIdx.x * kHfRows;
  const int tid = threadIdx.x;
  // staging: eight independent loads in flight per thread (a load -> store loop left to the compiler ran one L2 round
  // trip per iteration: 17 us for this kernel in the first version)
  for (int i0 = tid; i0 < n_out * k; i0 += 8 * kT) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = i0 + u * kT < n_out * k ? __ldg(wh + i0 + u * kT) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u * kT < n_out * k) wh_s[i0 + u * kT] = v[u];
  }
  for (int j = tid; j < k; j += kT) {
    for (int base = 0; base < kHfRows; base += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = b0 + base + u < n_samples ? __ldg(fp + (b0 + base + u) * ld_fp + j) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) fp_s[(base + u) * kp + j] = v[u];
    }
  }
  __syncthreads();
  // ---- output layer.  Few outputs: one WARP per (sample, output), lanes over the fingerprint; many: one thread each
  if (kHfRows * n_out <= kT) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int task = warp; task < kHfRows * n_out; task += kT / 32) {
      const int bl = task & (kHfRows - 1), t = task >> 5;
      float acc = 0.f;
      for (int j = lane; j < k; j += 32) acc = fmaf(fp_s[bl * kp + j], wh_s[t * k + j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        acc += bh ? __ldg(bh + t) : 0.f;
        out_s[bl * n_out + t] = acc;
        dout_s[bl * n_out + t] = 0.f;
        if (b0 + bl < n_samples) out[(b0 + bl) * n_out + t] = acc;
      }
    }
  } else {
    for (int task = tid; task < kHfRows * n_out; task += kT) {
      const int bl = task & (kHfRows - 1), t = task >> 5;
      const float* fr = fp_s + bl * kp;
      const float* wr = wh_s + t * k;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;       // four independent chains
      int j = 0;
      for (; j + 4 <= k; j += 4) {
        a0 = fmaf(fr[j], wr[j], a0); a1 = fmaf(fr[j + 1], wr[j + 1], a1);
        a2 = fmaf(fr[j + 2], wr[j + 2], a2); a3 = fmaf(fr[j + 3], wr[j + 3], a3);
      }
      for (; j < k; ++j) a0 = fmaf(fr[j], wr[j], a0);
      float acc = (a0 + a1) + (a2 + a3);
      acc += bh ? __ldg(bh + t) : 0.f;
      out_s[bl * n_out + t] = acc;
      dout_s[bl * n_out + t] = 0.f;
      if (b0 + bl < n_samples) out[(b0 + bl) * n_out + t] = acc;
    }
  }
  __syncthreads();
  // ---- loss and d(out) per (sample, loss element); torch_model.py:1275-1294, losses.py:76-94 / 236-259
  const float inv = 1.f / (float)n_elems;
  for (int task = tid; task < kHfRows * per_row; task += kT) {
    const int bl = task / per_row, te = task - bl * per_row;
    float l = 0.f;
    if (b0 + bl < n_samples) {
      const int64_t i = (b0 + bl) * per_row + te;
      const float wi = w ? __ldg(w + i) : 1.f;
      if (mode == 0) {
        const float d = out_s[bl * n_out + te] - __ldg(y + i);
        l = wi * d * d;
        dout_s[bl * n_out + te] = 2.f * wi * d * inv;
      } else {
        const float* lg = out_s + bl * n_out + te * n_classes;
        const float* yy = y + i * n_classes;
        float mx = -INFINITY;
        for (int c = 0; c < n_classes; ++c) mx = fmaxf(mx, lg[c]);
        float se = 0.f, sy = 0.f;
        for (int c = 0; c < n_classes; ++c) { se += expf(lg[c] - mx); sy += __ldg(yy + c); }
        const float lse = mx + logf(se);
        for (int c = 0; c < n_classes; ++c) {
          const float logp = lg[c] - lse, yc = __ldg(yy + c);
          l -= yc * logp;
          dout_s[bl * n_out + te * n_classes + c] = wi * inv * (expf(logp) * sy - yc);
        }
        l *= wi;
      }
    }
    le_s[task] = l;
  }
  __syncthreads();
  // ---- d(fingerprint) rows (zero for the padded segments) and this block's partial dWh / dbh
  for (int j = tid; j < k; j += kT) {
    for (int bl = 0; bl < kHfRows; ++bl) {
      if (b0 + bl >= n_seg) break;
      float acc = 0.f;
      for (int t = 0; t < n_out; ++t) acc = fmaf(dout_s[bl * n_out + t], wh_s[t * k + j], acc);
      dfp[(b0 + bl) * k + j] = acc;
    }
    for (int t = 0; t < n_out; ++t) {
      float acc = 0.f;
      for (int bl = 0; bl < kHfRows; ++bl) acc = fmaf(dout_s[bl * n_out + t], fp_s[bl * kp + j], acc);
      part[((int64_t)blockIdx.x * n_out + t) * kp + j] = acc;
    }
  }
  if (tid < n_out) {
    float acc = 0.f;
    for (int bl = 0; bl < kHfRows; ++bl) acc += dout_s[bl * n_out + tid];
    part[((int64_t)blockIdx.x * n_out + tid) * kp + k] = acc;
  }
  if (tid == 0) {
    double acc = 0.0;
    for (int i = 0; i < kHfRows * per_row; ++i) acc += (double)le_s[i];
    loss_part[blockIdx.x] = acc;
  }
}
// partials -> dWh, dbh and the mean loss.  One WARP per output element: lane l adds the partials of blocks l, l + 32,
// ... in block order (independent loads: one memory latency, where a single thread per element walked all blocks one
// dependent round trip after the other) and the 32 lane sums are combined by a fixed shuffle tree: deterministic.
__global__ void __launch_bounds__(kT)
head_fused_final(const float* __restrict__ part, const double* __restrict__ loss_part, int n_blocks, int k, int n_out,
                 int64_t n_elems, float* __restrict__ dwh, float* __restrict__ dbh, float* __restrict__ loss) {
  dcgc_griddep_wait();
  const int64_t i = ((int64_t)blockIdx.x * kT + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_el = (int64_t)n_out * (k + 1);
  if (i < n_el) {
    float acc = 0.f;
    for (int c = lane; c < n_blocks; c += 32) acc += __ldg(part + (int64_t)c * n_el + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const int t = (int)(i / (k + 1)), j = (int)(i - (int64_t)t * (k + 1));
      if (j < k) dwh[(int64_t)t * k + j] = acc;
      else if (dbh) dbh[t] = acc;
    }
  } else if (i == n_el) {
    double acc = 0.0;
    for (int c = lane; c < n_blocks; c += 32) acc += loss_part[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) *loss = n_elems > 0 ? (float)(acc / (double)n_elems) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// Adam on the flat slab (torch.optim.Adam semantics, models/optimizers.py:190-241)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// four parameters per thread (the slab and the optimizer state are 16-byte aligned allocations)
__global__ void __launch_bounds__(kT)
adam_kernel_vec(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                int64_t n4, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n4) return;
  const float4 gv = g[i];
  float4 mv = m[i], vv = v[i], pv = p[i];
  const float step = lr / bc1;
#define DCGC_ADAM1(X)                                           \
  {                                                             \
    const float gi = gv.X * grad_scale;                         \
    mv.X = b1 * mv.X + (1.f - b1) * gi;                         \
    vv.X = b2 * vv.X + (1.f - b2) * gi * gi;                    \
    pv.X -= step * (mv.X / (sqrtf(vv.X) / bc2_sqrt + eps));     \
  }
  DCGC_ADAM1(x) DCGC_ADAM1(y) DCGC_ADAM1(z) DCGC_ADAM1(w)
#undef DCGC_ADAM1
  m[i] = mv; v[i] = vv; p[i] = pv;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Arena {
  char* base;
  int64_t off, cap;
  bool ok = true;
  template <typename T>
  T* take(int64_t n) {
    const int64_t bytes = dcgc_align_up(n * (int64_t)sizeof(T), 256);
    if (off + bytes > cap) { ok = false; off += bytes; return nullptr; }
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

struct Layout {
  int L;
  int f[DCGC_MODEL_MAX_LAYERS + 1];   // input width of conv l (f[0] = n_feat), f[L] = last conv width
  int fp[DCGC_MODEL_MAX_LAYERS + 1];  // padded to a multiple of 4
  int64_t conv_w[DCGC_MODEL_MAX_LAYERS], conv_b[DCGC_MODEL_MAX_LAYERS];
  int64_t bn_g[DCGC_MODEL_MAX_LAYERS + 1], bn_b[DCGC_MODEL_MAX_LAYERS + 1];
  int64_t dense_w, dense_b, head_w, head_b, n_params;
  int64_t bn_mean[DCGC_MODEL_MAX_LAYERS + 1], bn_var[DCGC_MODEL_MAX_LAYERS + 1], n_bn;
};

int make_layout(const dcgc_gcmodel_config* cfg, Layout* lo) {
  DCGC_CHECK_ARG(cfg, "dcgc_gcmodel: null config");
  DCGC_CHECK_ARG(cfg->n_layers >= 1 && cfg->n_layers <= DCGC_MODEL_MAX_LAYERS, "dcgc_gcmodel: 1..%d conv layers",
                 DCGC_MODEL_MAX_LAYERS);
  DCGC_CHECK_ARG(cfg->n_feat > 0 && cfg->dense > 0 && cfg->n_out > 0, "dcgc_gcmodel: bad widths");
  DCGC_CHECK_ARG(cfg->dense % 4 == 0, "dcgc_gcmodel: dense width must be a multiple of 4");
  DCGC_CHECK_ARG(cfg->mode == 0 || (cfg->mode == 1 && cfg->n_classes >= 2 && cfg->n_out % cfg->n_classes == 0),
                 "dcgc_gcmodel: bad mode / n_classes");
  const int L = cfg->n_layers;
  lo->L = L;
  lo->f[0] = cfg->n_feat;
  for (int l = 0; l < L; ++l) {
    DCGC_CHECK_ARG(cfg->widths[l] > 0 && cfg->widths[l] % 4 == 0,
                   "dcgc_gcmodel: conv widths must be positive multiples of 4");
    lo->f[l + 1] = cfg->widths[l];
  }
  for (int l = 0; l <= L; ++l) lo->fp[l] = (lo->f[l] + 3) / 4 * 4;
  int64_t off = 0, boff = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += (n + 3) / 4 * 4; return o; };  // 16-byte aligned tensors
  for (int l = 0; l < L; ++l) {
    const int c = cfg->widths[l];
    lo->conv_w[l] = take((int64_t)DCGC_N_DEG * 2 * lo->fp[l] * c);
    lo->conv_b[l] = take((int64_t)21 * c);
    if (cfg->batch_norm) {
      lo->bn_g[l] = take(c); lo->bn_b[l] = take(c);
      lo->bn_mean[l] = boff; boff += c; lo->bn_var[l] = boff; boff += c;
    } else {
      lo->bn_g[l] = lo->bn_b[l] = lo->bn_mean[l] = lo->bn_var[l] = -1;
    }
  }
  lo->dense_w = take((int64_t)cfg->dense * lo->f[L]);
  lo->dense_b = take(cfg->dense);
  if (cfg->batch_norm) {
    lo->bn_g[L] = take(cfg->dense); lo->bn_b[L] = take(cfg->dense);
    lo->bn_mean[L] = boff; boff += cfg->dense; lo->bn_var[L] = boff; boff += cfg->dense;
  } else {
    lo->bn_g[L] = lo->bn_b[L] = lo->bn_mean[L] = lo->bn_var[L] = -1;
  }
  lo->head_w = take((int64_t)cfg->n_out * 2 * cfg->dense);
  lo->head_b = take(cfg->n_out);
  lo->n_params = off;
  lo->n_bn = boff;
  return DCGC_OK;
}

struct Saved {
  const float* h[DCGC_MODEL_MAX_LAYERS + 1];  // conv inputs (h[0] = x), h[L] = last pool output
  int64_t ld_h[DCGC_MODEL_MAX_LAYERS + 1];
  float* s[DCGC_MODEL_MAX_LAYERS];
  float* y[DCGC_MODEL_MAX_LAYERS];
  uint8_t* arg[DCGC_MODEL_MAX_LAYERS];
  float* z;        // dense pre-BN (post-ReLU)
  float* stats;    // per BN: mean, invstd, scale, shift (4 * width each)
  int64_t stats_off[DCGC_MODEL_MAX_LAYERS + 1];
  float* fp;       // fingerprint [n_seg, 2D]
  float* zc_sum;   // training + BatchNorm: per molecule, sum over its rows of (z - batch mean) and the same for the
  float* zc_arg;   // arg-max row, [n_seg, D] each (dcgc_gather_fwd_train) — for the molecule-level BatchNorm-backward sums
  int32_t* argrow;
  float* out;      // [n_samples, n_out]
  float* b11[DCGC_MODEL_MAX_LAYERS];
  double* part;    // column-moment partials
  int n_chunks;
  // weight images built ahead on the side stream (train step, tensor-core modes; null = built in front of the GEMM)
  const float* img_fwd[DCGC_MODEL_MAX_LAYERS];
  const float* img_dense;
  cudaEvent_t img_fwd_ready;   // main stream waits on it before the first GEMM
  unsigned int* fin_counters;  // train step: zeroed counters of the last-CTA BatchNorm finalizes (2 per BatchNorm), or null
  const dcgc_bn_sync* sync;    // synchronised BatchNorm (world > 1), or null
};
static SyncArgs sync_args(const dcgc_bn_sync* sy, unsigned long long seq) {
  SyncArgs a{};
  a.world = sy->world; a.rank = sy->rank; a.cap = sy->cap; a.seq = seq;
  for (int r = 0; r < sy->world; ++r) a.mailbox[r] = static_cast<double*>(sy->mailbox[r]);
  return a;
}

// ---- early weight images: one side stream + three events per device, created on first use
struct SideStream { cudaStream_t s; cudaEvent_t e0, e1, e2; };
std::mutex g_side_mu;
std::map<int, SideStream> g_side;
int side_stream(SideStream** out) {
  int dev = 0;
  DCGC_CUDA_CALL(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_side_mu);
  auto it = g_side.find(dev);
  if (it == g_side.end()) {
    SideStream ss{};
    DCGC_CUDA_CALL(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
    DCGC_CUDA_CALL(cudaEventCreateWithFlags(&ss.e0, cudaEventDisableTiming));
    DCGC_CUDA_CALL(cudaEventCreateWithFlags(&ss.e1, cudaEventDisableTiming));
    DCGC_CUDA_CALL(cudaEventCreateWithFlags(&ss.e2, cudaEventDisableTiming));
    it = g_side.emplace(dev, ss).first;
  }
  *out = &it->second;
  return DCGC_OK;
}
// Forward GEMMs of the TF32x3 mode in fp16x3 (tc_gemm_kernel_v6: half the tensor-core instructions, the same 22-bit
// products; activations are in fp16's range by construction, gradients are not, so the backward GEMMs stay TF32x3).
// DCGC_FWD_F16X3=0 keeps the forward pass on tf32 halves.
bool fwd_f16x3_on(const dcgc_gcmodel_config* cfg) {
  static const bool on = [] { const char* e = getenv("DCGC_FWD_F16X3"); return !(e && e[0] == '0'); }();
  return on && cfg->gemm_mode == DCGC_GEMM_TF32X3;
}
bool early_images_on() {
  // measured (profiles/r3i_*): forward GEMM scopes 239 -> 210 us, dgrad 141 -> 131 us, step 1.2445 -> 1.2333 ms
  static const bool on = [] { const char* e = getenv("DCGC_EARLY_IMAGES"); return !(e && e[0] == '0'); }();
  return on;
}

int check_topo(const dcgc_topology* t) {
  DCGC_CHECK_ARG(t, "dcgc_gcmodel: null topology");
  DCGC_CHECK_ARG(t->n_atoms >= 0 && t->n_edges >= 0 && t->n_segments >= 0 && t->n_tiles >= 0,
                 "dcgc_gcmodel: negative topology size");
  if (t->n_atoms > 0)
    DCGC_CHECK_ARG(t->row_ptr && t->t_row_ptr && t->mol_ptr && t->mol_atoms && t->membership && t->tiles,
                   "dcgc_gcmodel: null topology array");
  if (t->n_edges > 0) DCGC_CHECK_ARG(t->col_idx && t->t_src && t->t_slot, "dcgc_gcmodel: null adjacency array");
  return DCGC_OK;
}

// staged (molecule-group) kernels unless DCGC_NO_STAGED=1 (A/B measurements)
static bool staged_enabled() {
  static const bool on = [] { const char* e = getenv("DCGC_NO_STAGED"); return !(e && e[0] == '1'); }();
  return on;
}
static bool use_mg(const dcgc_topology* t, int64_t ld, int64_t ld_arg, const void* p0, const void* p1) {
  return staged_enabled() && t->symmetric && dcgc_mg_supported(t, ld, ld_arg) &&
         ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15) == 0;
}

// the fused head (head_fused_kernel) serves small output widths; shared memory: 32 fingerprint rows + Wh + out/dout/loss
static bool head_fused_ok(const dcgc_gcmodel_config* cfg) {
  static const bool off = [] { const char* e = getenv("DCGC_NO_FUSED_HEAD"); return e && e[0] == '1'; }();
  return !off && cfg->n_out <= 32 && 2 * cfg->dense <= 512;
}
static size_t head_fused_smem(const dcgc_gcmodel_config* cfg) {
  const int k = 2 * cfg->dense, T = cfg->n_out;
  return (size_t)(kHfRows * (k + 1) + T * k + 3 * kHfRows * T) * 4;
}

#define RET_IF(expr)                \
  do {                              \
    int st__ = (expr);              \
    if (st__ != DCGC_OK) return st__; \
  } while (0)

// Measured (profiles/r5p_last_cta_fin.md): at the bench shape (102 k atoms, 148 rows of partials = 300 KB pulled through
// one SM, plus a threadfence in every CTA's tail) the last-CTA finalize costs 7 us per step MORE than the eight small
// finalize kernels it replaces; for small batches (a few CTAs, launch-latency-bound steps) it removes eight of ~50
// launches.  So: on below kLastCtaFinAtoms atoms, off above; DCGC_LAST_CTA_FIN=0 / 1 forces it.
constexpr int64_t kLastCtaFinAtoms = 16384;
static bool last_cta_fin_on(int64_t n_atoms) {
  static const int forced = [] { const char* e = getenv("DCGC_LAST_CTA_FIN"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
  return forced >= 0 ? forced == 1 : n_atoms <= kLastCtaFinAtoms;
}
// forward finalize of BatchNorm idx by the last CTA of the kernel that writes the statistics partials
static DcgcBnFin fwd_fin(const dcgc_gcmodel_config* cfg, const Layout& lo, int idx, int width, int64_t n,
                         const float* params, float* bn_running, const Saved& sv) {
  DcgcBnFin f{};
  if (!sv.fin_counters || width > 256) return f;     // (one thread per column in the last CTA)
  float* stats = sv.stats + sv.stats_off[idx];
  f.counter = sv.fin_counters + 2 * idx; f.kind = 1; f.width = width; f.n_rows = n; f.part = sv.part;
  f.gamma = params + lo.bn_g[idx]; f.beta = params + lo.bn_b[idx]; f.eps = cfg->bn_eps; f.momentum = cfg->bn_momentum;
  f.running_mean = bn_running ? bn_running + lo.bn_mean[idx] : nullptr;
  f.running_var = bn_running ? bn_running + lo.bn_var[idx] : nullptr;
  f.mean_out = stats; f.invstd_out = stats + width; f.scale_out = stats + 2 * width; f.shift_out = stats + 3 * width;
  return f;
}

// fused_chunks >= 0: the stage-1 partials were already written by the GEMM epilogue (dcgc_*_fwd_stats);
// finalized: ... and the last CTA of that kernel has already written mean / invstd / scale / shift
int bn_forward(const dcgc_gcmodel_config* cfg, const Layout& lo, int idx, const float* y, int64_t ld_y, int64_t n,
               int width, const float* params, float* bn_running, int training, Saved& sv, cudaStream_t st,
               int fused_chunks = -1, bool finalized = false) {
  if (finalized && training && fused_chunks >= 0) return DCGC_OK;
  float* stats = sv.stats + sv.stats_off[idx];
  float *mean = stats, *invstd = stats + width, *scale = stats + 2 * width, *shift = stats + 3 * width;
  const float* gamma = params + lo.bn_g[idx];
  const float* beta = params + lo.bn_b[idx];
  float* rm = bn_running ? bn_running + lo.bn_mean[idx] : nullptr;
  float* rv = bn_running ? bn_running + lo.bn_var[idx] : nullptr;
  DcgcProfScope prof_scope("bn_stats_fwd", st);
  if (training) {
    int chunks = fused_chunks;
    if (chunks < 0) {
      chunks = n > 0 ? sv.n_chunks : 0;
      if (n > 0) {
        dim3 grid((unsigned)sv.n_chunks, (unsigned)((width + 127) / 128));
        dcgc_launch(col_moments_partial, grid, 32 * kMomLanes, 0, st, y, ld_y, y, ld_y, n, width, mom_rows(n, sv.n_chunks),
                                                             sv.part);
        DCGC_CUDA_LAUNCH_CHECK("col_moments_partial");
      }
    }
    if (sv.sync) {
      // statistics over the rows of every rank: exchange through the peer mailboxes inside the finalize kernel
      dcgc_launch(bn_sync_kernel<1>, 1, 512, 0, st, sync_args(sv.sync, sv.sync->seq0 + (unsigned long long)idx), sv.part, chunks,
                                           width, (long long)n, gamma, beta, cfg->bn_eps, cfg->bn_momentum, rm, rv, mean,
                                           invstd, scale, shift, nullptr, nullptr, nullptr);
      DCGC_CUDA_LAUNCH_CHECK("bn_sync_kernel (forward)");
      return DCGC_OK;
    }
    dcgc_launch(bn_fwd_finalize, (width + kFinCols - 1) / kFinCols, 32 * kFinCols, 0, st, sv.part, chunks, width, n, gamma, beta, cfg->bn_eps,
                                                       cfg->bn_momentum, rm, rv, mean, invstd, scale, shift);
    DCGC_CUDA_LAUNCH_CHECK("bn_fwd_finalize");
  } else {
    DCGC_CHECK_ARG(rm && rv, "dcgc_gcmodel: eval-mode BatchNorm needs the running statistics");
    dcgc_launch(bn_eval_fold, (width + 127) / 128, 128, 0, st, gamma, beta, rm, rv, cfg->bn_eps, width, scale, shift);
    DCGC_CUDA_LAUNCH_CHECK("bn_eval_fold");
  }
  return DCGC_OK;
}

int forward_impl(const dcgc_gcmodel_config* cfg, const Layout& lo, const dcgc_topology* t, const float* x,
                 int64_t ld_x, int64_t n_samples, const float* params, float* bn_running, int training,
                 int keep_arg, Arena& ws, Saved& sv, cudaStream_t st, bool skip_head = false) {
  const int L = lo.L, D = cfg->dense;
  const int64_t N = t->n_atoms, S = t->n_segments;
  DCGC_CHECK_ARG(ld_x >= lo.fp[0] || N == 0, "dcgc_gcmodel: x must be zero-padded to a leading dimension >= %d",
                 lo.fp[0]);
  DCGC_CHECK_ARG(n_samples >= 0 && n_samples <= S, "dcgc_gcmodel: n_samples outside [0, n_segments]");
  // ---- carve the workspace
  sv.n_chunks = mom_chunks(N);
  int wmax = D;
  for (int l = 0; l < L; ++l) wmax = wmax > cfg->widths[l] ? wmax : cfg->widths[l];
  sv.part = ws.take<double>((int64_t)part_chunks(N) * 2 * wmax);
  int64_t so = 0;
  for (int l = 0; l <= L; ++l) { sv.stats_off[l] = so; so += 4 * (int64_t)(l < L ? cfg->widths[l] : D); }
  sv.stats = ws.take<float>(so);
  sv.h[0] = x; sv.ld_h[0] = ld_x;
  for (int l = 0; l < L; ++l) {
    const int c = cfg->widths[l];
    sv.s[l] = ws.take<float>(N * lo.fp[l]);
    sv.y[l] = ws.take<float>(N * c);
    sv.arg[l] = keep_arg ? ws.take<uint8_t>(N * c) : nullptr;
    sv.h[l + 1] = ws.take<float>(N * c); sv.ld_h[l + 1] = c;
    sv.b11[l] = ws.take<float>(DCGC_N_DEG * c);
  }
  sv.z = ws.take<float>(N * D);
  sv.fp = ws.take<float>(S * 2 * D);
  sv.argrow = keep_arg ? ws.take<int32_t>(S * D) : nullptr;
  const bool want_zc = keep_arg && training && cfg->batch_norm && S > 0 && N > 0 && (S + 127) / 128 <= part_chunks(N);
  sv.zc_sum = want_zc ? ws.take<float>(S * D) : nullptr;
  sv.zc_arg = want_zc ? ws.take<float>(S * D) : nullptr;
  sv.out = ws.take<float>((n_samples > 0 ? n_samples : 1) * cfg->n_out);
  if (!ws.ok) {
    dcgc_set_error("dcgc_gcmodel: workspace too small (%lld bytes needed so far, %lld given)", (long long)ws.off,
                   (long long)ws.cap);
    return DCGC_ERR_NOMEM;
  }
  // ---- conv stack
  {
    BiasPackAll bp{};
    int cmax = 0;
    for (int l = 0; l < L; ++l) {
      bp.b21[l] = params + lo.conv_b[l]; bp.b11[l] = sv.b11[l]; bp.c[l] = cfg->widths[l];
      cmax = cmax > cfg->widths[l] ? cmax : cfg->widths[l];
    }
    dim3 grid(blocks_for(DCGC_N_DEG * cmax), (unsigned)L);
    dcgc_launch(conv_bias_pack_all, grid, kT, 0, st, bp);
    DCGC_CUDA_LAUNCH_CHECK("conv_bias_pack_all");
  }
  for (int l = 0; l < L; ++l) {
    const int c = cfg->widths[l], fp = lo.fp[l];
    const float* h = sv.h[l];
    const int64_t ld = sv.ld_h[l];
    if (use_mg(t, ld, 0, h, sv.s[l]))
      RET_IF(dcgc_mg_gather_sum(h, ld, t, 0, fp, nullptr, 0, sv.s[l], fp, st));
    else
      RET_IF(dcgc_gather_sum_bucketed(h, ld, t->deg_count, t->col_idx, N, fp, nullptr, 0, sv.s[l], fp, st));
    const bool fuse_stats = cfg->batch_norm && training && dcgc_tc_terms(cfg->gemm_mode) != 0;
    int32_t fused = -1;
    if (l == 0 && sv.img_fwd_ready) DCGC_CUDA_CALL(cudaStreamWaitEvent(st, sv.img_fwd_ready, 0));
    DcgcGemmOpts go;
    go.img = sv.img_fwd[l];
    go.a_exact = (l == 0 && cfg->input_exact) ? 1 : 0;      // integer-valued features: [X | S] is exact in tf32 / fp16
    go.f16x3 = fwd_f16x3_on(cfg) ? 1 : 0;
    const DcgcBnFin fin = fuse_stats ? fwd_fin(cfg, lo, l, c, N, params, bn_running, sv) : DcgcBnFin{};
    if (fin.kind) go.fin = &fin;
    RET_IF(dcgc_group_gemm_fwd_opts(cfg->gemm_mode, h, ld, fp, sv.s[l], fp, fp, params + lo.conv_w[l], sv.b11[l], c,
                                    t->tiles, t->n_tiles, 128, N, DCGC_ACT_RELU, sv.y[l], c,
                                    fuse_stats ? sv.part : nullptr, fuse_stats ? &fused : nullptr, go, st));
    const float *scale = nullptr, *shift = nullptr;
    if (cfg->batch_norm) {
      RET_IF(bn_forward(cfg, lo, l, sv.y[l], c, N, c, params, bn_running, training, sv, st, fused, fin.kind != 0));
      scale = sv.stats + sv.stats_off[l] + 2 * c;
      shift = scale + c;
    }
    if (use_mg(t, c, 0, sv.y[l], sv.h[l + 1]))
      RET_IF(dcgc_mg_pool_fwd(sv.y[l], c, scale, shift, t, c, const_cast<float*>(sv.h[l + 1]), c, sv.arg[l], c, st));
    else
      RET_IF(dcgc_pool_fwd(sv.y[l], c, scale, shift, t->row_ptr, t->col_idx, N, c, const_cast<float*>(sv.h[l + 1]), c,
                           sv.arg[l], c, st));
  }
  // ---- atom-level dense + ReLU (+BN folded into the gather), GraphGather(tanh), head
  int32_t fused_d = -1;
  DcgcBnFin fin_d{};
  {
    const bool fuse_d = cfg->batch_norm && training && dcgc_tc_terms(cfg->gemm_mode) != 0;
    DcgcGemmOpts go;
    go.img = sv.img_dense;
    go.f16x3 = fwd_f16x3_on(cfg) ? 1 : 0;
    fin_d = fuse_d ? fwd_fin(cfg, lo, L, D, N, params, bn_running, sv) : DcgcBnFin{};
    if (fin_d.kind) go.fin = &fin_d;
    RET_IF(dcgc_linear_fwd_opts(cfg->gemm_mode, sv.h[L], sv.ld_h[L], lo.f[L], params + lo.dense_w, params + lo.dense_b,
                                D, N, DCGC_ACT_RELU, sv.z, D, fuse_d ? sv.part : nullptr, fuse_d ? &fused_d : nullptr,
                                go, st));
  }
  const float *scale = nullptr, *shift = nullptr;
  if (cfg->batch_norm) {
    RET_IF(bn_forward(cfg, lo, L, sv.z, D, N, D, params, bn_running, training, sv, st, fused_d, fin_d.kind != 0));
    scale = sv.stats + sv.stats_off[L] + 2 * D;
    shift = scale + D;
  }
  RET_IF(dcgc_gather_fwd_train(sv.z, D, scale, shift, t->mol_ptr, t->mol_atoms, S, D, DCGC_ACT_TANH, sv.fp, 2 * D,
                               sv.argrow, sv.zc_sum ? sv.stats + sv.stats_off[L] : nullptr, sv.zc_sum, sv.zc_arg, st));
  if (n_samples > 0 && !skip_head) {
    dcgc_launch(head_fwd, blocks_for(n_samples * cfg->n_out * 32), kT, 0, st, sv.fp, 2 * D, params + lo.head_w,
                                                                    params + lo.head_b, n_samples, 2 * D, cfg->n_out,
                                                                    sv.out);
    DCGC_CUDA_LAUNCH_CHECK("head_fwd");
  }
  return DCGC_OK;
}

}  // namespace

extern "C" int dcgc_gcmodel_layout(const dcgc_gcmodel_config* cfg, int64_t* param_offsets, int64_t* bn_offsets,
                                   int64_t* n_params, int64_t* n_bn) {
  Layout lo;
  RET_IF(make_layout(cfg, &lo));
  if (param_offsets) {
    int k = 0;
    for (int l = 0; l < lo.L; ++l) {
      param_offsets[k++] = lo.conv_w[l]; param_offsets[k++] = lo.conv_b[l];
      param_offsets[k++] = lo.bn_g[l]; param_offsets[k++] = lo.bn_b[l];
    }
    param_offsets[k++] = lo.dense_w; param_offsets[k++] = lo.dense_b;
    param_offsets[k++] = lo.bn_g[lo.L]; param_offsets[k++] = lo.bn_b[lo.L];
    param_offsets[k++] = lo.head_w; param_offsets[k++] = lo.head_b;
  }
  if (bn_offsets) {
    int k = 0;
    for (int l = 0; l <= lo.L; ++l) { bn_offsets[k++] = lo.bn_mean[l]; bn_offsets[k++] = lo.bn_var[l]; }
  }
  if (n_params) *n_params = lo.n_params;
  if (n_bn) *n_bn = lo.n_bn;
  return DCGC_OK;
}

extern "C" int64_t dcgc_gcmodel_workspace_bytes(const dcgc_gcmodel_config* cfg, int64_t n_atoms, int64_t n_segments) {
  Layout lo;
  if (make_layout(cfg, &lo) != DCGC_OK || n_atoms < 0 || n_segments < 0) return -1;
  const int L = lo.L, D = cfg->dense;
  int wmax = D;
  int64_t per_atom = 0;  // floats per atom
  for (int l = 0; l < L; ++l) {
    const int c = cfg->widths[l];
    wmax = wmax > c ? wmax : c;
    per_atom += lo.fp[l] + 2 * c + (c + 3) / 4;  // S, Y, P, arg (bytes/4)
  }
  per_atom += D;                 // Z
  int fmax = 0;
  for (int l = 0; l <= L; ++l) fmax = fmax > lo.fp[l] ? fmax : lo.fp[l];
  per_atom += 2 * (int64_t)wmax + 2 * (int64_t)fmax;  // backward scratch: dA, dP, d1, d2
  int64_t bytes = n_atoms * per_atom * 4;
  bytes += n_segments * (int64_t)(2 * D * 2 + D + 2 * D) * 4;    // fp, dfp, argrow, zc_sum, zc_arg
  bytes += n_segments * (int64_t)cfg->n_out * 4 * 3;              // out, dout, per-element loss
  bytes += (int64_t)(part_chunks(n_atoms) + 1) * 2 * (int64_t)wmax * 8;
  bytes += ((n_segments + kHfRows - 1) / kHfRows + 1) * ((int64_t)cfg->n_out * (2 * D + 1) * 4 + 8) + 256;
  int64_t wg = 0;
  for (int l = 0; l < L; ++l) {
    const int64_t b = dcgc_group_gemm_wgrad_workspace(lo.fp[l], lo.fp[l], cfg->widths[l], DCGC_N_DEG);
    wg = wg > b ? wg : b;
  }
  const int64_t bd = dcgc_linear_wgrad_workspace(lo.f[L], D);
  wg = wg > bd ? wg : bd;
  bytes += wg;
  bytes += (int64_t)(L + 1) * 16 * wmax * 4 + (int64_t)L * DCGC_N_DEG * wmax * 8 + 3 * (int64_t)wmax * 4;
  const int nt = dcgc_tc_terms(cfg->gemm_mode);
  if (nt != 0) {   // early weight images of the train step
    for (int l = 0; l < L; ++l) {
      bytes += dcgc_tc_image_bytes(nt, lo.fp[l], lo.fp[l], cfg->widths[l], DCGC_N_DEG) + 256;
      if (l > 0) bytes += dcgc_tc_image_bytes(nt, cfg->widths[l], 0, 2 * lo.fp[l], DCGC_N_DEG) + 256;
    }
    bytes += dcgc_tc_image_bytes(nt, lo.f[L], 0, D, 1) + dcgc_tc_image_bytes(nt, D, 0, lo.f[L], 1) + 512;
  }
  return bytes + 256 * 64;  // alignment slack for every carve
}

extern "C" int dcgc_gcmodel_forward(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x,
                                    int64_t ld_x, int64_t n_samples, const float* params, float* bn_running,
                                    int32_t training, void* workspace, int64_t workspace_bytes, float* out,
                                    float* probs, float* fingerprint, void* stream) {
  Layout lo;
  RET_IF(make_layout(cfg, &lo));
  RET_IF(check_topo(topo));
  DCGC_CHECK_ARG(params && (workspace || workspace_bytes == 0), "dcgc_gcmodel_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Arena ws{(char*)workspace, 0, workspace_bytes};
  Saved sv{};
  RET_IF(forward_impl(cfg, lo, topo, x, ld_x, n_samples, params, bn_running, training, 0, ws, sv, st));
  const int64_t S = topo->n_segments;
  if (fingerprint && S > 0)
    DCGC_CUDA_CALL(cudaMemcpyAsync(fingerprint, sv.fp, (size_t)S * 2 * cfg->dense * 4, cudaMemcpyDeviceToDevice, st));
  if (out && n_samples > 0)
    DCGC_CUDA_CALL(cudaMemcpyAsync(out, sv.out, (size_t)n_samples * cfg->n_out * 4, cudaMemcpyDeviceToDevice, st));
  if (probs && n_samples > 0 && cfg->mode == 1) {
    const int64_t rows = n_samples * (cfg->n_out / cfg->n_classes);
    dcgc_launch(softmax_rows, blocks_for(rows), kT, 0, st, sv.out, rows, cfg->n_classes, probs);
    DCGC_CUDA_LAUNCH_CHECK("softmax_rows");
  }
  return DCGC_OK;
}

extern "C" int dcgc_gcmodel_train_step(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x,
                                       int64_t ld_x, const float* y, const float* w, int64_t n_samples,
                                       const float* params, float* grads, float* bn_running, void* workspace,
                                       int64_t workspace_bytes, float* loss_dev, float* out, void* stream) {
  return dcgc_gcmodel_train_step_ev(cfg, topo, x, ld_x, y, w, n_samples, params, grads, bn_running, workspace,
                                    workspace_bytes, loss_dev, out, nullptr, nullptr, 0, stream);
}

extern "C" int dcgc_gcmodel_train_step_ev(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x,
                                          int64_t ld_x, const float* y, const float* w, int64_t n_samples,
                                          const float* params, float* grads, float* bn_running, void* workspace,
                                          int64_t workspace_bytes, float* loss_dev, float* out, void* forward_event,
                                          void* const* grad_events, int32_t n_grad_events, void* stream) {
  return dcgc_gcmodel_train_step_sync(cfg, topo, x, ld_x, y, w, n_samples, params, grads, bn_running, workspace,
                                      workspace_bytes, loss_dev, out, forward_event, grad_events, n_grad_events, nullptr,
                                      stream);
}

extern "C" int dcgc_gcmodel_train_step_sync(const dcgc_gcmodel_config* cfg, const dcgc_topology* topo, const float* x,
                                            int64_t ld_x, const float* y, const float* w, int64_t n_samples,
                                            const float* params, float* grads, float* bn_running, void* workspace,
                                            int64_t workspace_bytes, float* loss_dev, float* out, void* forward_event,
                                            void* const* grad_events, int32_t n_grad_events, const dcgc_bn_sync* sync,
                                            void* stream) {
  Layout lo;
  RET_IF(make_layout(cfg, &lo));
  RET_IF(check_topo(topo));
  if (sync && (sync->world <= 1 || !cfg->batch_norm)) sync = nullptr;
  if (sync) {
    int wmax_s = cfg->dense;
    for (int l = 0; l < cfg->n_layers; ++l) wmax_s = wmax_s > cfg->widths[l] ? wmax_s : cfg->widths[l];
    DCGC_CHECK_ARG(sync->world <= DCGC_SYNC_MAX_RANKS && sync->rank >= 0 && sync->rank < sync->world && sync->seq0 > 0 &&
                       sync->cap >= wmax_s && wmax_s <= 1024,
                   "dcgc_gcmodel_train_step_sync: bad dcgc_bn_sync (world %d, rank %d, cap %d for width %d)", sync->world,
                   sync->rank, sync->cap, wmax_s);
    for (int r = 0; r < sync->world; ++r)
      DCGC_CHECK_ARG(sync->mailbox[r], "dcgc_gcmodel_train_step_sync: mailbox of rank %d is not mapped", r);
    // The exchange kernels wait for a peer inside the kernel.  With lazy module loading the FIRST launch of a kernel
    // may need the device idle: were that to happen behind a waiting exchange kernel of the same device (two ranks
    // sharing one GPU), the host would block and the peer's step would never be enqueued.  Load both variants now.
    static bool loaded = false;
    if (!loaded) {
      cudaFuncAttributes fa;
      DCGC_CUDA_CALL(cudaFuncGetAttributes(&fa, bn_sync_kernel<1>));
      DCGC_CUDA_CALL(cudaFuncGetAttributes(&fa, bn_sync_kernel<2>));
      loaded = true;
    }
  }
  DCGC_CHECK_ARG(params && grads && y && loss_dev && workspace, "dcgc_gcmodel_train_step: null pointer");
  DCGC_CHECK_ARG(n_grad_events >= 0 && n_grad_events <= cfg->n_layers + 1 && (n_grad_events == 0 || grad_events),
                 "dcgc_gcmodel_train_step_ev: at most n_layers + 1 gradient events");
  cudaStream_t st = (cudaStream_t)stream;
  auto slice_done = [&](int i) -> int {
    if (i < n_grad_events && grad_events[i]) DCGC_CUDA_CALL(cudaEventRecord((cudaEvent_t)grad_events[i], st));
    return DCGC_OK;
  };
  const dcgc_topology* t = topo;
  const int L = lo.L, D = cfg->dense, T = cfg->n_out;
  const int64_t N = t->n_atoms, S = t->n_segments;
  Arena ws{(char*)workspace, 0, workspace_bytes};
  Saved sv{};
  sv.sync = sync;
  // ---- weight images of all seven GEMMs, built now on the side stream (the weights only change in the Adam launch,
  // which precedes this call on `st`): they run beside the first gather-sum instead of in front of every GEMM
  const float* img_dgrad[DCGC_MODEL_MAX_LAYERS] = {};
  const float* img_dense_dgrad = nullptr;
  SideStream* side = nullptr;
  const int nt = dcgc_tc_terms(cfg->gemm_mode);
  if (nt != 0 && early_images_on() && N > 0) {
    RET_IF(side_stream(&side));
    float* imf[DCGC_MODEL_MAX_LAYERS]; float* imd[DCGC_MODEL_MAX_LAYERS] = {};
    for (int l = 0; l < L; ++l) {
      imf[l] = ws.take<float>(dcgc_tc_image_bytes(nt, lo.fp[l], lo.fp[l], cfg->widths[l], DCGC_N_DEG) / 4);
      if (l > 0) imd[l] = ws.take<float>(dcgc_tc_image_bytes(nt, cfg->widths[l], 0, 2 * lo.fp[l], DCGC_N_DEG) / 4);
    }
    float* im_dense = ws.take<float>(dcgc_tc_image_bytes(nt, lo.f[L], 0, D, 1) / 4);
    float* im_dense_d = ws.take<float>(dcgc_tc_image_bytes(nt, D, 0, lo.f[L], 1) / 4);
    if (!ws.ok) {
      dcgc_set_error("dcgc_gcmodel_train_step: workspace too small for the weight images");
      return DCGC_ERR_NOMEM;
    }
    DCGC_CUDA_CALL(cudaEventRecord(side->e0, st));
    DCGC_CUDA_CALL(cudaStreamWaitEvent(side->s, side->e0, 0));
    const bool f16 = fwd_f16x3_on(cfg);     // (the fp16 images are never larger than the tf32 ones they replace)
    // ONE launch for all of them (forward images first in block order); DCGC_IMAGE_BATCH=0: one launch per image
    static const bool batch = [] { const char* e = getenv("DCGC_IMAGE_BATCH"); return !(e && e[0] == '0'); }();
    DcgcImgJob jobs[2 * DCGC_MODEL_MAX_LAYERS + 2];
    int nj = 0;
    for (int l = 0; l < L; ++l)       // forward: [X | S] . W[g], W stored [G][2 fp][c]
      jobs[nj++] = DcgcImgJob{params + lo.conv_w[l], DCGC_N_DEG, 1, lo.fp[l], lo.fp[l], cfg->widths[l], imf[l], f16 ? 1 : 0};
    jobs[nj++] = DcgcImgJob{params + lo.dense_w, 1, 0, lo.f[L], 0, D, im_dense, f16 ? 1 : 0};     // nn.Linear layout
    const int n_fwd = nj;
    jobs[nj++] = DcgcImgJob{params + lo.dense_w, 1, 1, D, 0, lo.f[L], im_dense_d, 0};             // dx = g . w
    for (int l = L - 1; l >= 1; --l)  // dgrad: G . W[g]^T, the same W read as [G][2 fp][c] = [n1 + n2][k1]
      jobs[nj++] = DcgcImgJob{params + lo.conv_w[l], DCGC_N_DEG, 0, cfg->widths[l], 0, 2 * lo.fp[l], imd[l], 0};
    if (batch) {
      RET_IF(dcgc_tc_prep_weights_batch(nt, jobs, nj, side->s));
      DCGC_CUDA_CALL(cudaEventRecord(side->e1, side->s));
    } else {
      for (int j = 0; j < nj; ++j) {
        const DcgcImgJob& q = jobs[j];
        if (q.f16) RET_IF(dcgc_tc_prep_weights_f16(q.w, q.n_groups, q.trans_w, q.k1, q.k2, q.N, q.img, side->s));
        else RET_IF(dcgc_tc_prep_weights(nt, q.w, q.n_groups, q.trans_w, q.k1, q.k2, q.N, q.img, side->s));
        if (j == n_fwd - 1) DCGC_CUDA_CALL(cudaEventRecord(side->e1, side->s));
      }
    }
    DCGC_CUDA_CALL(cudaEventRecord(side->e2, side->s));
    for (int l = 0; l < L; ++l) { sv.img_fwd[l] = imf[l]; img_dgrad[l] = imd[l]; }
    sv.img_dense = im_dense;
    img_dense_dgrad = im_dense_d;
    sv.img_fwd_ready = side->e1;
  }
  if (cfg->batch_norm && nt != 0 && N > 0 && last_cta_fin_on(N) && !sync) {
    // counters of the last-CTA BatchNorm finalizes (forward 2 idx, backward 2 idx + 1): zero at every launch
    sv.fin_counters = ws.take<unsigned int>(2 * (DCGC_MODEL_MAX_LAYERS + 1));
    if (!ws.ok) {
      dcgc_set_error("dcgc_gcmodel_train_step: workspace too small");
      return DCGC_ERR_NOMEM;
    }
    DCGC_CUDA_CALL(cudaMemsetAsync(sv.fin_counters, 0, 2 * (DCGC_MODEL_MAX_LAYERS + 1) * sizeof(unsigned int), st));
  }
  const bool fused_head = head_fused_ok(cfg);
  RET_IF(forward_impl(cfg, lo, t, x, ld_x, n_samples, params, bn_running, 1, 1, ws, sv, st, fused_head));
  if (side) DCGC_CUDA_CALL(cudaStreamWaitEvent(st, side->e2, 0));     // the backward's images (long done by now)
  if (forward_event) DCGC_CUDA_CALL(cudaEventRecord((cudaEvent_t)forward_event, st));

  // ---- backward scratch
  int wmax = D, fmax = 0;
  for (int l = 0; l < L; ++l) wmax = wmax > cfg->widths[l] ? wmax : cfg->widths[l];
  for (int l = 0; l <= L; ++l) fmax = fmax > lo.fp[l] ? fmax : lo.fp[l];
  float* dA = ws.take<float>(N * wmax);   // grad wrt BN output / then GEMM gradient G (in place)
  float* dP = ws.take<float>(N * wmax);   // grad wrt pool output of the layer below
  float* d2 = ws.take<float>(N * fmax);   // neighbour-path input gradient before the transposed gather
  float* dfp = ws.take<float>(S * 2 * D);
  float* dout = ws.take<float>((n_samples > 0 ? n_samples : 1) * T);
  const int n_loss_elems_per_row = cfg->mode == 1 ? T / cfg->n_classes : T;
  float* per_elem = ws.take<float>((n_samples > 0 ? n_samples : 1) * n_loss_elems_per_row);
  const int head_chunks = (int)((n_samples + kHeadChunk - 1) / kHeadChunk);
  const int head_blocks = (int)((S + kHfRows - 1) / kHfRows);
  const int64_t part_rows = fused_head ? head_blocks : head_chunks;
  float* head_part = ws.take<float>((part_rows > 0 ? part_rows : 1) * T * (2 * D + 1));
  double* loss_part = ws.take<double>(head_blocks > 0 ? head_blocks : 1);
  float* coef = ws.take<float>(3 * (int64_t)wmax);
  float* db11 = ws.take<float>(DCGC_N_DEG * (int64_t)wmax);
  int64_t wg_bytes = dcgc_linear_wgrad_workspace(lo.f[L], D);
  for (int l = 0; l < L; ++l) {
    const int64_t b = dcgc_group_gemm_wgrad_workspace(lo.fp[l], lo.fp[l], cfg->widths[l], DCGC_N_DEG);
    wg_bytes = wg_bytes > b ? wg_bytes : b;
  }
  char* wg = ws.take<char>(wg_bytes);
  if (!ws.ok) {
    dcgc_set_error("dcgc_gcmodel_train_step: workspace too small (%lld bytes needed, %lld given)", (long long)ws.off,
                   (long long)ws.cap);
    return DCGC_ERR_NOMEM;
  }

  // ---- loss and head
  const int64_t n_elems = n_samples * n_loss_elems_per_row;
  {
  DcgcProfScope prof_scope("head_loss_bwd", st);
  if (fused_head) {
    static bool attr_done = false;       // per process; 65 KB at the widest supported head
    if (!attr_done) {
      DCGC_CUDA_CALL(cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_done = true;
    }
    if (head_blocks > 0) {
      dcgc_launch(head_fused_kernel, head_blocks, kT, head_fused_smem(cfg), st, 
          sv.fp, 2 * D, params + lo.head_w, params + lo.head_b, y, w, n_samples, S, 2 * D, T,
          cfg->mode == 1 ? cfg->n_classes : 1, cfg->mode, n_elems, sv.out, dfp, head_part, loss_part);
      DCGC_CUDA_LAUNCH_CHECK("head_fused_kernel");
    }
    dcgc_launch(head_fused_final, blocks_for(((int64_t)T * (2 * D + 1) + 1) * 32), kT, 0, st, head_part, loss_part, head_blocks, 2 * D, T,
                                                                          n_elems, grads + lo.head_w,
                                                                          grads + lo.head_b, loss_dev);
    DCGC_CUDA_LAUNCH_CHECK("head_fused_final");
    if (out && n_samples > 0)
      DCGC_CUDA_CALL(cudaMemcpyAsync(out, sv.out, (size_t)n_samples * T * 4, cudaMemcpyDeviceToDevice, st));
  } else {
  if (n_elems > 0) {
    dcgc_launch(loss_fwd_bwd, blocks_for(n_elems), kT, 0, st, sv.out, y, w, n_elems, cfg->mode == 1 ? cfg->n_classes : 1,
                                                     cfg->mode, per_elem, dout);
    DCGC_CUDA_LAUNCH_CHECK("loss_fwd_bwd");
  }
  dcgc_launch(loss_reduce, 1, 1024, 0, st, per_elem, n_elems, loss_dev);
  DCGC_CUDA_LAUNCH_CHECK("loss_reduce");
  if (out && n_samples > 0)
    DCGC_CUDA_CALL(cudaMemcpyAsync(out, sv.out, (size_t)n_samples * T * 4, cudaMemcpyDeviceToDevice, st));
  if (head_chunks > 0) {
    dim3 grid((unsigned)head_chunks, (unsigned)T);
    dcgc_launch(head_bwd_partial, grid, kT, 0, st, dout, sv.fp, 2 * D, n_samples, 2 * D, T, head_part);
    DCGC_CUDA_LAUNCH_CHECK("head_bwd_partial");
  }
  dcgc_launch(head_bwd_final, blocks_for((int64_t)T * (2 * D + 1)), kT, 0, st, head_part, head_chunks, 2 * D, T,
                                                                      grads + lo.head_w, grads + lo.head_b);
  DCGC_CUDA_LAUNCH_CHECK("head_bwd_final");
  if (S > 0) {
    dcgc_launch(head_bwd_input, blocks_for(S * 2 * D), kT, 0, st, dout, params + lo.head_w, n_samples, S, 2 * D, T, dfp);
    DCGC_CUDA_LAUNCH_CHECK("head_bwd_input");
  }
  }
  }
  // backward finalize of BatchNorm idx by the last CTA of the kernel that writes dA and its column sums
  auto bwd_fin = [&](int idx, int width) -> DcgcBnFin {
    DcgcBnFin f{};
    if (!sv.fin_counters || !cfg->batch_norm || width > 256) return f;
    const float* stats = sv.stats + sv.stats_off[idx];
    f.counter = sv.fin_counters + 2 * idx + 1; f.kind = 2; f.width = width; f.n_rows = N; f.part = sv.part;
    f.mean = stats; f.invstd = stats + width; f.scale = stats + 2 * width;
    f.dgamma = grads + lo.bn_g[idx]; f.dbeta = grads + lo.bn_b[idx]; f.coef = coef;
    return f;
  };
  // fused_chunks >= 0: the stage-1 partials were already written by the kernel that produced dA
  // (dcgc_mg_pool_bwd_stats), one row per CTA
  auto bn_backward = [&](int idx, const float* yv, int width, int fused_chunks, bool finalized = false,
                         bool skip_apply = false) -> int {
    // dA (ld = width) -> G in place; dgamma / dbeta into the gradient slab
    if (cfg->batch_norm) {
      const float* stats = sv.stats + sv.stats_off[idx];
      {
      DcgcProfScope prof_scope("bn_stats_bwd", st);
      if (N > 0 && fused_chunks < 0) {
        dim3 grid((unsigned)sv.n_chunks, (unsigned)((width + 127) / 128));
        dcgc_launch(col_moments_partial, grid, 32 * kMomLanes, 0, st, dA, width, yv, width, N, width,
                                                             mom_rows(N, sv.n_chunks), sv.part);
        DCGC_CUDA_LAUNCH_CHECK("col_moments_partial (bwd)");
      }
      if (sv.sync) {
        const int chunks = fused_chunks >= 0 ? fused_chunks : (N > 0 ? sv.n_chunks : 0);
        float* stats_w = sv.stats + sv.stats_off[idx];
        dcgc_launch(bn_sync_kernel<2>, 1, 512, 0, st, 
            sync_args(sv.sync, sv.sync->seq0 + (unsigned long long)(L + 1 + (L - idx))), sv.part, chunks, width,
            (long long)N, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, stats_w, stats_w + width, stats_w + 2 * width,
            nullptr, grads + lo.bn_g[idx], grads + lo.bn_b[idx], coef);
        DCGC_CUDA_LAUNCH_CHECK("bn_sync_kernel (backward)");
      } else if (!(finalized && fused_chunks >= 0)) {
        dcgc_launch(bn_bwd_finalize, (width + kFinCols - 1) / kFinCols, 32 * kFinCols, 0, st, sv.part, fused_chunks >= 0 ? fused_chunks : (N > 0 ? sv.n_chunks : 0), width, N, stats,
                                                             stats + width, stats + 2 * width, grads + lo.bn_g[idx],
                                                             grads + lo.bn_b[idx], coef);
        DCGC_CUDA_LAUNCH_CHECK("bn_bwd_finalize");
      }
      }
      if (skip_apply) return DCGC_OK;     // the consumer applies the coefficients itself (dcgc_gather_bwd_apply)
      DcgcProfScope prof_scope("bn_relu_bwd_apply", st);
      if (N > 0) {
        const int groups = width / 4;
        DCGC_CHECK_ARG(groups <= kT, "dcgc_gcmodel: layer width above %d is not supported by bn_relu_bwd_apply", 4 * kT);
        const int64_t rows_per_block = (int64_t)(kT / groups) * kApplyRows;
        dcgc_launch(bn_relu_bwd_apply, (unsigned)((N + rows_per_block - 1) / rows_per_block), kT, 0, st, 
            dA, width, yv, width, stats, stats + width, coef, N, width, 1, dA, width);
        DCGC_CUDA_LAUNCH_CHECK("bn_relu_bwd_apply");
      }
    } else if (N > 0) {
      dcgc_launch(relu_bwd_apply, blocks_for(N * (width / 4)), kT, 0, st, dA, width, yv, width, N, width, dA, width);
      DCGC_CUDA_LAUNCH_CHECK("relu_bwd_apply");
    }
    return DCGC_OK;
  };

  // ---- GraphGather backward -> dA (grad wrt the BN output of the dense layer)
  // With BatchNorm and the molecule-level sums of the forward pass: column sums from 4 096 molecules instead of 102 k
  // atoms, finalize, then ONE kernel writes G = relu'(z) * BatchNorm-backward(dA) without dA ever being stored
  // (DCGC_NO_DENSE_FUSION=1: the previous route, dA + sums in one kernel, then the apply pass).
  static const bool no_dense_fusion = [] { const char* e = getenv("DCGC_NO_DENSE_FUSION"); return e && e[0] == '1'; }();
  const bool al16 = ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(sv.z) | reinterpret_cast<uintptr_t>(dfp) |
                      reinterpret_cast<uintptr_t>(sv.fp) | reinterpret_cast<uintptr_t>(sv.argrow) |
                      reinterpret_cast<uintptr_t>(sv.stats + sv.stats_off[L]) | reinterpret_cast<uintptr_t>(coef)) & 15) == 0;
  bool dense_done = false;
  if (cfg->batch_norm && sv.zc_sum && !no_dense_fusion && D % 4 == 0 && al16 && N > 0 && S > 0) {
    const float* stats_d = sv.stats + sv.stats_off[L];
    int32_t ch = 0;
    RET_IF(dcgc_dense_bn_sums(dfp, 2 * D, sv.fp, 2 * D, sv.argrow, t->mol_ptr, S, D, DCGC_ACT_TANH, sv.zc_sum, sv.zc_arg,
                              stats_d, sv.part, &ch, st));
    RET_IF(bn_backward(L, sv.z, D, ch, false, true));
    RET_IF(dcgc_gather_bwd_apply(dfp, 2 * D, sv.fp, 2 * D, sv.argrow, t->membership, N, D, DCGC_ACT_TANH, sv.z, D, stats_d,
                                 stats_d + D, coef, dA, D, st));
    dense_done = true;
  }
  int32_t fused_gather = -1;
  DcgcBnFin fin_gather{};
  if (dense_done) {
  } else if (cfg->batch_norm && N > 0 && D % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(sv.z) | reinterpret_cast<uintptr_t>(dfp) |
        reinterpret_cast<uintptr_t>(sv.fp) | reinterpret_cast<uintptr_t>(sv.argrow)) & 15) == 0)
  {
    fin_gather = bwd_fin(L, D);
    RET_IF(dcgc_gather_bwd_stats(dfp, 2 * D, sv.fp, 2 * D, sv.argrow, t->membership, N, D, DCGC_ACT_TANH, dA, D, sv.z, D,
                                 part_chunks(N), sv.part, &fused_gather, fin_gather.kind ? &fin_gather : nullptr, st));
  }
  else
    RET_IF(dcgc_gather_bwd(dfp, 2 * D, sv.fp, 2 * D, sv.argrow, t->membership, N, D, DCGC_ACT_TANH, dA, D, st));

  // ---- dense layer backward
  if (!dense_done) RET_IF(bn_backward(L, sv.z, D, fused_gather, fin_gather.kind != 0));
  RET_IF(dcgc_linear_wgrad(cfg->gemm_mode, sv.h[L], sv.ld_h[L], lo.f[L], dA, D, D, N, grads + lo.dense_w,
                           grads + lo.dense_b, wg, wg_bytes, st));
  RET_IF(slice_done(0));            // head, dense layer and its BatchNorm
  {
    DcgcGemmOpts go;
    go.img = img_dense_dgrad;
    RET_IF(dcgc_linear_dgrad_opts(cfg->gemm_mode, dA, D, D, params + lo.dense_w, lo.f[L], N, dP, lo.f[L], go, st));
  }

  // ---- conv stack backward
  for (int l = L - 1; l >= 0; --l) {
    const int c = cfg->widths[l], fp = lo.fp[l];
    // GraphPool backward over CSR^T: dP (ld c) -> dA (ld c)
    int32_t fused = -1;
    DcgcBnFin fin_pool{};
    if (use_mg(t, c, c, dP, sv.arg[l]) && (reinterpret_cast<uintptr_t>(dA) & 15) == 0) {
      // the BatchNorm-backward column sums of dA come out of the same kernel (DCGC_NO_FUSED_BN_BWD=1: separate pass)
      static const bool no_fuse = [] { const char* e = getenv("DCGC_NO_FUSED_BN_BWD"); return e && e[0] == '1'; }();
      if (cfg->batch_norm && !no_fuse && c % 4 == 0 && (reinterpret_cast<uintptr_t>(sv.y[l]) & 15) == 0)
      {
        fin_pool = bwd_fin(l, c);
        RET_IF(dcgc_mg_pool_bwd_stats_fin(dP, c, sv.arg[l], c, t, c, dA, c, sv.y[l], c, sv.stats + sv.stats_off[l], sv.part,
                                          &fused, fin_pool.kind ? &fin_pool : nullptr, st));
      }
      else
        RET_IF(dcgc_mg_pool_bwd(dP, c, sv.arg[l], c, nullptr, t, c, dA, c, st));
    } else {
      RET_IF(dcgc_pool_bwd(dP, c, sv.arg[l], c, nullptr, t->t_row_ptr, t->t_src, t->t_slot, N, c, dA, c, st));
    }
    RET_IF(bn_backward(l, sv.y[l], c, fused, fin_pool.kind != 0));
    // (the reduction of the weight-gradient partials also scatters the 11 per-degree bias sums to the reference's 21
    // bias rows)
    RET_IF(dcgc_group_gemm_wgrad_opts(cfg->gemm_mode, sv.h[l], sv.ld_h[l], fp, sv.s[l], fp, fp, dA, c, c, t->deg_count,
                                      DCGC_N_DEG, grads + lo.conv_w[l], db11, wg, wg_bytes,
                                      (l == 0 && cfg->input_exact) ? 1 : 0, st, grads + lo.conv_b[l]));
    RET_IF(slice_done(1 + (L - 1 - l)));
    if (l > 0) {
      // [dP | d2] = G . W^T, then dP += transposed gather of d2
      DcgcGemmOpts go;
      go.img = img_dgrad[l];
      RET_IF(dcgc_group_gemm_dgrad_opts(cfg->gemm_mode, dA, c, c, params + lo.conv_w[l], fp, fp, t->tiles, t->n_tiles,
                                        128, N, dP, fp, d2, fp, go, st));
      if (use_mg(t, fp, 0, d2, dP))
        RET_IF(dcgc_mg_gather_sum(d2, fp, t, 1, fp, dP, fp, dP, fp, st));
      else if (t->symmetric)
        RET_IF(dcgc_gather_sum_bucketed(d2, fp, t->deg_count, t->t_src, N, fp, dP, fp, dP, fp, st));
      else
        RET_IF(dcgc_gather_sum(d2, fp, t->t_row_ptr, t->t_src, N, fp, dP, fp, dP, fp, st));
    }
  }
  return DCGC_OK;
}

extern "C" int dcgc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                              void* stream) {
  DCGC_CHECK_ARG(n >= 0 && step >= 1, "dcgc_adam_step: bad sizes (step counts from 1)");
  if (n == 0) return DCGC_OK;
  DCGC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "dcgc_adam_step: null pointer");
  DcgcProfScope prof_scope("dcgc_adam_step", (cudaStream_t)stream);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 ? n / 4 : 0;
  if (n4 > 0)
    dcgc_launch(adam_kernel_vec, blocks_for(n4), kT, 0, (cudaStream_t)stream, 
        reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(exp_avg),
        reinterpret_cast<float4*>(exp_avg_sq), n4, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
  if (n - 4 * n4 > 0)      // the tail (or everything, for unaligned slabs): identical arithmetic per element
    dcgc_launch(adam_kernel, blocks_for(n - 4 * n4), kT, 0, (cudaStream_t)stream, params + 4 * n4, grads + 4 * n4, exp_avg + 4 * n4,
                                                                       exp_avg_sq + 4 * n4, n - 4 * n4, lr, beta1, beta2,
                                                                       eps, (float)bc1, (float)sqrt(bc2), grad_scale);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_adam_step");
  return DCGC_OK;
}

// Internal helpers shared by the libdcgc translation units (not part of the ABI).
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "dcgc.h"

void dcgc_set_error(const char* fmt, ...);

#define DCGC_CHECK_ARG(cond, ...)   \
  do {                              \
    if (!(cond)) {                  \
      dcgc_set_error(__VA_ARGS__);  \
      return DCGC_ERR_INVALID;      \
    }                               \
  } while (0)

#ifdef __CUDACC__
#include <cuda_runtime.h>

#include <atomic>
extern std::atomic<long long> g_dcgc_launches;
// RAII: if profiling of entry point `name` is on, brackets the enclosed launches with CUDA events
struct DcgcProfScope {
  DcgcProfScope(const char* name, cudaStream_t st);
  ~DcgcProfScope();
  cudaEvent_t e1_;
  cudaStream_t st_;
};
// tensor-core modes -> number of MMA terms (0 = not a tensor-core mode)
static inline int dcgc_tc_terms(int mode) {
  return (mode == DCGC_GEMM_TF32X3 || mode == DCGC_GEMM_F16X3) ? 3 : (mode == DCGC_GEMM_BF16 ? 1 : 0);
}
#define DCGC_CUDA_LAUNCH_CHECK(what)                                        \
  do {                                                                      \
    g_dcgc_launches.fetch_add(1, std::memory_order_relaxed);                \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      dcgc_set_error("%s: %s", what, cudaGetErrorString(e__));              \
      return DCGC_ERR_CUDA;                                                 \
    }                                                                       \
  } while (0)
#define DCGC_CUDA_CALL(expr)                                                \
  do {                                                                      \
    cudaError_t e__ = (expr);                                               \
    if (e__ != cudaSuccess) {                                               \
      dcgc_set_error("%s: %s", #expr, cudaGetErrorString(e__));             \
      return DCGC_ERR_CUDA;                                                 \
    }                                                                       \
  } while (0)
#endif

#ifdef __CUDACC__
// stage-1 arguments of the weight-gradient contraction (shared by the SIMT and tcgen05 kernels):
// chunk c of group g covers rows [group_row0[g] + (c - chunk_prefix[g]) * chunk_rows, ...) and writes
// the partial dW to ws[c][k1+k2][n] and the partial column sums of grad to wsb[c][n]
struct DcgcWgradArgs {
  const float* a1; int64_t ld_a1; int k1;
  const float* a2; int64_t ld_a2; int k2;
  const float* g; int64_t ld_g; int n;
  float* ws; float* wsb;
  int chunk_rows, n_groups, tiles_n;
  int64_t group_row0[DCGC_N_DEG + 1];
  int chunk_prefix[DCGC_N_DEG + 1];
  int a1_vec, a2_vec, g_vec;
  int a_exact;      // [a1 | a2] holds tf32-exact values (see DcgcGemmOpts): the lo(A) tiles and their MMA are skipped
  long long* dbg;   // optional timeline buffer (dcgcdbg_tc_timeline)
  int knob;         // measurement switches of tc_wgrad_kernel_v2 (env DCGC_WG2_KNOB), 0 in production
};
struct DcgcBnFin;
// Explicit per-call options of the tensor-core GEMMs (no thread-local side channels):
//   img      a weight image built ahead with dcgc_tc_prep_weights (e.g. on another stream); null = the call builds
//            its own in a stream-ordered scratch allocation (cudaMallocAsync / cudaFreeAsync on the launch stream)
//   a_exact  every value of the A operand is exactly representable in tf32 (integer-valued atom features and
//            their neighbour sums below 2^11): its lo part is identically zero, so the lo(A) tile and the
//            lo(A) * hi(W) term are skipped — a third of the tensor time, bit-identical results
struct DcgcGemmOpts {
  const float* img = nullptr;
  int a_exact = 0;
  const DcgcBnFin* fin = nullptr;   // BatchNorm finalize by the last CTA (only with the fused column statistics)
  int f16x3 = 0;                    // forward GEMMs: fp16 hi / lo halves instead of tf32 ones (tc_gemm_kernel_v6); img, if
                                    // given, must then come from dcgc_tc_prep_weights_f16
};
int dcgc_gather_bwd_stats(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const int32_t* argrow,
                          const int32_t* membership, int64_t n_rows, int32_t width, int32_t act, float* dx,
                          int64_t ld_dx, const float* z, int64_t ld_z, int max_chunks, double* part,
                          int32_t* n_chunks_out, const DcgcBnFin* fin, void* stream);
int dcgc_gather_fwd_train(const float* x, int64_t ld_x, const float* scale, const float* shift, const int32_t* mol_ptr,
                          const int32_t* mol_atoms, int64_t n_segments, int32_t width, int32_t act, float* out,
                          int64_t ld_out, int32_t* argrow, const float* mean, float* zc_sum, float* zc_arg,
                          void* stream);
int dcgc_dense_bn_sums(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const int32_t* argrow,
                       const int32_t* mol_ptr, int64_t n_segments, int32_t width, int32_t act, const float* zc_sum,
                       const float* zc_arg, const float* mean, double* part, int32_t* n_chunks_out, void* stream);
int dcgc_gather_bwd_apply(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const int32_t* argrow,
                          const int32_t* membership, int64_t n_rows, int32_t width, int32_t act, const float* z,
                          int64_t ld_z, const float* mean, const float* invstd, const float* coef, float* g,
                          int64_t ld_g, void* stream);
int dcgc_mg_pool_bwd_stats_fin(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t ld_arg, const dcgc_topology* t,
                               int32_t width, float* dx, int64_t ld_dx, const float* y, int64_t ld_y, const float* stats,
                               double* part, int32_t* n_chunks, const DcgcBnFin* fin, void* stream);
int dcgc_tc_wgrad_stage1(int nt, const DcgcWgradArgs& p, int chunks, cudaStream_t st);
int dcgc_tc_wgrad_grid_y(int k_total, int n);
int dcgc_tc_num_sms();
// tcgen05 path (gemm_tc.cu); see the comment there for the argument convention
int dcgc_tc_gemm(int nt, const float* a1, int64_t ld_a1, int k1, const float* a2, int64_t ld_a2, int k2, const float* w,
                 int n_groups, int trans_w, const float* bias, int n1, int n2, const int32_t* tiles, int64_t n_tiles,
                 int64_t n_rows, int act, float* c1, int64_t ld_c1, float* c2, int64_t ld_c2, cudaStream_t st,
                 double* stats = nullptr, int* stats_chunks = nullptr, const DcgcGemmOpts* opts = nullptr);
// early weight images (gemm_tc.cu): build on any stream, pass to the GEMM in DcgcGemmOpts::img
int64_t dcgc_tc_image_bytes(int nt, int k1, int k2, int N, int n_groups);
int dcgc_tc_prep_weights(int nt, const float* w, int n_groups, int trans_w, int k1, int k2, int N, float* img, cudaStream_t st);
int64_t dcgc_tc_image_bytes_f16(int k1, int k2, int N, int n_groups);
int dcgc_tc_prep_weights_f16(const float* w, int n_groups, int trans_w, int k1, int k2, int N, float* img, cudaStream_t st);
// several images in one launch (arguments as dcgc_tc_prep_weights / _f16 per job; f16 jobs need nt == 3)
struct DcgcImgJob { const float* w; int n_groups, trans_w, k1, k2, N; float* img; int f16; };
int dcgc_tc_prep_weights_batch(int nt, const DcgcImgJob* jobs, int n_jobs, cudaStream_t st);
// the C-ABI entry points of gemm_simt.cu with explicit options (used by the fused engines)
int dcgc_group_gemm_fwd_opts(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2, int64_t ld_a2,
                             int32_t k2, const float* w, const float* bias, int32_t n, const int32_t* tiles,
                             int64_t n_tiles, int32_t tile_rows, int64_t n_rows, int32_t act, float* y, int64_t ld_y,
                             double* stats_part, int32_t* n_chunks_out, const DcgcGemmOpts& opts, void* stream);
int dcgc_group_gemm_dgrad_opts(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w, int32_t k1,
                               int32_t k2, const int32_t* tiles, int64_t n_tiles, int32_t tile_rows, int64_t n_rows,
                               float* d1, int64_t ld_d1, float* d2, int64_t ld_d2, const DcgcGemmOpts& opts, void* stream);
int dcgc_linear_fwd_opts(int32_t mode, const float* x, int64_t ld_x, int32_t k, const float* w, const float* bias,
                         int32_t n, int64_t n_rows, int32_t act, float* y, int64_t ld_y, double* stats_part,
                         int32_t* n_chunks_out, const DcgcGemmOpts& opts, void* stream);
int dcgc_linear_dgrad_opts(int32_t mode, const float* g, int64_t ld_g, int32_t n, const float* w, int32_t k,
                           int64_t n_rows, float* dx, int64_t ld_dx, const DcgcGemmOpts& opts, void* stream);
int dcgc_group_gemm_wgrad_opts(int32_t mode, const float* a1, int64_t ld_a1, int32_t k1, const float* a2, int64_t ld_a2,
                               int32_t k2, const float* g, int64_t ld_g, int32_t n, const int64_t* deg_count,
                               int32_t n_groups, float* dw, float* dbias, void* workspace, int64_t workspace_bytes,
                               int a_exact, void* stream, float* dbias21 = nullptr);
#endif

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+): every kernel of the library starts with dcgc_griddep_wait() — wait until the
// kernel in front of it on the stream has completed and its writes are visible (a no-op for a launch without the
// attribute), then allow the NEXT kernel to be launched — and is launched through dcgc_launch with
// programmatic stream serialization: the next grid is set up and its CTAs are placed as SMs free up while this one
// still runs, instead of after it has drained.  Nothing before the wait may touch global memory.  DCGC_PDL=0: plain
// stream-ordered launches.
__device__ __forceinline__ void dcgc_griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// the wait alone: for a kernel that itself WAITS FOR ANOTHER STREAM OR DEVICE (bn_sync_kernel) — the grid behind it must
// not be placed on the SMs while it spins, or two ranks emulated on one device could starve each other
__device__ __forceinline__ void dcgc_griddep_wait_only() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool dcgc_pdl_on();   // profile.cu
template <typename... KArgs, typename... Args>
inline void dcgc_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = dcgc_pdl_on() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // (errors surface in DCGC_CUDA_LAUNCH_CHECK)
}

// mbarrier wait shared by the tcgen05 GEMMs and the molecule-group kernels.  The wait itself is the hardware's
// (try_wait suspends the thread); the guard around it is WALL-CLOCK based (%globaltimer, 20 s): a protocol bug still
// traps instead of hanging the GPU, but a context that is preempted, time-sliced (MPS) or replayed by a profiler is
// not killed by a spin count it happened to exceed.
__device__ __forceinline__ void dcgc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spin & 0xfffu) == 0xfffu) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}
#endif

#ifdef __CUDACC__
// ---- BatchNorm finalize by the LAST CTA of the kernel that produced the column-sum partials ---------------------------
// The kernels that emit per-CTA partial sums (GEMM epilogue: sum y, sum y^2; GraphPool / GraphGather backward: sum dA,
// sum dA y) used to be followed by a one-block-per-16-columns finalize kernel: 4.5-5.6 us each, eight per training
// step, nothing but launch and memory latency.  With a DcgcBnFin the CTAs count themselves on `counter` after their
// partials are written (threadfence + atomic: the classic last-block pattern); the last one adds the partial rows in
// ROW ORDER (deterministic, whichever CTA it is) and writes what the finalize kernel wrote.  counter must be zero at
// launch; the last CTA resets it.  kind 0 = off.
struct DcgcBnFin {
  unsigned int* counter;
  int kind;                 // 1: forward statistics -> mean / invstd / scale / shift + running statistics
                            // 2: backward sums -> dgamma / dbeta + the three coefficients of the apply pass
  int width;
  long long n_rows;
  const double* part;       // [chunks][2][width]
  const float* gamma; const float* beta; float eps, momentum; float* running_mean; float* running_var;
  float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
  const float* mean; const float* invstd; const float* scale; float* dgamma; float* dbeta; float* coef;
};
// Called by every thread of the group that wrote the partials (t = 0 .. n - 1 inside named barrier `bar`), after the
// writes; n_chunks = partial rows, n_ctas = CTAs of the grid that make this call; scratch = shared memory the group no
// longer needs, at least 2 * (n / width) * width doubles (n >= width).
// The last CTA pulls the whole partial table (chunks x 2 x width doubles, 300 KB at the bench shape) through ONE SM,
// so the loads must be spread over every thread and kept 32 deep: thread t owns column t % width and the
// (t / width)-th slice of the rows; the slices are then added in slice order.  (A first version with one thread per
// column and 16 loads in flight needed ~9 us — more than the finalize kernel it replaced.)
__device__ __forceinline__ void dcgc_bn_fin_last_cta(const DcgcBnFin& f, int n_chunks, unsigned n_ctas, int bar, int n, int t,
                                                     double* scratch) {
  if (f.kind == 0) return;
  __shared__ int last_flag;
  __threadfence();
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory");
  if (t == 0) last_flag = atomicAdd(f.counter, 1u) == n_ctas - 1u;
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory");
  if (!last_flag) return;
  __threadfence();
  const int slices = n / f.width;
  const int per = (n_chunks + slices - 1) / slices;
  if (t < slices * f.width) {
    const int c = t % f.width, sl = t / f.width;
    const int k0 = sl * per, k1 = min(n_chunks, k0 + per);
    double a = 0.0, b = 0.0;
    int k = k0;
    for (; k + 16 <= k1; k += 16) {
      double va[16], vb[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        va[u] = __ldcg(f.part + ((size_t)(k + u) * 2) * f.width + c);
        vb[u] = __ldcg(f.part + ((size_t)(k + u) * 2 + 1) * f.width + c);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) { a += va[u]; b += vb[u]; }
    }
    {
      double va[16], vb[16];                   // the ragged tail, still one round of loads
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        va[u] = k + u < k1 ? __ldcg(f.part + ((size_t)(k + u) * 2) * f.width + c) : 0.0;
        vb[u] = k + u < k1 ? __ldcg(f.part + ((size_t)(k + u) * 2 + 1) * f.width + c) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) { a += va[u]; b += vb[u]; }
    }
    scratch[(size_t)(2 * sl) * f.width + c] = a;
    scratch[(size_t)(2 * sl + 1) * f.width + c] = b;
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory");
  const double rows = (double)f.n_rows;
  for (int c = t; c < f.width; c += n) {
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < slices; ++sl) {
      a += scratch[(size_t)(2 * sl) * f.width + c];
      b += scratch[(size_t)(2 * sl + 1) * f.width + c];
    }
    if (f.kind == 1) {
      // torch semantics (graphconvmodel.py:150-158): normalise with the biased batch variance, running_var unbiased
      const double mean = rows > 0 ? a / rows : 0.0;
      double var = rows > 0 ? b / rows - mean * mean : 0.0;
      if (var < 0) var = 0;
      const double invstd = 1.0 / sqrt(var + (double)f.eps);
      const float g = f.gamma ? f.gamma[c] : 1.f, be = f.beta ? f.beta[c] : 0.f;
      const float sc = (float)(g * invstd);
      f.mean_out[c] = (float)mean;
      f.invstd_out[c] = (float)invstd;
      f.scale_out[c] = sc;
      f.shift_out[c] = (float)(be - mean * (double)sc);
      if (f.running_mean) {
        const double unbiased = rows > 1 ? var * rows / (rows - 1.0) : var;
        f.running_mean[c] = (float)((1.0 - f.momentum) * f.running_mean[c] + f.momentum * mean);
        f.running_var[c] = (float)((1.0 - f.momentum) * f.running_var[c] + f.momentum * unbiased);
      }
    } else {
      const double dg = (double)f.invstd[c] * (b - (double)f.mean[c] * a);      // sum dA * xhat
      if (f.dgamma) f.dgamma[c] = (float)dg;
      if (f.dbeta) f.dbeta[c] = (float)a;
      f.coef[c] = f.scale[c];
      f.coef[f.width + c] = rows > 0 ? (float)(a / rows) : 0.f;
      f.coef[2 * f.width + c] = rows > 0 ? (float)(dg / rows) : 0.f;
    }
  }
  if (t == 0) *f.counter = 0u;
}
#endif

static inline int64_t dcgc_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

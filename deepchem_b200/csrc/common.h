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
static inline int dcgc_tc_terms(int mode) { return mode == DCGC_GEMM_TF32X3 ? 3 : (mode == DCGC_GEMM_BF16 ? 1 : 0); }
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
// Explicit per-call options of the tensor-core GEMMs (no thread-local side channels):
//   img      a weight image built ahead with dcgc_tc_prep_weights (e.g. on another stream); null = the call builds
//            its own in a stream-ordered scratch allocation (cudaMallocAsync / cudaFreeAsync on the launch stream)
//   a_exact  every value of the A operand is exactly representable in tf32 (integer-valued atom features and
//            their neighbour sums below 2^11): its lo part is identically zero, so the lo(A) tile and the
//            lo(A) * hi(W) term are skipped — a third of the tensor time, bit-identical results
struct DcgcGemmOpts {
  const float* img = nullptr;
  int a_exact = 0;
};
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
                               int a_exact, void* stream);
#endif

#ifdef __CUDACC__
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

static inline int64_t dcgc_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

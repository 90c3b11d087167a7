// Fused D-MPNN training step (dcgc_dmpnn_model_*): DMPNN.forward (deepchem/models/torch_models/dmpnn.py:246-449 =
// DMPNNEncoderLayer, torch_models/layers.py:1585-1649, + PositionwiseFeedForward, :795-910) + L2 loss
// (models/losses.py:76-94 under _StandardLoss, torch_model.py:1267-1294) + the backward of all of it, as ONE C call
// over flat parameter / gradient slabs — the D-MPNN counterpart of dcgc_gcmodel_train_step.  The per-layer
// autograd path (deepchem_b200/dmpnn.py over the same kernels) needed 2 ms of Python per step to issue 1.6 ms of GPU
// work (scripts/dmpnn_profile.py); this call issues the same kernels from C++ and folds the element-wise steps
// (ReLU, residual add, ReLU masks of the backward) into four small fused kernels.
//
// Covered configuration (everything else stays on the autograd path): ReLU activations, dropout 0, encoder bias
// False (W_i, W_h without bias; W_o with), regression, no global features, hidden widths that are multiples of 4,
// FFN with >= 2 linears.  depth >= 2 as in the reference (depth 1 leaves h_message unbound there).
#include <cuda_runtime.h>

#include <stdlib.h>

#include "common.h"

namespace {

constexpr int kT = 256;

#define RET_IF(expr)                  \
  do {                                \
    int st__ = (expr);                \
    if (st__ != DCGC_OK) return st__; \
  } while (0)

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kT - 1) / kT); }

// all element-wise kernels work on [rows, width] matrices with width % 4 == 0 and 16-byte aligned rows
// h = relu(a + b)                                                   (layers.py:1632: act(input + W_h(message)))
__global__ void __launch_bounds__(kT) add_relu_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                      float4* __restrict__ y, int64_t n4) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n4) return;
  const float4 p = a[i], q = b[i];
  y[i] = make_float4(fmaxf(p.x + q.x, 0.f), fmaxf(p.y + q.y, 0.f), fmaxf(p.z + q.z, 0.f), fmaxf(p.w + q.w, 0.f));
}
// g = (y > 0) ? g : 0      (ReLU backward from the saved OUTPUT, in place)
__global__ void __launch_bounds__(kT) relu_mask_kernel(float4* __restrict__ g, const float4* __restrict__ y, int64_t n4) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n4) return;
  float4 v = g[i];
  const float4 t = y[i];
  v.x = t.x > 0.f ? v.x : 0.f; v.y = t.y > 0.f ? v.y : 0.f; v.z = t.z > 0.f ? v.z : 0.f; v.w = t.w > 0.f ? v.w : 0.f;
  g[i] = v;
}
// CSR gather-sum with the element-wise neighbours of the D-MPNN step folded in (one thread per 16-byte column group
// of one output row, entries summed in index order exactly as dcgc_gather_sum does):
//   RELU_IN  the source rows are read through ReLU        (message = act(input), layers.py:1624, never materialised)
//   MODE 0   out = sum
//   MODE 1   out = (mask > 0) ? sum : 0                   (ReLU backward of h_message on the scattered gradient)
//   MODE 2   out = addend + ((mask > 0) ? sum : 0)        (d input = d s + act'(input) * d message; out may alias addend)
template <bool RELU_IN, int MODE>
__global__ void __launch_bounds__(kT) fused_gather_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_ptr,
                                                          const int32_t* __restrict__ idx, int64_t n_rows, int width,
                                                          const float* __restrict__ mask, const float* addend,
                                                          float* out) {
  dcgc_griddep_wait();
  const int groups = width >> 2;
  const int64_t t = (int64_t)blockIdx.x * kT + threadIdx.x;
  const int64_t row = t / groups;
  if (row >= n_rows) return;
  const int c = (int)(t - row * groups) * 4;
  int e = __ldg(row_ptr + row);
  const int e1 = __ldg(row_ptr + row + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto add = [&](float4 v) {
    if (RELU_IN) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  };
  const float* xc = x + c;
  for (; e + 4 <= e1; e += 4) {      // four independent row loads in flight
    const int j0 = __ldg(idx + e), j1 = __ldg(idx + e + 1), j2 = __ldg(idx + e + 2), j3 = __ldg(idx + e + 3);
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j0 * width));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j1 * width));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j2 * width));
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j3 * width));
    add(v0); add(v1); add(v2); add(v3);
  }
  const int rem = e1 - e;
  if (rem > 0) {
    const int j0 = __ldg(idx + e), j1 = rem > 1 ? __ldg(idx + e + 1) : j0, j2 = rem > 2 ? __ldg(idx + e + 2) : j0;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j0 * width));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j1 * width));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(xc + (int64_t)j2 * width));
    add(v0);
    if (rem > 1) add(v1);
    if (rem > 2) add(v2);
  }
  const int64_t o = row * width + c;
  if (MODE != 0) {
    const float4 m = *reinterpret_cast<const float4*>(mask + o);
    acc.x = m.x > 0.f ? acc.x : 0.f; acc.y = m.y > 0.f ? acc.y : 0.f;
    acc.z = m.z > 0.f ? acc.z : 0.f; acc.w = m.w > 0.f ? acc.w : 0.f;
  }
  if (MODE == 2) {
    const float4 a = *reinterpret_cast<const float4*>(addend + o);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  *reinterpret_cast<float4*>(out + o) = acc;
}

template <bool RELU_IN, int MODE>
int fused_gather(const float* x, const int32_t* row_ptr, const int32_t* idx, int64_t n_rows, int width, const float* mask,
                 const float* addend, float* out, cudaStream_t st) {
  if (n_rows == 0) return DCGC_OK;
  DcgcProfScope prof_scope("dcgc_gather_sum", st);
  dcgc_launch(fused_gather_kernel<RELU_IN, MODE>, blocks_for(n_rows * (width >> 2)), kT, 0, st, x, row_ptr, idx, n_rows, width, mask,
                                                                                      addend, out);
  DCGC_CUDA_LAUNCH_CHECK("dmpnn fused_gather");
  return DCGC_OK;
}

// dst[c, r] = src[r, c]   (nn.Linear weight [n, k] <-> the [k, n] layout of the two-operand GEMM entry points)
__global__ void __launch_bounds__(kT) transpose_kernel(const float* __restrict__ src, int rows, int cols,
                                                       float* __restrict__ dst) {
  dcgc_griddep_wait();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[tx][j];
  }
}
// per-element weighted squared error and its gradient: L = mean(w * (out - y)^2) over all n elements
__global__ void __launch_bounds__(kT) l2_loss_kernel(const float* __restrict__ out, const float* __restrict__ y,
                                                     const float* __restrict__ w, int64_t n, float inv_n,
                                                     float* __restrict__ per_elem, float* __restrict__ dout) {
  dcgc_griddep_wait();
  const int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  if (i >= n) return;
  const float d = out[i] - y[i], ww = w ? w[i] : 1.f;
  per_elem[i] = ww * d * d;
  dout[i] = 2.f * ww * d * inv_n;
}
// fixed-order sum of per_elem (float64), one block: deterministic
__global__ void __launch_bounds__(1024) loss_sum_kernel(const float* __restrict__ per_elem, int64_t n, float inv_n,
                                                        float* __restrict__ loss) {
  dcgc_griddep_wait();
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)per_elem[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(sh[0] * (double)inv_n);
}

inline int64_t pad4(int64_t x) { return (x + 3) / 4 * 4; }
inline int64_t pad64(int64_t x) { return (x + 63) / 64 * 64; }

struct Layout {
  int64_t w_i, w_h, w_o, b_o;
  int64_t ffn_w[DCGC_DMPNN_MAX_FFN], ffn_b[DCGC_DMPNN_MAX_FFN];
  int ffn_in[DCGC_DMPNN_MAX_FFN], ffn_out[DCGC_DMPNN_MAX_FFN];
  int64_t n_params;
};

int check_cfg(const dcgc_dmpnn_model_config* c) {
  DCGC_CHECK_ARG(c, "dcgc_dmpnn_model: null config");
  DCGC_CHECK_ARG(c->atom_fdim > 0 && c->bond_fdim >= 0 && c->hidden > 0 && c->hidden % 4 == 0,
                 "dcgc_dmpnn_model: hidden width must be a positive multiple of 4");
  DCGC_CHECK_ARG(c->depth >= 2, "dcgc_dmpnn_model: depth must be at least 2 (the reference leaves h_message unbound for 1)");
  DCGC_CHECK_ARG(c->ffn_layers >= 2 && c->ffn_layers <= DCGC_DMPNN_MAX_FFN && c->ffn_hidden > 0 &&
                 c->ffn_hidden % 4 == 0 && c->n_out > 0, "dcgc_dmpnn_model: unsupported feed-forward shape");
  DCGC_CHECK_ARG(c->aggregation >= 0 && c->aggregation <= 2, "dcgc_dmpnn_model: Invalid aggregation");
  DCGC_CHECK_ARG(c->gemm_mode == DCGC_GEMM_FP32 || dcgc_tc_terms(c->gemm_mode) != 0, "dcgc_dmpnn_model: bad GEMM mode");
  return DCGC_OK;
}

void make_layout(const dcgc_dmpnn_model_config* c, Layout* lo) {
  const int H = c->hidden, fa = c->atom_fdim, fi = c->atom_fdim + c->bond_fdim;
  int64_t off = 0;
  auto take = [&](int64_t n) { const int64_t o = off; off += pad64(n); return o; };
  lo->w_i = take((int64_t)H * fi);
  lo->w_h = take((int64_t)H * H);
  lo->w_o = take((int64_t)H * (fa + H));
  lo->b_o = take(H);
  for (int i = 0; i < c->ffn_layers; ++i) {
    lo->ffn_in[i] = i == 0 ? H : c->ffn_hidden;
    lo->ffn_out[i] = i == c->ffn_layers - 1 ? c->n_out : c->ffn_hidden;
    lo->ffn_w[i] = take((int64_t)lo->ffn_in[i] * lo->ffn_out[i]);
    lo->ffn_b[i] = take(lo->ffn_out[i]);
  }
  lo->n_params = off;
}

// workspace carve-up (floats unless noted)
struct Work {
  float *r0, *r1, *r2, *r3;          // [n_rows, H]: inp, two message buffers, h
  float *a0, *a1, *a2;               // [n_atoms, H]: m2a, atoms_hidden, gradient scratch
  float *enc, *denc;                 // [n_mols, H]
  float *x[DCGC_DMPNN_MAX_FFN];      // FFN activations x[i] = output of linear i ([n_mols, ffn_out[i] padded])
  float *dx[2];                      // [n_mols, max width] ping-pong gradients
  float *per_elem;                   // [n_mols * n_out]
  float *wo_t, *dwo_t, *dbo;         // W_o^T [fa+H, H], its gradient, bias gradient
  void* wgrad_ws; int64_t wgrad_bytes;
  // split-weight images of the tensor-core GEMMs (null in DCGC_GEMM_FP32 mode), all built by ONE launch at the start
  // of the call instead of one launch in front of every GEMM: W_i, W_h, W_o forward, FFN forward; then the dgrad ones
  float *img_wi, *img_wh, *img_wo, *img_ffn[DCGC_DMPNN_MAX_FFN];
  float *img_ffn_d[DCGC_DMPNN_MAX_FFN], *img_wo_d, *img_wh_d;
  int64_t total_bytes;
};

int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

void carve(const dcgc_dmpnn_model_config* c, const Layout& lo, int64_t n_rows, int64_t n_atoms, int64_t n_mols,
           uint8_t* base, Work* w) {
  const int H = c->hidden;
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += dcgc_align_up(floats * 4, 256);
    return p;
  };
  w->r0 = take(n_rows * H); w->r1 = take(n_rows * H); w->r2 = take(n_rows * H); w->r3 = take(n_rows * H);
  w->a0 = take(n_atoms * H); w->a1 = take(n_atoms * H); w->a2 = take(n_atoms * H);
  w->enc = take(n_mols * H); w->denc = take(n_mols * H);
  int64_t wmax = H;
  for (int i = 0; i < c->ffn_layers; ++i) {
    w->x[i] = take(n_mols * pad4(lo.ffn_out[i]));
    wmax = max64(wmax, pad4(lo.ffn_out[i]));
  }
  w->dx[0] = take(n_mols * wmax); w->dx[1] = take(n_mols * wmax);
  w->per_elem = take(n_mols * c->n_out);
  w->wo_t = take((int64_t)(c->atom_fdim + H) * H);
  w->dwo_t = take((int64_t)(c->atom_fdim + H) * H);
  w->dbo = take(H);
  int64_t wb = dcgc_group_gemm_wgrad_workspace(c->atom_fdim, H, H, 1);
  wb = max64(wb, dcgc_linear_wgrad_workspace(c->atom_fdim + c->bond_fdim, H));
  wb = max64(wb, dcgc_linear_wgrad_workspace(H, H));
  for (int i = 0; i < c->ffn_layers; ++i) wb = max64(wb, dcgc_linear_wgrad_workspace(lo.ffn_in[i], lo.ffn_out[i]));
  w->wgrad_bytes = wb;
  w->wgrad_ws = base ? base + off : nullptr;
  off += dcgc_align_up(wb, 256);
  w->img_wi = w->img_wh = w->img_wo = w->img_wo_d = w->img_wh_d = nullptr;
  for (int i = 0; i < DCGC_DMPNN_MAX_FFN; ++i) w->img_ffn[i] = w->img_ffn_d[i] = nullptr;
  const int nt = dcgc_tc_terms(c->gemm_mode);
  if (nt != 0) {
    const int fa = c->atom_fdim, fi = c->atom_fdim + c->bond_fdim;
    auto take_img = [&](int k1, int k2, int n) {
      return take(max64(dcgc_tc_image_bytes(nt, k1, k2, n, 1), dcgc_tc_image_bytes_f16(k1, k2, n, 1)) / 4);
    };
    w->img_wi = take_img(fi, 0, H);
    w->img_wh = take_img(H, 0, H);
    w->img_wo = take_img(fa, H, H);
    for (int i = 0; i < c->ffn_layers; ++i) w->img_ffn[i] = take_img(lo.ffn_in[i], 0, lo.ffn_out[i]);
    for (int i = 0; i < c->ffn_layers; ++i) w->img_ffn_d[i] = take_img(lo.ffn_out[i], 0, lo.ffn_in[i]);
    w->img_wo_d = take_img(H, 0, fa + H);
    w->img_wh_d = take_img(H, 0, H);
  }
  w->total_bytes = off;
}

// W_o^T (the two-operand GEMM wants [fa + H, H]) and the weight images of the call, in front of its first GEMM.
// Job arguments follow the dcgc_tc_gemm calls behind dcgc_linear_fwd / _dgrad and dcgc_group_gemm_fwd / _dgrad
// (csrc/gemm_simt.cu): forward of nn.Linear weights [n, k] = (trans 0, k, n); dx = g . w = (trans 1, K = n, N = k).
// Forward GEMMs of the TF32x3 mode with fp16 operand halves (DCGC_GEMM_F16X3, csrc/gemm_tc.cu tc_gemm_kernel_v6: half
// the tensor-core instructions for the same 22-bit products; 0/1 features, messages and their sums are far inside
// the range it needs, and the range flag is checked by DMPNNModel).  DCGC_FWD_F16X3=0 keeps tf32 halves.
int forward_mode(const dcgc_dmpnn_model_config* c) {
  static const bool on = [] { const char* e = getenv("DCGC_FWD_F16X3"); return !(e && e[0] == '0'); }();
  return (on && c->gemm_mode == DCGC_GEMM_TF32X3) ? DCGC_GEMM_F16X3 : c->gemm_mode;
}
int build_images(const dcgc_dmpnn_model_config* c, const Layout& lo, const float* params, const Work& w, bool backward,
                 cudaStream_t st) {
  const int H = c->hidden, fa = c->atom_fdim, fi = c->atom_fdim + c->bond_fdim;
  const int f16 = forward_mode(c) == DCGC_GEMM_F16X3 ? 1 : 0;
  {
    dim3 grid((unsigned)((fa + H + 31) / 32), (unsigned)((H + 31) / 32));
    dcgc_launch(transpose_kernel, grid, kT, 0, st, params + lo.w_o, H, fa + H, w.wo_t);
    DCGC_CUDA_LAUNCH_CHECK("dmpnn transpose W_o");
  }
  const int nt = dcgc_tc_terms(c->gemm_mode);
  if (nt == 0) return DCGC_OK;
  DcgcImgJob jobs[5 + 2 * DCGC_DMPNN_MAX_FFN];
  int nj = 0;
  jobs[nj++] = DcgcImgJob{params + lo.w_i, 1, 0, fi, 0, H, w.img_wi, f16};
  jobs[nj++] = DcgcImgJob{params + lo.w_h, 1, 0, H, 0, H, w.img_wh, f16};
  jobs[nj++] = DcgcImgJob{w.wo_t, 1, 1, fa, H, H, w.img_wo, f16};
  for (int i = 0; i < c->ffn_layers; ++i)
    jobs[nj++] = DcgcImgJob{params + lo.ffn_w[i], 1, 0, lo.ffn_in[i], 0, lo.ffn_out[i], w.img_ffn[i], f16};
  if (backward) {
    for (int i = 0; i < c->ffn_layers; ++i)
      jobs[nj++] = DcgcImgJob{params + lo.ffn_w[i], 1, 1, lo.ffn_out[i], 0, lo.ffn_in[i], w.img_ffn_d[i], 0};
    jobs[nj++] = DcgcImgJob{w.wo_t, 1, 0, H, 0, fa + H, w.img_wo_d, 0};
    jobs[nj++] = DcgcImgJob{params + lo.w_h, 1, 1, H, 0, H, w.img_wh_d, 0};
  }
  return dcgc_tc_prep_weights_batch(nt, jobs, nj, st);
}
inline DcgcGemmOpts with_img(const float* img) {
  DcgcGemmOpts o{};
  o.img = img;
  return o;
}

int check_tables(const dcgc_dmpnn_tables* t) {
  DCGC_CHECK_ARG(t && t->n_mols >= 0 && t->n_atoms >= 0 && t->n_rows >= 0, "dcgc_dmpnn_model: bad tables");
  if (t->n_rows > 0 && t->n_atoms > 0)
    DCGC_CHECK_ARG(t->mol_ptr && t->a2b_ptr && t->a2b_idx && t->a2b_t_ptr && t->a2b_t_idx && t->map_ptr && t->map_idx &&
                   t->map_t_ptr && t->map_t_idx, "dcgc_dmpnn_model: null index table");
  return DCGC_OK;
}

// message buffer after `depth - 1` gathers starting from r1: r1 -> r2 -> r1 -> ...
inline float* msg_final(const Work& w, int depth) { return (depth - 1) % 2 == 0 ? w.r1 : w.r2; }
inline float* msg_other(const Work& w, int depth) { return (depth - 1) % 2 == 0 ? w.r2 : w.r1; }

int forward_impl(const dcgc_dmpnn_model_config* c, const Layout& lo, const dcgc_dmpnn_tables* t, const float* af,
                 int64_t ld_af, const float* fini, int64_t ld_fi, const float* params, const Work& w, float* out,
                 int64_t ld_out, cudaStream_t st) {
  const int H = c->hidden, fa = c->atom_fdim, fi = c->atom_fdim + c->bond_fdim, mode = forward_mode(c);
  const int64_t R = t->n_rows, A = t->n_atoms, B = t->n_mols;
  const int64_t r4 = R * H / 4;
  // input = W_i(f_ini)  (layers.py:1622);  message = act(input)  (:1624)
  RET_IF(dcgc_linear_fwd_opts(mode, fini, ld_fi, fi, params + lo.w_i, nullptr, H, R, DCGC_ACT_NONE, w.r0, H, nullptr, nullptr,
                              with_img(w.img_wi), st));
  // message = message[mapping].sum(1), depth - 1 times  (:1627-1629); the first gather reads act(input) on the fly
  float *src = w.r1, *dst = w.r2;
  for (int d = 1; d < c->depth; ++d) {
    if (d == 1)
      RET_IF((fused_gather<true, 0>(w.r0, t->map_ptr, t->map_idx, R, H, nullptr, nullptr, dst, st)));
    else
      RET_IF(dcgc_gather_sum(src, H, t->map_ptr, t->map_idx, R, H, nullptr, 0, dst, H, st));
    float* tmp = src; src = dst; dst = tmp;
  }
  // src == msg_final; h_message = act(input + W_h(message))  (:1630-1633, only the last product is live)
  RET_IF(dcgc_linear_fwd_opts(mode, src, H, H, params + lo.w_h, nullptr, H, R, DCGC_ACT_NONE, dst, H, nullptr, nullptr,
                              with_img(w.img_wh), st));
  if (r4 > 0) {
    dcgc_launch(add_relu_kernel, blocks_for(r4), kT, 0, st, reinterpret_cast<const float4*>(w.r0), reinterpret_cast<const float4*>(dst),
                                                  reinterpret_cast<float4*>(w.r3), r4);
    DCGC_CUDA_LAUNCH_CHECK("dmpnn add_relu");
  }
  // messages to atoms (:1539), atoms_hidden = act(W_o(cat(atom_features, m2a)))  (:1541-1547)
  RET_IF(dcgc_gather_sum(w.r3, H, t->a2b_ptr, t->a2b_idx, A, H, nullptr, 0, w.a0, H, st));
  RET_IF(dcgc_group_gemm_fwd_opts(mode, af, ld_af, fa, w.a0, H, H, w.wo_t, params + lo.b_o, H, nullptr, 0, 128, A,
                                  DCGC_ACT_RELU, w.a1, H, nullptr, nullptr, with_img(w.img_wo), st));
  // readout (:1550-1583)
  RET_IF(dcgc_segment_readout_fwd(w.a1, H, t->mol_ptr, B, H, c->aggregation, c->aggregation_norm, w.enc, H, st));
  // feed-forward (layers.py:880-910 with dropout 0)
  const float* x = w.enc;
  int64_t ld_x = H;
  for (int i = 0; i < c->ffn_layers; ++i) {
    const bool last = i == c->ffn_layers - 1;
    float* y = last && out ? out : w.x[i];
    const int64_t ld_y = last && out ? ld_out : pad4(lo.ffn_out[i]);
    RET_IF(dcgc_linear_fwd_opts(mode, x, ld_x, lo.ffn_in[i], params + lo.ffn_w[i], params + lo.ffn_b[i], lo.ffn_out[i], B,
                                last ? DCGC_ACT_NONE : DCGC_ACT_RELU, y, ld_y, nullptr, nullptr, with_img(w.img_ffn[i]), st));
    x = y; ld_x = ld_y;
  }
  return DCGC_OK;
}

}  // namespace

extern "C" int dcgc_dmpnn_model_layout(const dcgc_dmpnn_model_config* cfg, int64_t* offsets, int64_t* n_params) {
  RET_IF(check_cfg(cfg));
  DCGC_CHECK_ARG(offsets && n_params, "dcgc_dmpnn_model_layout: null output");
  Layout lo;
  make_layout(cfg, &lo);
  offsets[0] = lo.w_i; offsets[1] = lo.w_h; offsets[2] = lo.w_o; offsets[3] = lo.b_o;
  for (int i = 0; i < cfg->ffn_layers; ++i) { offsets[4 + 2 * i] = lo.ffn_w[i]; offsets[5 + 2 * i] = lo.ffn_b[i]; }
  *n_params = lo.n_params;
  return DCGC_OK;
}

extern "C" int64_t dcgc_dmpnn_model_workspace_bytes(const dcgc_dmpnn_model_config* cfg, int64_t n_rows, int64_t n_atoms,
                                                    int64_t n_mols) {
  if (check_cfg(cfg) != DCGC_OK || n_rows < 0 || n_atoms < 0 || n_mols < 0) return -1;
  Layout lo;
  make_layout(cfg, &lo);
  Work w;
  carve(cfg, lo, n_rows, n_atoms, n_mols, nullptr, &w);
  return w.total_bytes;
}

extern "C" int dcgc_dmpnn_model_forward(const dcgc_dmpnn_model_config* cfg, const dcgc_dmpnn_tables* t,
                                        const float* atom_feat, int64_t ld_af, const float* f_ini, int64_t ld_fi,
                                        const float* params, void* workspace, int64_t workspace_bytes, float* out,
                                        float* encoding_out, void* stream) {
  RET_IF(check_cfg(cfg));
  RET_IF(check_tables(t));
  DCGC_CHECK_ARG(ld_af >= cfg->atom_fdim && ld_fi >= cfg->atom_fdim + cfg->bond_fdim, "dcgc_dmpnn_model_forward: bad leading dimensions");
  if (t->n_mols == 0) return DCGC_OK;
  DCGC_CHECK_ARG(atom_feat && f_ini && params && workspace && out, "dcgc_dmpnn_model_forward: null pointer");
  DCGC_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "dcgc_dmpnn_model_forward: workspace must be 256-byte aligned");
  Layout lo;
  make_layout(cfg, &lo);
  Work w;
  carve(cfg, lo, t->n_rows, t->n_atoms, t->n_mols, static_cast<uint8_t*>(workspace), &w);
  DCGC_CHECK_ARG(workspace_bytes >= w.total_bytes, "dcgc_dmpnn_model_forward: workspace too small (%lld < %lld)",
                 (long long)workspace_bytes, (long long)w.total_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  RET_IF(build_images(cfg, lo, params, w, false, st));
  RET_IF(forward_impl(cfg, lo, t, atom_feat, ld_af, f_ini, ld_fi, params, w, out, cfg->n_out, st));
  if (encoding_out)
    DCGC_CUDA_CALL(cudaMemcpyAsync(encoding_out, w.enc, (size_t)t->n_mols * cfg->hidden * 4, cudaMemcpyDeviceToDevice, st));
  return DCGC_OK;
}

extern "C" int dcgc_dmpnn_model_train_step(const dcgc_dmpnn_model_config* cfg, const dcgc_dmpnn_tables* t,
                                           const float* atom_feat, int64_t ld_af, const float* f_ini, int64_t ld_fi,
                                           const float* y, const float* wts, const float* params, float* grads,
                                           void* workspace, int64_t workspace_bytes, float* loss_dev, float* out,
                                           void* stream) {
  RET_IF(check_cfg(cfg));
  RET_IF(check_tables(t));
  DCGC_CHECK_ARG(ld_af >= cfg->atom_fdim && ld_fi >= cfg->atom_fdim + cfg->bond_fdim, "dcgc_dmpnn_model_train_step: bad leading dimensions");
  DCGC_CHECK_ARG(t->n_mols > 0 && t->n_atoms > 0 && t->n_rows > 0, "dcgc_dmpnn_model_train_step: empty batch");
  DCGC_CHECK_ARG(atom_feat && f_ini && y && params && grads && workspace && loss_dev, "dcgc_dmpnn_model_train_step: null pointer");
  DCGC_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "dcgc_dmpnn_model_train_step: workspace must be 256-byte aligned");
  Layout lo;
  make_layout(cfg, &lo);
  Work w;
  carve(cfg, lo, t->n_rows, t->n_atoms, t->n_mols, static_cast<uint8_t*>(workspace), &w);
  DCGC_CHECK_ARG(workspace_bytes >= w.total_bytes, "dcgc_dmpnn_model_train_step: workspace too small (%lld < %lld)",
                 (long long)workspace_bytes, (long long)w.total_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = cfg->hidden, fa = cfg->atom_fdim, fi = cfg->atom_fdim + cfg->bond_fdim, mode = cfg->gemm_mode;
  const int L = cfg->ffn_layers, T = cfg->n_out;
  const int64_t R = t->n_rows, A = t->n_atoms, B = t->n_mols;
  const int64_t a4 = A * H / 4;

  // ---------------- forward (the last linear writes into x[L-1], [B, pad4(T)])
  RET_IF(build_images(cfg, lo, params, w, true, st));
  RET_IF(forward_impl(cfg, lo, t, atom_feat, ld_af, f_ini, ld_fi, params, w, nullptr, 0, st));
  const int64_t ld_o = pad4(T);
  if (out) DCGC_CUDA_CALL(cudaMemcpy2DAsync(out, (size_t)T * 4, w.x[L - 1], (size_t)ld_o * 4, (size_t)T * 4, (size_t)B,
                                             cudaMemcpyDeviceToDevice, st));

  // ---------------- loss: mean over B * T elements of w * (out - y)^2
  float* g = w.dx[0];      // gradient wrt the output of the current linear, [B, pad4(width)]
  float* g_next = w.dx[1];
  {
    // the loss kernels index [B, T] densely: compact copy of the output when T is not a multiple of 4
    const float* o = w.x[L - 1];
    if (ld_o != T) {
      DCGC_CUDA_CALL(cudaMemcpy2DAsync(g_next, (size_t)T * 4, w.x[L - 1], (size_t)ld_o * 4, (size_t)T * 4, (size_t)B,
                                       cudaMemcpyDeviceToDevice, st));
      o = g_next;
    }
    const int64_t n = B * T;
    const float inv_n = 1.0f / (float)n;
    // dout is written densely [B, T] into per_elem's neighbour, then spread to the padded layout if needed
    float* dense_d = ld_o == T ? g : w.denc;      // denc is free until the FFN backward reaches the encoder
    dcgc_launch(l2_loss_kernel, blocks_for(n), kT, 0, st, o, y, wts, n, inv_n, w.per_elem, dense_d);
    DCGC_CUDA_LAUNCH_CHECK("dmpnn l2_loss");
    dcgc_launch(loss_sum_kernel, 1, 1024, 0, st, w.per_elem, n, inv_n, loss_dev);
    DCGC_CUDA_LAUNCH_CHECK("dmpnn loss_sum");
    if (ld_o != T) {
      DCGC_CUDA_CALL(cudaMemsetAsync(g, 0, (size_t)B * ld_o * 4, st));
      DCGC_CUDA_CALL(cudaMemcpy2DAsync(g, (size_t)ld_o * 4, dense_d, (size_t)T * 4, (size_t)T * 4, (size_t)B,
                                       cudaMemcpyDeviceToDevice, st));
    }
  }

  // ---------------- feed-forward backward
  for (int i = L - 1; i >= 0; --i) {
    const int n = lo.ffn_out[i], k = lo.ffn_in[i];
    const int64_t ld_g = pad4(n);
    const float* x_in = i == 0 ? w.enc : w.x[i - 1];
    const int64_t ld_x = i == 0 ? H : pad4(lo.ffn_out[i - 1]);
    if (i != L - 1) {     // ReLU of this linear's output
      const int64_t n4 = B * ld_g / 4;
      dcgc_launch(relu_mask_kernel, blocks_for(n4), kT, 0, st, reinterpret_cast<float4*>(g), reinterpret_cast<const float4*>(w.x[i]), n4);
      DCGC_CUDA_LAUNCH_CHECK("dmpnn relu_mask ffn");
    }
    RET_IF(dcgc_linear_wgrad(mode, x_in, ld_x, k, g, ld_g, n, B, grads + lo.ffn_w[i], grads + lo.ffn_b[i], w.wgrad_ws,
                             w.wgrad_bytes, st));
    float* dxo = i == 0 ? w.denc : g_next;
    const int64_t ld_dx = i == 0 ? H : pad4(k);
    if (i != 0 && ld_dx != k) DCGC_CUDA_CALL(cudaMemsetAsync(dxo, 0, (size_t)B * ld_dx * 4, st));
    RET_IF(dcgc_linear_dgrad_opts(mode, g, ld_g, n, params + lo.ffn_w[i], k, B, dxo, ld_dx, with_img(w.img_ffn_d[i]), st));
    if (i != 0) { float* tmp = g; g = g_next; g_next = tmp; }
  }

  // ---------------- readout backward, ReLU of atoms_hidden, W_o
  RET_IF(dcgc_segment_readout_bwd(w.denc, H, t->mol_ptr, B, A, H, cfg->aggregation, cfg->aggregation_norm, w.a2, H, st));
  dcgc_launch(relu_mask_kernel, blocks_for(a4), kT, 0, st, reinterpret_cast<float4*>(w.a2), reinterpret_cast<const float4*>(w.a1), a4);
  DCGC_CUDA_LAUNCH_CHECK("dmpnn relu_mask atoms");
  {
    const int64_t rows[1] = {A};
    RET_IF(dcgc_group_gemm_wgrad(mode, atom_feat, ld_af, fa, w.a0, H, H, w.a2, H, H, rows, 1, w.dwo_t, grads + lo.b_o,
                                 w.wgrad_ws, w.wgrad_bytes, st));
    dim3 grid((unsigned)((H + 31) / 32), (unsigned)((fa + H + 31) / 32));
    dcgc_launch(transpose_kernel, grid, kT, 0, st, w.dwo_t, fa + H, H, grads + lo.w_o);
    DCGC_CUDA_LAUNCH_CHECK("dmpnn transpose dW_o");
  }
  // d(m2a) = g . W_o[:, fa:]  -> a1 (atoms_hidden is not needed any more); atom features need no gradient
  RET_IF(dcgc_group_gemm_dgrad_opts(mode, w.a2, H, H, w.wo_t, fa, H, nullptr, 0, 128, A, nullptr, 0, w.a1, H,
                                    with_img(w.img_wo_d), st));

  // ---------------- bonds: transposed gathers, ReLU masks, W_h, W_i
  float* mf = msg_final(w, cfg->depth);   // message after the gathers (input of W_h)
  float* mo = msg_other(w, cfg->depth);   // held W_h(message); free now
  // d(h_message) = scatter of d(m2a) over a2b  -> mo;  d(pre-activation) = mask by h > 0
  RET_IF((fused_gather<false, 1>(w.a1, t->a2b_t_ptr, t->a2b_t_idx, R, H, w.r3, nullptr, mo, st)));
  RET_IF(dcgc_linear_wgrad(mode, mf, H, H, mo, H, H, R, grads + lo.w_h, nullptr, w.wgrad_ws, w.wgrad_bytes, st));
  // d(message) = ds . W_h  -> r3 (h is not needed any more), then depth - 1 transposed gathers r3 -> mf -> r3 ...
  RET_IF(dcgc_linear_dgrad_opts(mode, mo, H, H, params + lo.w_h, H, R, w.r3, H, with_img(w.img_wh_d), st));
  float *src = w.r3, *dst = mf;
  for (int d = 1; d < cfg->depth; ++d) {
    if (d == cfg->depth - 1)   // the last one lands on message_0 = act(input): d(input) = ds + (input > 0) * it, in place in mo
      RET_IF((fused_gather<false, 2>(src, t->map_t_ptr, t->map_t_idx, R, H, w.r0, mo, mo, st)));
    else
      RET_IF(dcgc_gather_sum(src, H, t->map_t_ptr, t->map_t_idx, R, H, nullptr, 0, dst, H, st));
    float* tmp = src; src = dst; dst = tmp;
  }
  RET_IF(dcgc_linear_wgrad(mode, f_ini, ld_fi, fi, mo, H, H, R, grads + lo.w_i, nullptr, w.wgrad_ws, w.wgrad_bytes, st));
  return DCGC_OK;
}

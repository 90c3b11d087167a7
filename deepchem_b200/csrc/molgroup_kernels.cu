// Molecule-group staged kernels for sm_100a: neighbour gather-sum (K1/K5), GraphPool forward (K3) and
// backward (K7) with the atom rows of a group of molecules staged in shared memory by bulk (TMA) copies.
//
// Why: the one-thread-per-16-bytes gathers of graph_kernels.cu read every activation row once per reference
// (1 + mean degree = 3.2 times for the pool) through dependent index -> row loads; ncu shows them bound by
// bytes in flight, DRAM 35-55 % busy (profiles/r1p_ncu_gather_sum_bucketed.md).  In the ConvMol layout
// (deepchem/feat/mol_graphs.py:256-349) rows are ordered (degree, molecule, position), so the rows of a
// range of consecutive molecules are ONE contiguous range per degree bucket, and no edge leaves a molecule.
// The host layout builder cuts the batch into molecule groups of at most R rows (dcgc.h).
//
//   dcgc_mg_prepare (once per batch, after the slab upload): one 32-byte record per row for the forward lists
//     and one 48-byte record for the transposed lists: {global row, degree, the SHARED-MEMORY SLOT of each
//     neighbour inside the row's group (u16 x 10) [, the slot of this row in the neighbour's list (u8 x 10)]}.
//     Records are stored in row order, so the records of a group are the same contiguous ranges as its rows.
//   mg_kernel: a persistent CTA per SM walks groups g = blockIdx.x, blockIdx.x + gridDim.x, ... through an
//     S-stage shared-memory ring.  Producer warp s owns stage s: it waits until the consumers released the
//     stage, reads the group's table row and issues one cp.async.bulk per non-empty degree bucket and staged
//     tensor (activation rows, records, argmax bytes) completing on the stage's `full` mbarrier.  Every row is
//     read from HBM / L2 exactly once, as a contiguous burst, with S groups in flight per SM.  16 consumer
//     warps wait on `full`; each thread keeps four (row, 16-byte column group) items in flight: one 16-byte
//     record load gives the row id, the degree and five neighbour slots, so all row loads of an item are
//     independent shared-memory reads (consecutive lanes = consecutive columns: conflict free); each output
//     row is written once with coalesced 128-bit stores; the stage is released through its `empty` mbarrier.
// No float atomics; summation / comparison order = index order, identical to graph_kernels.cu (bit-identical
// results, tests/test_gpu_staged.py).
#include "common.h"

namespace {

constexpr int kConsumerWarps = 16;
constexpr int kConsumers = 32 * kConsumerWarps;
constexpr int kMaxStages = 4;
constexpr int kMaxSmem = 220 * 1024;
constexpr int kItems = 4;             // items in flight per consumer thread
constexpr int kRecF = 32, kRecB = 48; // record bytes (forward / transposed lists)

struct MgBuckets {
  int row0[DCGC_N_DEG + 1];   // first row of bucket d; row0[11] = number of rows
  int e0[DCGC_N_DEG + 1];     // first entry of bucket d
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { dcgc_mbar_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// debugging aid (dcgcdbg_mg_timeline): CTA 0 stores clock64() at dbg[8 * it + kind] (plain stores: no round trip)
__device__ __forceinline__ void mg_mark(long long* dbg, int kind, int it) {
  if (dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it < 500) dbg[8 * it + kind] = clock64();
}

__host__ __device__ inline int64_t mg_align(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// prepare: per-row records
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mg_prepare_kernel(const MgBuckets bk, const int32_t* __restrict__ groups, int n_groups,
                  const int32_t* __restrict__ col_idx, const int32_t* __restrict__ t_src,
                  const int32_t* __restrict__ t_slot, int n_rows, uint4* __restrict__ rec_f, uint4* __restrict__ rec_b) {
  dcgc_griddep_wait();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n_rows) return;
  int d = 0, r0 = bk.row0[0], e0 = bk.e0[0];
#pragma unroll
  for (int k = 1; k < DCGC_N_DEG; ++k) {
    const bool ge = r >= bk.row0[k];
    d = ge ? k : d; r0 = ge ? bk.row0[k] : r0; e0 = ge ? bk.e0[k] : e0;
  }
  // group: tab[g][d] <= r < tab[g+1][d]  (first g' with tab[g'][d] > r, minus one)
  int lo_g = 0, hi_g = n_groups;   // the closing row n_groups holds the bucket end > r
  while (lo_g < hi_g) {
    const int mid = (lo_g + hi_g) >> 1;
    if (__ldg(groups + (int64_t)mid * DCGC_GROUP_STRIDE + d) > r) hi_g = mid; else lo_g = mid + 1;
  }
  const int g = lo_g - 1;
  int lo[DCGC_N_DEG], base[DCGC_N_DEG], cnt[DCGC_N_DEG];
  int acc = 0;
#pragma unroll
  for (int q = 0; q < DCGC_N_DEG; ++q) {
    lo[q] = __ldg(groups + (int64_t)g * DCGC_GROUP_STRIDE + q);
    cnt[q] = __ldg(groups + (int64_t)(g + 1) * DCGC_GROUP_STRIDE + q) - lo[q];
    base[q] = acc;
    acc += cnt[q];
  }
  auto slot_of = [&](int j) {
    int sl = 0;   // a neighbour outside the group cannot happen (validated by the layout builder)
#pragma unroll
    for (int q = 0; q < DCGC_N_DEG; ++q) {
      const int off = j - lo[q];
      if (off >= 0 && off < cnt[q]) sl = base[q] + off;
    }
    return (uint32_t)sl;
  };
  const int e = e0 + (r - r0) * d;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int32_t* idx = pass == 0 ? col_idx : t_src;
    if (pass == 1 && !rec_b) continue;
    uint32_t sl[10], ts[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      sl[k] = k < d ? slot_of(__ldg(idx + e + k)) : 0u;
      ts[k] = (pass == 1 && k < d) ? (uint32_t)__ldg(t_slot + e + k) : 0u;
    }
    const uint4 a = make_uint4((uint32_t)r, (uint32_t)d | (sl[0] << 16), sl[1] | (sl[2] << 16), sl[3] | (sl[4] << 16));
    const uint4 b = make_uint4(sl[5] | (sl[6] << 16), sl[7] | (sl[8] << 16), sl[9], 0u);
    if (pass == 0) {
      rec_f[2 * (int64_t)r] = a;
      rec_f[2 * (int64_t)r + 1] = b;
    } else {
      rec_b[3 * (int64_t)r] = a;
      rec_b[3 * (int64_t)r + 1] = b;
      rec_b[3 * (int64_t)r + 2] = make_uint4(ts[0] | (ts[1] << 8) | (ts[2] << 16) | (ts[3] << 24),
                                             ts[4] | (ts[5] << 8) | (ts[6] << 16) | (ts[7] << 24), ts[8] | (ts[9] << 8), 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// staged kernel
// ------------------------------------------------------------------------------------------
// Shared memory: header (barriers) | S stages; one stage: n_rows (128 bytes) | up to three staged tensors, each
// 128-byte aligned: 0 = records, 1 = activation rows, 2 = argmax bytes (pool backward)
struct MgGeom {
  int rb[3];        // row bytes of the staged tensors (0: unused)
  int off[3];
  int stage_bytes, n_stages, max_rows;
};
constexpr int kHeaderBytes = 128;   // full[kMaxStages], empty[kMaxStages] (8 bytes each)
static_assert(16 * kMaxStages <= kHeaderBytes, "header too small");

inline MgGeom mg_geom(int max_rows, int rb_rec, int rb_rows, int rb_arg) {
  MgGeom g;
  g.max_rows = max_rows;
  g.rb[0] = rb_rec; g.rb[1] = rb_rows; g.rb[2] = rb_arg;
  int off = 128;
  for (int t = 0; t < 3; ++t) {
    g.off[t] = off;
    off += (int)mg_align((int64_t)max_rows * g.rb[t], 128);
  }
  g.stage_bytes = off;
  const int s = (kMaxSmem - kHeaderBytes) / g.stage_bytes;
  g.n_stages = s > kMaxStages ? kMaxStages : s;
  return g;
}

struct MgSrc { const char* p[3]; };

// Producer warp of stage s: fills the stage for groups it = s, s + S, ... of this CTA's sequence.
__device__ __forceinline__ void mg_producer(char* smem, const MgGeom& geo, int s, const int32_t* __restrict__ groups,
                                            int n_groups, const MgSrc& src, long long* dbg, int dbg_mode) {
  const int lane = threadIdx.x & 31;
  const uint32_t full = smem_u32(smem + 8 * s), empty = smem_u32(smem + 8 * (kMaxStages + s));
  char* stage = smem + kHeaderBytes + (size_t)s * geo.stage_bytes;
  const int S = geo.n_stages;
  const int rb_sum = geo.rb[0] + geo.rb[1] + geo.rb[2];
  int use = 0;   // how many times this stage has been filled
  for (int it = s; ; it += S, ++use) {
    const int g = blockIdx.x + it * gridDim.x;
    if (g >= n_groups) break;
    int lo = 0, hi = 0;   // the table row does not depend on the stage: load it before waiting
    if (lane < DCGC_N_DEG) {
      lo = __ldg(groups + (int64_t)g * DCGC_GROUP_STRIDE + lane);
      hi = __ldg(groups + (int64_t)(g + 1) * DCGC_GROUP_STRIDE + lane);
    }
    if (use > 0) mbar_wait(empty, (use - 1) & 1);
    mg_mark(dbg, 1, it);
    const int cnt = hi - lo;
    int base = cnt;   // inclusive prefix sum over the buckets
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, base, o);
      if (lane >= o) base += t;
    }
    const int n_rows = __shfl_sync(0xffffffffu, base, DCGC_N_DEG);
    base -= cnt;
    if (lane == 0) {
      *reinterpret_cast<int*>(stage) = n_rows;
      mbar_expect_tx(full, (uint32_t)n_rows * (uint32_t)(dbg_mode == 3 ? geo.rb[0] : rb_sum));
    }
    __syncwarp();
    if (lane < DCGC_N_DEG && cnt > 0) {
#pragma unroll
      for (int t = 0; t < 3; ++t)
        if (geo.rb[t] && (dbg_mode != 3 || t == 0))
          bulk_g2s(smem_u32(stage + geo.off[t]) + (uint32_t)base * geo.rb[t], src.p[t] + (int64_t)lo * geo.rb[t],
                   (uint32_t)cnt * geo.rb[t], full);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(full);   // release: n_rows is visible to whoever observes the phase
    mg_mark(dbg, 2, it);
  }
}

__device__ __forceinline__ void vadd(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// ---- consumers ---------------------------------------------------------------------------------------------------
// The consumer path is issue-bound if written carelessly (ncu r2b: 500 warp instructions per warp and group, 50 %
// issue slots busy, 2400 cycles per group).  Hence: the column group of a thread is FIXED for the whole kernel
// (thread t < used: cg = t % cgroups, rows i0 + n * istride), all per-thread addresses are computed once, shared
// memory is addressed with 32-bit .shared addresses, the stage / phase are tracked incrementally (no division),
// and the five neighbour slots that arrive with the record's first 16 bytes are handled by straight-line
// predicated code (a record's unused slots are 0: always a valid staged row).
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// what a consumer thread knows about the current stage (32-bit shared addresses, already offset by its column group)
struct MgView {
  uint32_t recs;       // records of the stage
  uint32_t rows_cg;    // rows + 16 * cg
  uint32_t args_cg;    // argmax bytes + 4 * cg
  int rb_rec, rb_rows, rb_arg;
};
// one (row slot, column group) item in flight: first 16 bytes of its record
struct MgItem {
  uint32_t rec;        // shared address of the record
  int slot;            // row slot i
  int row, d;
  uint32_t w1, w2, w3; // degree | slot0, slot1 | slot2, slot3 | slot4
};
template <int K>
__device__ __forceinline__ int mg_slot5(const MgItem& m) {
  return K == 0 ? (int)(m.w1 >> 16) : K == 1 ? (int)(m.w2 & 0xffffu) : K == 2 ? (int)(m.w2 >> 16)
       : K == 3 ? (int)(m.w3 & 0xffffu) : (int)(m.w3 >> 16);
}

struct GatherSumOp {
  const float* addend; float* out; int ld_add, ld_out;              // addend / out already offset by 4 * cg
  struct St { float4 acc; };
  __device__ __forceinline__ void bind(int cg) {
    if (addend) addend += 4 * cg;
    out += 4 * cg;
  }
  __device__ __forceinline__ void init(St& s, const MgItem& m, const MgView&, bool valid) const {
    s.acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (addend && valid) s.acc = *reinterpret_cast<const float4*>(addend + (int64_t)m.row * ld_add);   // may alias out
  }
  __device__ __forceinline__ void step(St& s, const MgItem&, const MgView& v, int, int sl, bool on) const {
    if (on) {
      const float4 q = lds128f(v.rows_cg + sl * v.rb_rows);
      s.acc.x += q.x; s.acc.y += q.y; s.acc.z += q.z; s.acc.w += q.w;
    }
  }
  __device__ __forceinline__ void pre(St&, const MgItem&, bool) const {}
  __device__ __forceinline__ void fin(const St& s, const MgItem& m, int) {
    *reinterpret_cast<float4*>(out + (int64_t)m.row * ld_out) = s.acc;
  }
  __device__ __forceinline__ void finish(char*, int, int, int) const {}
};

// GraphPool forward with the folded BatchNorm affine; first slot attaining the max wins (strict >)
template <bool AFFINE, bool ARG>
struct PoolFwdOp {
  float4 sc, sh;                                                     // this thread's column group
  float* out; uint8_t* arg; int ld_out, ld_arg;                      // already offset by the column group
  struct St { float4 m; uint32_t a; };
  const float* scale; const float* shift;
  __device__ __forceinline__ void bind(int cg) {
    if (AFFINE) {
      sc = __ldg(reinterpret_cast<const float4*>(scale) + cg);
      sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
    }
    out += 4 * cg;
    if (ARG) arg += 4 * cg;
  }
  __device__ __forceinline__ float4 aff(float4 v) const {
    if (AFFINE) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w); }
    return v;
  }
  __device__ __forceinline__ void init(St& s, const MgItem& m, const MgView& v, bool) const {
    s.m = aff(lds128f(v.rows_cg + m.slot * v.rb_rows));
    s.a = 0u;
  }
  __device__ __forceinline__ void step(St& s, const MgItem&, const MgView& v, int k, int sl, bool on) const {
    if (on) {
      const float4 u = aff(lds128f(v.rows_cg + sl * v.rb_rows));
      const uint32_t c = (uint32_t)(k + 1);
      if (u.x > s.m.x) { s.m.x = u.x; s.a = (s.a & 0xffffff00u) | c; }
      if (u.y > s.m.y) { s.m.y = u.y; s.a = (s.a & 0xffff00ffu) | (c << 8); }
      if (u.z > s.m.z) { s.m.z = u.z; s.a = (s.a & 0xff00ffffu) | (c << 16); }
      if (u.w > s.m.w) { s.m.w = u.w; s.a = (s.a & 0x00ffffffu) | (c << 24); }
    }
  }
  __device__ __forceinline__ void pre(St&, const MgItem&, bool) const {}
  __device__ __forceinline__ void fin(const St& s, const MgItem& m, int) {
    *reinterpret_cast<float4*>(out + (int64_t)m.row * ld_out) = s.m;
    if (ARG) *reinterpret_cast<uint32_t*>(arg + (int64_t)m.row * ld_arg) = s.a;
  }
  __device__ __forceinline__ void finish(char*, int, int, int) const {}
};

// GraphPool backward over the transposed lists of a symmetric adjacency (dy rows and arg rows staged).
// STATS: the kernel also produces the per-column sums the BatchNorm backward needs from the rows it writes,
//   sum_r dx[r,c]   and   sum_r dx[r,c] * y[r,c]      (y = the BatchNorm INPUT of the layer, read once, coalesced),
// instead of a separate pass over dx and y (col_moments_partial: 38 us per layer in situ).  The column group of a
// thread is fixed, so the sums live in 8 registers: fp32 over the ~N / (148 * 16) rows of the thread, centred on the
// batch mean (sum dx * (y - mean): no cancellation later), float64 from the cross-thread reduction on; one
// [2][width] row of partials per CTA, reduced by bn_bwd_finalize in CTA order (deterministic).
template <bool AFFINE, bool STATS>
struct PoolBwdOp {
  float4 sc;                                                         // this thread's column group
  float* dx; int ld_dx;                                              // already offset by the column group
  struct St { float4 acc; uint32_t t0, t1; float4 yv; };             // t0, t1: slots of this row in its neighbours' lists
  const float* scale;
  const float* y; int ld_y; const float* mean; double* part; int width;   // STATS only
  DcgcBnFin bnfin;                                                   // STATS only: finalize by the last CTA (kind 0 = off)
  float4 mu, sa, sc2;                                                // column means, sum dx, sum dx * (y - mean)
  __device__ __forceinline__ void bind(int cg) {
    if (AFFINE) sc = __ldg(reinterpret_cast<const float4*>(scale) + cg);
    dx += 4 * cg;
    if (STATS) {
      y += 4 * cg;
      mu = __ldg(reinterpret_cast<const float4*>(mean) + cg);
      sa = make_float4(0.f, 0.f, 0.f, 0.f);
      sc2 = sa;
    }
  }
  static __device__ __forceinline__ float4 pick(const float4 q, uint32_t b, uint32_t c) {
    float4 r;
    r.x = (b & 0xffu) == c ? q.x : 0.f; r.y = ((b >> 8) & 0xffu) == c ? q.y : 0.f;
    r.z = ((b >> 16) & 0xffu) == c ? q.z : 0.f; r.w = (b >> 24) == c ? q.w : 0.f;
    return r;
  }
  __device__ __forceinline__ void init(St& s, const MgItem& m, const MgView& v, bool) const {
    const uint2 t = lds64(m.rec + 32);
    s.t0 = t.x; s.t1 = t.y;
    s.acc = pick(lds128f(v.rows_cg + m.slot * v.rb_rows), lds32(v.args_cg + m.slot * v.rb_arg), 0u);
  }
  __device__ __forceinline__ void step(St& s, const MgItem& m, const MgView& v, int k, int sl, bool on) const {
    if (on) {
      const uint32_t c = (k < 4 ? (s.t0 >> (8 * k)) & 0xffu : k < 8 ? (s.t1 >> (8 * (k - 4))) & 0xffu : lds8(m.rec + 32 + k)) + 1u;
      const float4 r = pick(lds128f(v.rows_cg + sl * v.rb_rows), lds32(v.args_cg + sl * v.rb_arg), c);
      s.acc.x += r.x; s.acc.y += r.y; s.acc.z += r.z; s.acc.w += r.w;
    }
  }
  // all y loads of the items in flight are issued before the first one is used
  __device__ __forceinline__ void pre(St& s, const MgItem& m, bool valid) const {
    if (STATS) s.yv = valid ? __ldg(reinterpret_cast<const float4*>(y + (int64_t)m.row * ld_y)) : mu;
  }
  __device__ __forceinline__ void fin(const St& s, const MgItem& m, int) {
    float4 acc = s.acc;
    if (AFFINE) { acc.x *= sc.x; acc.y *= sc.y; acc.z *= sc.z; acc.w *= sc.w; }
    *reinterpret_cast<float4*>(dx + (int64_t)m.row * ld_dx) = acc;
    if (STATS) {
      sa.x += acc.x; sa.y += acc.y; sa.z += acc.z; sa.w += acc.w;
      sc2.x = fmaf(acc.x, s.yv.x - mu.x, sc2.x); sc2.y = fmaf(acc.y, s.yv.y - mu.y, sc2.y);
      sc2.z = fmaf(acc.z, s.yv.z - mu.z, sc2.z); sc2.w = fmaf(acc.w, s.yv.w - mu.w, sc2.w);
    }
  }
  // CTA-level reduction (consumer threads only; the producers have left): thread (cg, row lane) -> shared memory,
  // then one thread per (quantity, column) adds the row lanes in lane order in float64 and writes this CTA's row of
  // partials: part[cta][0][c] = sum dx, part[cta][1][c] = sum dx * y  ( = centred sum + mean * sum dx ).
  __device__ __forceinline__ void finish(char* scratch, int ct, int cgroups, int used) const {
    if (!STATS) return;
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");      // every consumer is done with the stages
    float* sh = reinterpret_cast<float*>(scratch);                     // [row lane][2][width]
    const int cg = ct % cgroups, lane_r = ct / cgroups;
    if (ct < used) {
      float* q = sh + (size_t)lane_r * 2 * width + 4 * cg;
      *reinterpret_cast<float4*>(q) = sa;
      *reinterpret_cast<float4*>(q + width) = sc2;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    const int lanes = used / cgroups;
    for (int c = ct; c < width; c += kConsumers) {
      double a = 0.0, b = 0.0;
      for (int l = 0; l < lanes; ++l) {
        a += (double)sh[(size_t)l * 2 * width + c];
        b += (double)sh[(size_t)l * 2 * width + width + c];
      }
      part[((int64_t)blockIdx.x * 2) * width + c] = a;
      part[((int64_t)blockIdx.x * 2 + 1) * width + c] = b + (double)__ldg(mean + c) * a;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");      // the float staging above is no longer read
    dcgc_bn_fin_last_cta(bnfin, (int)gridDim.x, gridDim.x, 1, kConsumers, ct, reinterpret_cast<double*>(scratch));
  }
};

template <class Op>
__global__ void __launch_bounds__(32 * (kMaxStages + kConsumerWarps), 1)
mg_kernel(const MgGeom geo, const int32_t* __restrict__ groups, int n_groups, const MgSrc src, int cgroups,
          const Op op_in, long long* dbg, int dbg_mode) {
  extern __shared__ __align__(128) char smem[];
  const int S = geo.n_stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(smem + 8 * s), 1);                                   // full: the producer's arrive (+ tx bytes)
      mbar_init(smem_u32(smem + 8 * (kMaxStages + s)), kConsumerWarps);       // empty: one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  dcgc_griddep_wait();     // (everything above is set-up in shared / tensor memory: it overlaps the kernel in front)
  const int warp = threadIdx.x >> 5;
  if (warp >= kConsumerWarps) {
    const int s = warp - kConsumerWarps;
    if (s < S) mg_producer(smem, geo, s, groups, n_groups, src, dbg, dbg_mode);
    return;
  }
  // thread t < used owns column group t % cgroups of rows t / cgroups + n * istride
  const int ct = threadIdx.x;
  const int istride = kConsumers / cgroups;
  const int used = istride * cgroups;
  const int cg = ct % cgroups;
  const int i0 = ct < used ? ct / cgroups : (1 << 30);
  Op op = op_in;
  op.bind(cg);
  MgView v;
  v.rb_rec = geo.rb[0]; v.rb_rows = geo.rb[1]; v.rb_arg = geo.rb[2];
  const uint32_t smem0 = smem_u32(smem);
  const bool lane0 = (threadIdx.x & 31) == 0;
  int s = 0;
  uint32_t phase = 0;
  uint32_t stage = smem0 + kHeaderBytes;
  for (int it = 0, g = blockIdx.x; g < n_groups; ++it, g += gridDim.x) {
    v.recs = stage + geo.off[0];
    v.rows_cg = stage + geo.off[1] + 16 * cg;
    v.args_cg = stage + geo.off[2] + 4 * cg;
    if (warp == 0) mg_mark(dbg, 4, it);
    mbar_wait(smem0 + 8 * s, phase);
    if (warp == 0) mg_mark(dbg, 5, it);
    const int n_rows = (int)lds32(stage);
    for (int ib = i0; ib < n_rows; ib += kItems * istride) {
      MgItem m[kItems];
      typename Op::St st[kItems];
      bool valid[kItems];
      int dmax = 0;
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        m[u].slot = ib + u * istride;
        valid[u] = m[u].slot < n_rows;
        m[u].rec = v.recs + (valid[u] ? m[u].slot : 0) * v.rb_rec;
        m[u].slot = valid[u] ? m[u].slot : 0;
      }
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        const uint4 w = lds128(m[u].rec);
        m[u].row = (int)w.x; m[u].d = valid[u] ? (int)(w.y & 0xffffu) : 0; m[u].w1 = w.y; m[u].w2 = w.z; m[u].w3 = w.w;
        dmax = max(dmax, m[u].d);
      }
#pragma unroll
      for (int u = 0; u < kItems; ++u) op.init(st[u], m[u], v, valid[u]);
      if (dbg_mode != 2) {
#pragma unroll
        for (int u = 0; u < kItems; ++u) op.step(st[u], m[u], v, 0, mg_slot5<0>(m[u]), 0 < m[u].d);
#pragma unroll
        for (int u = 0; u < kItems; ++u) op.step(st[u], m[u], v, 1, mg_slot5<1>(m[u]), 1 < m[u].d);
        if (dmax > 2) {
#pragma unroll
          for (int u = 0; u < kItems; ++u) op.step(st[u], m[u], v, 2, mg_slot5<2>(m[u]), 2 < m[u].d);
#pragma unroll
          for (int u = 0; u < kItems; ++u) op.step(st[u], m[u], v, 3, mg_slot5<3>(m[u]), 3 < m[u].d);
          if (dmax > 4) {
#pragma unroll
            for (int u = 0; u < kItems; ++u) op.step(st[u], m[u], v, 4, mg_slot5<4>(m[u]), 4 < m[u].d);
            for (int k = 5; k < dmax; ++k) {
#pragma unroll
              for (int u = 0; u < kItems; ++u)
                op.step(st[u], m[u], v, k, (int)lds16(m[u].rec + 6 + 2 * k), k < m[u].d);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kItems; ++u) op.pre(st[u], m[u], valid[u]);
#pragma unroll
      for (int u = 0; u < kItems; ++u)
        if (valid[u] && (dbg_mode != 1 || m[u].d == 77)) op.fin(st[u], m[u], cg);
    }
    if (warp == 0) mg_mark(dbg, 6, it);
    __syncwarp();
    if (lane0) mbar_arrive(smem0 + 8 * (kMaxStages + s));
    stage += geo.stage_bytes;
    if (++s == S) { s = 0; phase ^= 1u; stage = smem0 + kHeaderBytes; }
  }
  op.finish(smem + kHeaderBytes, ct, cgroups, used);
}

long long* g_mg_timeline = nullptr;   // debugging aid, see dcgcdbg_mg_timeline
int g_mg_dbg_mode = 0;                // knock-outs for bound analysis: 1 no stores, 2 no row reads, 3 no row copies

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int fill_buckets(const dcgc_topology* t, MgBuckets* bk) {
  int64_t r = 0, e = 0;
  for (int d = 0; d < DCGC_N_DEG; ++d) {
    DCGC_CHECK_ARG(t->deg_count[d] >= 0, "dcgc_mg: negative bucket size");
    bk->row0[d] = (int)r;
    bk->e0[d] = (int)e;
    r += t->deg_count[d];
    e += (int64_t)d * t->deg_count[d];
  }
  bk->row0[DCGC_N_DEG] = (int)r;
  bk->e0[DCGC_N_DEG] = (int)e;
  DCGC_CHECK_ARG(r == t->n_atoms && e == t->n_edges, "dcgc_mg: bucket sizes do not add up to the topology's rows / entries");
  return DCGC_OK;
}

template <class Op>
int mg_launch(const dcgc_topology* t, const MgGeom& geo, const void* recs, const void* rows, const void* args,
              int cgroups, const Op& op, cudaStream_t st, const char* what) {
  static thread_local bool attr_done = false;   // per instantiation
  if (!attr_done) {
    DCGC_CUDA_CALL(cudaFuncSetAttribute(mg_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_done = true;
  }
  const int sms = dcgc_tc_num_sms();
  const int grid = t->n_groups < sms ? t->n_groups : sms;
  const size_t smem = (size_t)kHeaderBytes + (size_t)geo.n_stages * geo.stage_bytes;
  MgSrc src;
  src.p[0] = reinterpret_cast<const char*>(recs);
  src.p[1] = reinterpret_cast<const char*>(rows);
  src.p[2] = reinterpret_cast<const char*>(args);
  dcgc_launch(mg_kernel<Op>, grid, 32 * (kMaxStages + kConsumerWarps), smem, st, geo, t->groups, t->n_groups, src, cgroups, op,
                                                                        g_mg_timeline, g_mg_dbg_mode);
  DCGC_CUDA_LAUNCH_CHECK(what);
  return DCGC_OK;
}

const char* mg_rec(const dcgc_topology* t, int transposed) {
  return reinterpret_cast<const char*>(t->mg_records) + (transposed ? t->n_atoms * kRecF : 0);
}

}  // namespace

// debugging aid (scripts/mg_timeline.py): CTA 0 of the following staged launches stores clock64() marks
extern "C" void dcgcdbg_mg_timeline(long long* dev_buf) { g_mg_timeline = dev_buf; }
extern "C" void dcgcdbg_mg_mode(int mode) { g_mg_dbg_mode = mode; }

extern "C" int64_t dcgc_mg_record_bytes(int64_t n_atoms) { return n_atoms < 0 ? -1 : n_atoms * (kRecF + kRecB); }

extern "C" int dcgc_mg_prepare(const dcgc_topology* t, void* records, void* stream) {
  DCGC_CHECK_ARG(t, "dcgc_mg_prepare: null topology");
  if (t->n_atoms == 0 || t->n_groups <= 0) return DCGC_OK;
  DCGC_CHECK_ARG(records && al16(records) && t->groups && t->col_idx && t->t_src && t->t_slot,
                 "dcgc_mg_prepare: null or misaligned pointer");
  DCGC_CHECK_ARG(t->group_max_rows > 0 && t->group_max_rows <= 65535, "dcgc_mg_prepare: group of %d rows", t->group_max_rows);
  MgBuckets bk;
  const int st0 = fill_buckets(t, &bk);
  if (st0 != DCGC_OK) return st0;
  uint4* rec_f = reinterpret_cast<uint4*>(records);
  uint4* rec_b = t->symmetric ? reinterpret_cast<uint4*>(reinterpret_cast<char*>(records) + t->n_atoms * kRecF) : nullptr;
  dcgc_launch(mg_prepare_kernel, (unsigned)((t->n_atoms + 255) / 256), 256, 0, (cudaStream_t)stream, 
      bk, t->groups, t->n_groups, t->col_idx, t->t_src, t->t_slot, (int)t->n_atoms, rec_f, rec_b);
  DCGC_CUDA_LAUNCH_CHECK("dcgc_mg_prepare");
  return DCGC_OK;
}

extern "C" int dcgc_mg_supported(const dcgc_topology* t, int64_t ld_floats, int64_t ld_arg_bytes) {
  if (!t || !t->groups || !t->mg_records || t->n_groups <= 0 || t->group_max_rows <= 0 || t->group_max_rows > 65535) return 0;
  if (ld_floats <= 0 || ld_floats % 4 != 0 || ld_arg_bytes % 16 != 0 || ld_floats > 4 * kConsumers || ld_arg_bytes > (1 << 20)) return 0;
  const MgGeom g = mg_geom(t->group_max_rows, kRecB, (int)ld_floats * 4, (int)ld_arg_bytes);
  return g.n_stages >= 2 ? 1 : 0;   // at least a double buffer
}

extern "C" int dcgc_mg_gather_sum(const float* x, int64_t ld_x, const dcgc_topology* t, int32_t transposed,
                                  int32_t width, const float* addend, int64_t ld_add, float* out, int64_t ld_out,
                                  void* stream) {
  DCGC_CHECK_ARG(t && width >= 0 && ld_x >= width && ld_out >= width && (!addend || ld_add >= width),
                 "dcgc_mg_gather_sum: bad sizes");
  if (t->n_atoms == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && out, "dcgc_mg_gather_sum: null pointer");
  DCGC_CHECK_ARG(!transposed || t->symmetric, "dcgc_mg_gather_sum: the transposed lists are bucketed only for a symmetric adjacency");
  DCGC_CHECK_ARG(dcgc_mg_supported(t, ld_x, 0) && width % 4 == 0 && ld_out % 4 == 0 && al16(x) && al16(out) &&
                     (!addend || (ld_add % 4 == 0 && al16(addend))),
                 "dcgc_mg_gather_sum: layout not supported by the staged kernel (see dcgc_mg_supported)");
  cudaStream_t st = (cudaStream_t)stream;
  DcgcProfScope prof_scope("dcgc_gather_sum", st);
  const MgGeom geo = mg_geom(t->group_max_rows, transposed ? kRecB : kRecF, (int)ld_x * 4, 0);
  GatherSumOp op;
  op.addend = addend; op.ld_add = (int)ld_add; op.out = out; op.ld_out = (int)ld_out;
  return mg_launch(t, geo, mg_rec(t, transposed), x, nullptr, width / 4, op, st, "dcgc_mg_gather_sum");
}

extern "C" int dcgc_mg_pool_fwd(const float* x, int64_t ld_x, const float* scale, const float* shift,
                                const dcgc_topology* t, int32_t width, float* out, int64_t ld_out, uint8_t* arg,
                                int64_t ld_arg, void* stream) {
  DCGC_CHECK_ARG(t && width >= 0 && ld_x >= width && ld_out >= width, "dcgc_mg_pool_fwd: bad sizes");
  DCGC_CHECK_ARG((scale == nullptr) == (shift == nullptr), "dcgc_mg_pool_fwd: scale and shift go together");
  if (t->n_atoms == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(x && out, "dcgc_mg_pool_fwd: null pointer");
  DCGC_CHECK_ARG(dcgc_mg_supported(t, ld_x, 0) && width % 4 == 0 && ld_out % 4 == 0 && al16(x) && al16(out) &&
                     (!arg || (ld_arg >= width && ld_arg % 4 == 0 && (reinterpret_cast<uintptr_t>(arg) & 3) == 0)) &&
                     (!scale || (al16(scale) && al16(shift))),
                 "dcgc_mg_pool_fwd: layout not supported by the staged kernel (see dcgc_mg_supported)");
  cudaStream_t st = (cudaStream_t)stream;
  DcgcProfScope prof_scope("dcgc_pool_fwd", st);
  const MgGeom geo = mg_geom(t->group_max_rows, kRecF, (int)ld_x * 4, 0);
  const int cg = width / 4;
#define DCGC_MG_POOL(A, R)                                                                     \
  do {                                                                                         \
    PoolFwdOp<A, R> op;                                                                        \
    op.scale = scale; op.shift = shift; op.out = out; op.ld_out = (int)ld_out; op.arg = arg; op.ld_arg = (int)ld_arg; \
    op.sc = op.sh = make_float4(0.f, 0.f, 0.f, 0.f);                                           \
    return mg_launch(t, geo, mg_rec(t, 0), x, nullptr, cg, op, st, "dcgc_mg_pool_fwd");        \
  } while (0)
  if (scale) { if (arg) DCGC_MG_POOL(true, true); else DCGC_MG_POOL(true, false); }
  else { if (arg) DCGC_MG_POOL(false, true); else DCGC_MG_POOL(false, false); }
#undef DCGC_MG_POOL
}

extern "C" int dcgc_mg_pool_bwd(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t ld_arg, const float* scale,
                                const dcgc_topology* t, int32_t width, float* dx, int64_t ld_dx, void* stream) {
  DCGC_CHECK_ARG(t && width >= 0 && ld_dy >= width && ld_dx >= width && ld_arg >= width, "dcgc_mg_pool_bwd: bad sizes");
  if (t->n_atoms == 0 || width == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dy && arg && dx, "dcgc_mg_pool_bwd: null pointer");
  DCGC_CHECK_ARG(t->symmetric, "dcgc_mg_pool_bwd: the transposed lists are bucketed only for a symmetric adjacency");
  DCGC_CHECK_ARG(dcgc_mg_supported(t, ld_dy, ld_arg) && width % 4 == 0 && ld_dx % 4 == 0 && al16(dy) && al16(dx) &&
                     al16(arg) && (!scale || al16(scale)),
                 "dcgc_mg_pool_bwd: layout not supported by the staged kernel (see dcgc_mg_supported)");
  cudaStream_t st = (cudaStream_t)stream;
  DcgcProfScope prof_scope("dcgc_pool_bwd", st);
  const MgGeom geo = mg_geom(t->group_max_rows, kRecB, (int)ld_dy * 4, (int)ld_arg);
  const int cg = width / 4;
  if (scale) {
    PoolBwdOp<true, false> op{};
    op.scale = scale; op.dx = dx; op.ld_dx = (int)ld_dx;
    return mg_launch(t, geo, mg_rec(t, 1), dy, arg, cg, op, st, "dcgc_mg_pool_bwd");
  }
  PoolBwdOp<false, false> op{};
  op.dx = dx; op.ld_dx = (int)ld_dx;
  return mg_launch(t, geo, mg_rec(t, 1), dy, arg, cg, op, st, "dcgc_mg_pool_bwd");
}

extern "C" int dcgc_mg_pool_bwd_stats(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t ld_arg,
                                      const dcgc_topology* t, int32_t width, float* dx, int64_t ld_dx, const float* y,
                                      int64_t ld_y, const float* mean, double* part, int32_t* n_chunks, void* stream) {
  return dcgc_mg_pool_bwd_stats_fin(dy, ld_dy, arg, ld_arg, t, width, dx, ld_dx, y, ld_y, mean, part, n_chunks, nullptr,
                                    stream);
}

// the same with an optional BatchNorm finalize by the last CTA (fused engine; not part of the ABI)
int dcgc_mg_pool_bwd_stats_fin(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t ld_arg, const dcgc_topology* t,
                               int32_t width, float* dx, int64_t ld_dx, const float* y, int64_t ld_y, const float* mean,
                               double* part, int32_t* n_chunks, const DcgcBnFin* fin, void* stream) {
  DCGC_CHECK_ARG(t && width > 0 && ld_dy >= width && ld_dx >= width && ld_arg >= width && ld_y >= width,
                 "dcgc_mg_pool_bwd_stats: bad sizes");
  DCGC_CHECK_ARG(n_chunks, "dcgc_mg_pool_bwd_stats: null n_chunks");
  *n_chunks = 0;
  if (t->n_atoms == 0) return DCGC_OK;
  DCGC_CHECK_ARG(dy && arg && dx && y && mean && part, "dcgc_mg_pool_bwd_stats: null pointer");
  DCGC_CHECK_ARG(t->symmetric, "dcgc_mg_pool_bwd_stats: the transposed lists are bucketed only for a symmetric adjacency");
  DCGC_CHECK_ARG(dcgc_mg_supported(t, ld_dy, ld_arg) && width % 4 == 0 && ld_dx % 4 == 0 && ld_y % 4 == 0 && al16(dy) &&
                     al16(dx) && al16(arg) && al16(y) && al16(mean),
                 "dcgc_mg_pool_bwd_stats: layout not supported by the staged kernel (see dcgc_mg_supported)");
  cudaStream_t st = (cudaStream_t)stream;
  DcgcProfScope prof_scope("dcgc_pool_bwd", st);
  const MgGeom geo = mg_geom(t->group_max_rows, kRecB, (int)ld_dy * 4, (int)ld_arg);
  // the CTA-level reduction reuses the stage area: [row lanes][2][width] floats
  DCGC_CHECK_ARG((int64_t)(kConsumers / (width / 4)) * 2 * width * 4 <= (int64_t)geo.n_stages * geo.stage_bytes,
                 "dcgc_mg_pool_bwd_stats: stage area too small for the reduction");
  PoolBwdOp<false, true> op{};
  op.dx = dx; op.ld_dx = (int)ld_dx;
  op.y = y; op.ld_y = (int)ld_y; op.mean = mean; op.part = part; op.width = width;
  if (fin) op.bnfin = *fin;
  const int sms = dcgc_tc_num_sms();
  *n_chunks = t->n_groups < sms ? t->n_groups : sms;     // = the grid of mg_launch: one row of partials per CTA
  return mg_launch(t, geo, mg_rec(t, 1), dy, arg, width / 4, op, st, "dcgc_mg_pool_bwd_stats");
}

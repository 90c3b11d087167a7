// Launch accounting and per-entry-point CUDA-event timing (bench.py: gpu_launches and the live
// roofline measurement).  Events are recorded on the stream the kernel is launched on.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.h"

std::atomic<long long> g_dcgc_launches{0};

// Launches inside an active timing scope (between its two events) are plain: the events cut the programmatic chain
// anyway, and a launch that carries the attribute behind an event record measured 1.7 us longer (r6j).
static thread_local int t_scope_depth = 0;
bool dcgc_pdl_on() {
  static const bool on = [] { const char* e = getenv("DCGC_PDL"); return !(e && e[0] == '0'); }();
  return on && t_scope_depth == 0;
}

namespace {
std::mutex g_mu;
char g_name[64] = "";
std::atomic<int> g_on{0};
struct Ev { cudaEvent_t e0, e1; const char* name; };
std::vector<Ev> g_events;
}  // namespace

DcgcProfScope::DcgcProfScope(const char* name, cudaStream_t st) : e1_(nullptr), st_(st) {
  if (!g_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (strcmp(g_name, "*") != 0 && strcmp(name, g_name) != 0) return;
  cudaEvent_t e0;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) { e1_ = nullptr; return; }
  cudaEventRecord(e0, st);
  g_events.push_back(Ev{e0, e1_, name});
  ++t_scope_depth;
}

DcgcProfScope::~DcgcProfScope() {
  if (e1_) {
    cudaEventRecord(e1_, st_);
    --t_scope_depth;
  }
}

extern "C" long long dcgc_launch_count(void) { return g_dcgc_launches.load(); }

extern "C" int dcgc_profile_begin(const char* entry_name) {
  DCGC_CHECK_ARG(entry_name && strlen(entry_name) < sizeof(g_name), "dcgc_profile_begin: bad name");
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& ev : g_events) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); }
  g_events.clear();
  strcpy(g_name, entry_name);
  g_on.store(1);
  return DCGC_OK;
}

extern "C" int dcgc_profile_end(double* total_ms, long long* launches) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on.store(0);
  double ms = 0.0;
  for (auto& ev : g_events) {
    float t = 0.f;
    DCGC_CUDA_CALL(cudaEventSynchronize(ev.e1));
    DCGC_CUDA_CALL(cudaEventElapsedTime(&t, ev.e0, ev.e1));
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = (long long)g_events.size();
  for (auto& ev : g_events) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); }
  g_events.clear();
  g_name[0] = 0;
  return DCGC_OK;
}

// Ends a dcgc_profile_begin("*") session: one text line "name total_ms calls" per profiled scope.
extern "C" int dcgc_profile_report(char* out, int64_t cap) {
  DCGC_CHECK_ARG(out && cap > 0, "dcgc_profile_report: null buffer");
  std::lock_guard<std::mutex> lk(g_mu);
  g_on.store(0);
  struct Agg { const char* name; double ms; long long n; };
  std::vector<Agg> agg;
  for (auto& ev : g_events) {
    float t = 0.f;
    DCGC_CUDA_CALL(cudaEventSynchronize(ev.e1));
    DCGC_CUDA_CALL(cudaEventElapsedTime(&t, ev.e0, ev.e1));
    bool found = false;
    for (auto& a : agg)
      if (strcmp(a.name, ev.name) == 0) { a.ms += t; a.n++; found = true; break; }
    if (!found) agg.push_back(Agg{ev.name, (double)t, 1});
  }
  int64_t off = 0;
  out[0] = 0;
  for (auto& a : agg) {
    int w = snprintf(out + off, (size_t)(cap - off), "%s %.6f %lld\n", a.name, a.ms, a.n);
    if (w < 0 || off + w >= cap) break;
    off += w;
  }
  for (auto& ev : g_events) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); }
  g_events.clear();
  g_name[0] = 0;
  return DCGC_OK;
}

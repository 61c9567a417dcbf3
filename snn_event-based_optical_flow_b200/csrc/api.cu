// Library-level plumbing: thread-local error string, launch counter, device queries.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace snnflow {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- per-launch profiler ----------------------------------------------------------------------
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
  double bytes, flops;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static thread_local cudaStream_t g_prof_stream = nullptr;
static thread_local bool g_prof_open = false;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void prof_begin(const char* name, cudaStream_t st, double algo_bytes, double algo_flops) {
  if (!g_prof_on) return;
  ProfRec r{name, get_event(), get_event(), algo_bytes, algo_flops};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  g_prof_stream = st;
  g_prof_open = true;
}

void prof_end() {
  if (!g_prof_on || !g_prof_open) return;
  cudaEventRecord(g_recs.back().b, g_prof_stream);
  g_prof_open = false;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("SNNFLOW_PDL");
    on = (v && v[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

}  // namespace snnflow

extern "C" int snnflow_abi_version(void) { return SNNFLOW_ABI_VERSION; }
extern "C" const char* snnflow_last_error(void) { return snnflow::g_err; }
extern "C" uint64_t snnflow_launch_count(void) { return snnflow::g_launches.load(std::memory_order_relaxed); }

// Per-launch profiler control.  enable(1) starts recording (and clears old records); summary() synchronises
// the device and writes one line per kernel name: "name launches total_ms algo_bytes algo_flops\n".
extern "C" int snnflow_profile_enable(int on) {
  using namespace snnflow;
  for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_recs.clear();
  g_prof_on = on != 0;
  return SNNFLOW_OK;
}

extern "C" int snnflow_profile_summary(char* buf, size_t cap) {
  using namespace snnflow;
  if (!buf || cap == 0) return SNNFLOW_EINVAL;
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("profile_summary: sync failed"); return SNNFLOW_ECUDA; }
  struct Agg { long n = 0; double ms = 0, bytes = 0, flops = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    Agg& a = agg[r.name];
    a.n += 1; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
  }
  size_t off = 0;
  buf[0] = 0;
  for (auto& kv : agg) {
    int w = snprintf(buf + off, cap - off, "%s %ld %.6f %.0f %.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms,
                     kv.second.bytes, kv.second.flops);
    if (w < 0 || (size_t)w >= cap - off) break;
    off += (size_t)w;
  }
  return SNNFLOW_OK;
}

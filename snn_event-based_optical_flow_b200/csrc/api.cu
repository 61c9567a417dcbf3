// Library-level plumbing: thread-local error string, launch counter, device queries.
#include <stdarg.h>

#include <atomic>

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

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

}  // namespace snnflow

extern "C" int snnflow_abi_version(void) { return SNNFLOW_ABI_VERSION; }
extern "C" const char* snnflow_last_error(void) { return snnflow::g_err; }
extern "C" uint64_t snnflow_launch_count(void) { return snnflow::g_launches.load(std::memory_order_relaxed); }

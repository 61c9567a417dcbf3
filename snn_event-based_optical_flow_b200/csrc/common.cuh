// Shared helpers for the snnflow CUDA translation units (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/snnflow.h"

namespace snnflow {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// Optional per-launch timing (snnflow_profile_enable): prof_begin records a CUDA event on the launching
// stream before the kernel, prof_end one after it.  algo_bytes / algo_flops are the ALGORITHMIC traffic and
// arithmetic of the launch (what DESIGN.md states per kernel), used for the roofline fractions in bench.py.
void prof_begin(const char* name, cudaStream_t st, double algo_bytes, double algo_flops = 0.0);
void prof_end();

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SNNFLOW_ECUDA;
  }
  count_launch();
  prof_end();
  return SNNFLOW_OK;
}

#define SNNFLOW_REQUIRE(cond, msg)                    \
  do {                                                \
    if (!(cond)) {                                    \
      snnflow::set_error("%s: %s", __func__, msg);    \
      return SNNFLOW_EINVAL;                          \
    }                                                 \
  } while (0)

#define SNNFLOW_CUDA(call)                                                   \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      snnflow::set_error("%s: %s", #call, cudaGetErrorString(e__));          \
      return SNNFLOW_ECUDA;                                                  \
    }                                                                        \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// number of SMs of the current device (cached)
int sm_count();
// programmatic dependent launch on (default) / off (SNNFLOW_PDL=0)
bool pdl_enabled();

// ---- programmatic dependent launch -----------------------------------------------------------------
// Consecutive launches of a window run back to back on one stream; each is short (20 - 100 us), so the drain of one grid
// and the ramp-up of the next (launch latency, barrier / TMEM set-up, weight staging) are a visible fraction of a step.
// Kernels launched through launch_pdl() may begin while the previous grid is still retiring; they call pdl_wait()
// before their first access to global memory the previous launches wrote (it returns once those grids have completed and
// flushed), and pdl_launch_dependents() as early as possible.  Both are no-ops for a normal launch.
// 256-bit global loads / stores (sm_100): eight fp32 of one pixel in the c8 layout = one 32-byte sector per lane.
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :: "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(p)
               : "memory");
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Cooperative launch: every CTA of the grid is resident at once (the launch fails otherwise), which is what the grid
// barrier of the persistent multi-bin kernels relies on.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- grid barrier of the persistent multi-bin kernels (one counter, zeroed by the launcher) --------------------------
// arrive: called by ONE thread of a CTA after a CTA-level barrier over the threads whose global writes must be published.
__device__ __forceinline__ void grid_bar_arrive(unsigned int* bar) {
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
// wait until `target` arrivals; traps instead of hanging the device if a CTA never arrives
__device__ __forceinline__ void grid_bar_wait(const unsigned int* bar, unsigned int target) {
  const long long t0 = clock64();
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    if (v >= target) break;
    __nanosleep(40);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }
// ---- per-tile progress flags of the time-fused recurrent kernels -------------------------------------------------------
// flag[tile] = number of time bins this launch has finished for the tile (monotonic; zeroed by the launcher).  The
// epilogue that completes (tile, bin) publishes bin + 1 with release semantics after its plane writes; the TMA producer
// of a CTA that needs halo rows of that tile for the next bin acquires it.  Point-to-point, so a CTA only ever waits for
// the tiles it reads - a wavefront through (tile, bin) space with no grid-wide barrier.
__device__ __forceinline__ void tile_flag_set(unsigned int* flag, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ void tile_flag_wait(const unsigned int* flag, unsigned int target) {
  const long long t0 = clock64();
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= target) break;
    __nanosleep(32);
    if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must not hang the device
  }
}
// coherent 256-bit load (data written earlier by this same launch)
__device__ __forceinline__ void ldg256_coherent(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p) : "memory");
}

__device__ __forceinline__ uint4 ldg128_coherent(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 4 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Gaussian pdf (models/spiking_util.py:6-10)
__device__ __forceinline__ float gauss_pdf(float x, float mu, float sigma) {
  const float d = x - mu;
  return expf(-(d * d) / (2.0f * sigma * sigma)) / (sigma * 2.50662827463100050242f);   // sigma * sqrt(2 pi)
}
// surrogate gradient d spike / d u  (models/spiking_util.py:42,60-64,78,92)
__device__ __forceinline__ float surrogate(float u, float width, int kind) {
  if (kind == SNNFLOW_SG_MULTIGAUSS)   // MultiGaussSpike, Yin et al. 2021 (:60-64)
    return 1.15f * gauss_pdf(u, 0.f, width) - 0.15f * gauss_pdf(u, width, 6.0f * width) - 0.15f * gauss_pdf(u, -width, 6.0f * width);
  if (kind == SNNFLOW_SG_ARCTAN) return 1.0f / (1.0f + width * u * u);
  if (kind == SNNFLOW_SG_SUPERSPIKE) {
    float d = 1.0f + width * fabsf(u);
    return 1.0f / (d * d);
  }
  return fmaxf(1.0f - width * fabsf(u), 0.0f);
}

// compile-time surrogate kind + fast reciprocal (MUFU.RCP, ~1 ulp): the gradient tier tolerates it (rel 1e-4)
__device__ __forceinline__ float fast_rcp(float x) {   // one MUFU.RCP; the arguments here are >= 1 and finite
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int KIND>
__device__ __forceinline__ float surrogate_fast(float u, float width) {
  if (KIND == SNNFLOW_SG_ARCTAN) return fast_rcp(fmaf(width * u, u, 1.0f));
  if (KIND == SNNFLOW_SG_SUPERSPIKE) {
    const float d = fmaf(width, fabsf(u), 1.0f);
    return fast_rcp(d * d);
  }
  return fmaxf(1.0f - width * fabsf(u), 0.0f);
}

// two fp32 values -> packed bf16 hi pair + bf16 lo pair (value = hi + lo to 16 mantissa bits); even element in the low half
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace snnflow

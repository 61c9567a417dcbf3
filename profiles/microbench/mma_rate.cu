// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, cta_group::1, M = 128) with NO-SWIZZLE shared-memory
// operands, as used by the window kernels, for several N.  One CTA per SM, one thread issues `n_mma` MMAs into one
// accumulator, commits and waits; cycles per MMA = (clock after the commit completes - clock before the first MMA) / n.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../snn_event-based_optical_flow_b200/csrc/tcgen05.cuh"
using namespace snnflow;
namespace snnflow { void set_error(const char*, ...) {} void count_launch(int) {} void prof_begin(const char*, cudaStream_t, double, double) {} void prof_end() {} int sm_count() { return 148; } }

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int n_mma, int a_major_mn, int vary, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  unsigned char* buf = smem + 1024;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(buf)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t base16 = smem_u32(buf) >> 4;
    const uint32_t idesc = make_idesc(128, N, 1, a_major_mn, a_major_mn);
    // K-major: LBO (K chunks) 8320 B, SBO 128 B.  MN-major: LBO 128 B (k groups), SBO 2176 B (chunk groups)
    const uint32_t lbo = a_major_mn ? 128u : 8320u, sbo = a_major_mn ? 2176u : 128u;
    const uint32_t lo_c = ((lbo >> 4) & 0x3FFF) << 16, hi = desc_hi(sbo);
    const uint32_t b16 = base16 + (100 * 1024 >> 4);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t ao = vary ? (uint32_t)((i % 9) * 131 + (i & 1) * 1040) : 0u;   // tap-like shifts / k-steps
      umma_f16_split(tmem, lo_c | (base16 + ao), hi, lo_c | (b16 + (vary ? (uint32_t)((i % 9) * 256) : 0u)), hi, idesc, i > 0);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  const int smem = 1024 + 160 * 1024 + 16 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 4096;
  for (int mn = 0; mn < 2; ++mn)
    for (int vary = 0; vary < 2; ++vary)
      for (int N : {16, 32, 64, 96, 128, 192, 256}) {
        for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<148, 128, smem>>>(N, n_mma, mn, vary, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("%s vary=%d N=%3d : %7.1f cycles/MMA  (%s)\n", mn ? "MN-major" : "K-major ", vary, N, (double)h / n_mma, cudaGetErrorString(e));
      }
  return 0;
}

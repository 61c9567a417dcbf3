// Microbenchmark: per-SM throughput of cp.async.bulk (1-D TMA) global -> shared as a function of the copy size and of how
// many copies one warp issues per pipeline slot.  148 CTAs x 1 warp; a ring of D slots; the warp re-uses a slot as soon as
// the slot's mbarrier reports the bytes of the copies issued D iterations earlier.  Source: a 2 GiB buffer walked linearly
// (every byte read once: DRAM stream, no reuse).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu && ./tma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// per iteration: n_copies copies of copy_bytes each (lane l issues copies l, l+32, ...), slot = n_copies * copy_bytes
// n_warps producer warps: warp w owns the iterations it = w, w + n_warps, ... (slot = it % depth)
__global__ void __launch_bounds__(128, 1) tma_rate_kernel(const unsigned char* src, size_t per_cta, uint32_t copy_bytes, int n_copies, int depth,
                                                          long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem + 1024;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const uint32_t slot_bytes = copy_bytes * n_copies;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
  const int iters = (int)(per_cta / slot_bytes);
  const long long t0 = clock64();
  for (int it = warp; it < iters; it += n_warps) {
    const int st = it % depth, use = it / depth;
    if (lane == 0) {
      if (use > 0) while (!mbar_try_wait(&full[st], (use - 1) & 1)) {}
      mbar_expect_tx(&full[st], slot_bytes);
    }
    __syncwarp();
    for (int c = lane; c < n_copies; c += 32)
      tma_bulk_g2s(ring + (size_t)st * slot_bytes + (size_t)c * copy_bytes, base + (size_t)it * slot_bytes + (size_t)c * copy_bytes, copy_bytes, &full[st]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int it = iters > depth ? iters - depth : 0; it < iters; ++it) {
      const int st = it % depth, use = it / depth;
      while (!mbar_try_wait(&full[st], use & 1)) {}
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t total = (size_t)2 << 30;
  unsigned char* src;
  cudaMalloc(&src, total);
  cudaMemset(src, 1, total);
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * sms);
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const size_t per_cta = (total / sms) & ~(size_t)((1 << 20) - 1);
  printf("SMs %d, %zu MiB per CTA\n", sms, per_cta >> 20);
  printf("%10s %8s %6s %6s %10s %12s %10s\n", "copy B", "copies", "depth", "warps", "ms", "GB/s", "B/clk/SM");
  const uint32_t sizes[] = {2080, 6240, 8192};
  for (uint32_t cb : sizes)
    for (int nc : {1, 4, 12}) {
      for (int depth : {4, 6})
      for (int nw : {1, 2, 4}) {
        const size_t slot = (size_t)cb * nc;
        if (slot * depth > 200 * 1024 || slot < 8192) continue;
        if (slot * depth > 200 * 1024) continue;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          tma_rate_kernel<<<sms, 32 * nw, 1024 + slot * depth>>>(src, per_cta, cb, nc, depth, cyc);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long h[256];
        cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i] / sms;
        const size_t iters = per_cta / slot;
        const double bytes = (double)iters * slot * sms;
        printf("%10u %8d %6d %6d %10.3f %12.1f %10.2f\n", cb, nc, depth, nw, ms, bytes / ms / 1e6, (double)iters * slot / avg);
        if (cudaGetLastError() != cudaSuccess) { printf("error\n"); return 1; }
      }
    }
  return 0;
}

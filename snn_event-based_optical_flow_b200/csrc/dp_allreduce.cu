// Data-parallel gradient exchange: SUM all-reduce of the flat gradient (~300 KB) over NVLink peer memory, as ONE kernel
// per rank that can sit inside the training step's CUDA graph.
//
// The reference has no distributed code; the step needs exactly one collective: d loss / d parameters is a SUM over the
// batch (loss/flow.py:228,261,291), so ranks that own disjoint samples add their gradients before clip_grad_norm_ and
// Adam (train_flow.py:262-271).  299 KB is latency bound: every rank simply reads every peer's buffer (one-shot
// all-reduce, R * 299 KB over NVLink 5 / NVSwitch) between two flag barriers, and adds the R terms in rank order -
// every rank computes bit-identical sums, run-to-run deterministic.
//
//   peer_bufs[r]  : rank r's symmetric gradient buffer (n floats), mapped into this process
//   peer_pads[r]  : rank r's symmetric signal pad (uint32 flags), slot [256 + cta * world + src_rank]
//   counter       : this rank's launch counter (device memory, starts at 0): flag values grow with every launch, so the
//                   kernel is replayable from a CUDA graph without host-side epoch arguments
// Each CTA owns a slice of the buffer and runs the two barriers on its own flag slots, so no grid-wide sync is needed.
// A spinning CTA waits for kernels on OTHER GPUs only (one rank per GPU).
#include "common.cuh"

namespace snnflow {

constexpr int AR_THREADS = 512;
constexpr int AR_MAX_WORLD = 16;
constexpr int AR_PAD_OFFSET = 256;   // uint32 slots left to the owner of the signal pad (torch's own barriers use the first ones)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void ar_barrier(uint32_t* const* pads, int rank, int world, int cta, uint32_t value) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(pads[peer] + AR_PAD_OFFSET + cta * world + rank, value);               // tell `peer` that this rank reached `value`
    const uint32_t* mine = pads[rank] + AR_PAD_OFFSET + cta * world + peer;
    unsigned long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - value) < 0) {                 // wait until `peer` reached it too
      if (++spins > (1ull << 31)) __trap();                               // a dead peer: fail instead of hanging
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AR_THREADS) dp_allreduce_kernel(const float* const* peer_bufs, uint32_t* const* peer_pads, float* out,
                                                                  unsigned int* counter, int rank, int world, size_t n) {
  __shared__ uint32_t s_epoch;
  const int cta = blockIdx.x;
  if (threadIdx.x == 0) s_epoch = counter[cta] + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  ar_barrier(peer_pads, rank, world, cta, 2 * epoch - 1);      // every rank's buffer holds this step's gradient
  const size_t n4 = n / 4, per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)cta * per, hi = lo + per < n4 ? lo + per : n4;
  for (size_t i = lo + threadIdx.x; i < hi; i += AR_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {                          // fixed order: identical result on every rank
      const float4 v = reinterpret_cast<const float4*>(peer_bufs[r])[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  if (cta == 0) {
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += AR_THREADS) {
      float acc = 0.f;
      for (int r = 0; r < world; ++r) acc += peer_bufs[r][i];
      out[i] = acc;
    }
  }
  ar_barrier(peer_pads, rank, world, cta, 2 * epoch);          // nobody overwrites its buffer while a peer still reads it
  if (threadIdx.x == 0) counter[cta] = epoch;
}

}  // namespace snnflow
using namespace snnflow;

extern "C" int snnflow_dp_allreduce_ctas(void) { return 8; }

extern "C" int snnflow_dp_allreduce_sum(const void* peer_bufs, const void* peer_pads, float* out, unsigned int* counter, int rank,
                                        int world, size_t n, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(peer_bufs && peer_pads && out && counter, "null pointer");
  SNNFLOW_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
  SNNFLOW_REQUIRE(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
  prof_begin("dp_allreduce", (cudaStream_t)stream, 4.0 * n * (world + 1));
  dp_allreduce_kernel<<<snnflow_dp_allreduce_ctas(), AR_THREADS, 0, (cudaStream_t)stream>>>(
      (const float* const*)peer_bufs, (uint32_t* const*)peer_pads, out, counter, rank, world, n);
  return check_launch("dp_allreduce_kernel");
}

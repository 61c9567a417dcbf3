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

// =================================================================================================
// The whole parameter update of a data-parallel step as ONE kernel per rank:
//   stage this rank's flat gradient into its symmetric buffer -> ONE cross-rank flag barrier -> SUM over the ranks (rank
//   order: bit-identical on every rank) -> sum of squares -> grid barrier on this GPU -> clip coefficient -> Adam.
// Replaces [copy into the symmetric buffer, dp_allreduce_kernel (two barriers), copy out, opt_sumsq, opt_clip_adam]
// = train_flow.py:262-271 with the gradient SUM of SURVEY.md section 8e in front.  The peers' gradients are read once and
// the reduced gradient never goes back to memory (optional `reduced` output for tests).
// Symmetric buffer of a rank: 2 slots x slot_floats floats; the slot alternates with the launch parity (device-side
// counter: replayable from a CUDA graph), so NO trailing barrier is needed: a rank rewrites a slot two launches later, after
// the next launch's barrier, which no peer passes before it has finished reading.  Slot layout: [n gradient | DPF_CTAS gate
// words]: a rank whose update gate is raised (inputs not bf16-exact, see snnflow_clip_adam) vetoes the update on ALL ranks,
// so the replicas stay identical.
// world == 1 (or peer_bufs == NULL): no staging, no barrier - the single-GPU optimizer step in one launch.
// =================================================================================================
constexpr int DPF_CTAS = 16;
constexpr int DPF_THREADS = 512;
constexpr int DPF_MAX_PER_THREAD = 16;   // elements per thread kept in registers (n <= 16 * 512 * 16 = 131072)

struct DpRank {   // everything that belongs to ONE rank
  const float* grad_local;
  unsigned int* counter;
  float *p, *m, *v;
  const float* hyper;
  long long* step;
  double* state;
  float* partials;
  unsigned int* grid_cnt;
  float* norm_out;
  const unsigned int* gate;
  float* reduced;
};

__device__ __forceinline__ void dp_clip_adam_body(const DpRank& R, float* const* peer_bufs, uint32_t* const* peer_pads, int rank,
                                                  int world, int64_t n, int64_t slot_floats, int cta) {
  const float* __restrict__ grad_local = R.grad_local;
  unsigned int* counter = R.counter;
  float *p = R.p, *m = R.m, *v = R.v;
  const float* hyper = R.hyper;
  long long* step = R.step;
  double* state = R.state;
  float* partials = R.partials;
  unsigned int* grid_cnt = R.grid_cnt;
  float* norm_out = R.norm_out;
  const unsigned int* gate = R.gate;
  float* reduced = R.reduced;
  __shared__ uint32_t s_epoch;
  __shared__ float s_red[DPF_THREADS / 32];
  __shared__ float s_coef;
  __shared__ int s_veto;
  const int tid = threadIdx.x;
  if (tid == 0) s_epoch = counter[cta] + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int64_t per = (n + DPF_CTAS - 1) / DPF_CTAS, lo = (int64_t)cta * per, hi = lo + per < n ? lo + per : n;
  const bool multi = world > 1 && peer_bufs != nullptr;
  const unsigned int my_gate = (gate != nullptr && *gate != 0u) ? 1u : 0u;
  int veto = (int)my_gate;
  float g[DPF_MAX_PER_THREAD];
  if (multi) {
    float* mine = peer_bufs[rank] + (int64_t)(epoch & 1u) * slot_floats;
    for (int64_t i = lo + tid; i < hi; i += DPF_THREADS) mine[i] = grad_local[i];
    if (tid == 0) mine[n + cta] = my_gate ? 1.f : 0.f;
    ar_barrier(peer_pads, rank, world, cta, epoch);          // every rank's slice `cta` of this step is in place
#pragma unroll
    for (int k = 0; k < DPF_MAX_PER_THREAD; ++k) {
      const int64_t i = lo + tid + (int64_t)k * DPF_THREADS;
      float acc = 0.f;
      if (i < hi)
        for (int r = 0; r < world; ++r)                       // fixed order: identical result on every rank
          acc += __ldcg(peer_bufs[r] + (int64_t)(epoch & 1u) * slot_floats + i);
      g[k] = acc;
    }
    if (tid < world) veto = __ldcg(peer_bufs[tid] + (int64_t)(epoch & 1u) * slot_floats + n + cta) != 0.f;
    else veto = 0;
    veto = __syncthreads_or(veto);
  } else {
#pragma unroll
    for (int k = 0; k < DPF_MAX_PER_THREAD; ++k) {
      const int64_t i = lo + tid + (int64_t)k * DPF_THREADS;
      g[k] = i < hi ? grad_local[i] : 0.f;
    }
  }
  // sum of squares of the reduced gradient: per-CTA partial -> grid barrier -> every CTA adds the partials in the same order
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < DPF_MAX_PER_THREAD; ++k) ss = fmaf(g[k], g[k], ss);
  ss = warp_sum(ss);
  if ((tid & 31) == 0) s_red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DPF_THREADS / 32; ++w) t += s_red[w];
    partials[cta] = t;
    if (cta == 0 && !veto) {   // Adam's step counter and bias corrections live on the device (graph replays advance them)
      const long long t0 = *step;
      const double p1 = (t0 == 0 ? 1.0 : state[0]) * (double)hyper[1], p2 = (t0 == 0 ? 1.0 : state[1]) * (double)hyper[2];
      state[0] = p1; state[1] = p2;
      state[2] = (double)hyper[0] / (1.0 - p1);
      state[3] = 1.0 / sqrt(1.0 - p2);
      *step = t0 + 1;
    }
    __threadfence();
    atomicAdd(grid_cnt, 1u);
    unsigned long long spins = 0;
    while (*(volatile unsigned int*)grid_cnt < (unsigned int)DPF_CTAS * epoch) {     // all CTAs are resident (cooperative launch)
      if (++spins > (1ull << 31)) __trap();
    }
    __threadfence();
    float tot = 0.f;
    for (int c = 0; c < DPF_CTAS; ++c) tot += __ldcg(partials + c);
    const float max_norm = hyper[4], total = sqrtf(tot);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.0f);      // clip_grad.py: clamp(max=1.0)
    if (max_norm > 0.f && total != total) coef = total;                      // ... which propagates a NaN norm
    s_coef = coef;
    s_veto = veto;
    if (cta == 0 && norm_out) *norm_out = total;
  }
  __syncthreads();
  if (!s_veto) {
    const float coef = s_coef, step_size = (float)__ldcg(state + 2), inv_bc2_sqrt = (float)__ldcg(state + 3);
    const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
#pragma unroll
    for (int k = 0; k < DPF_MAX_PER_THREAD; ++k) {
      const int64_t i = lo + tid + (int64_t)k * DPF_THREADS;
      if (i >= hi) break;
      const float gc = g[k] * coef;
      const float mn = fmaf(1.0f - b1, gc - m[i], m[i]);
      const float vn = fmaf(1.0f - b2, gc * gc, v[i] * b2);
      m[i] = mn; v[i] = vn;
      p[i] = p[i] - step_size * (mn / (sqrtf(vn) * inv_bc2_sqrt + eps));
    }
  }
  if (reduced) {
#pragma unroll
    for (int k = 0; k < DPF_MAX_PER_THREAD; ++k) {
      const int64_t i = lo + tid + (int64_t)k * DPF_THREADS;
      if (i < hi) reduced[i] = g[k];
    }
  }
  if (tid == 0) counter[cta] = epoch;
}

__global__ void __launch_bounds__(DPF_THREADS) dp_clip_adam_kernel(const DpRank R, float* const* peer_bufs, uint32_t* const* peer_pads,
                                                                   int rank, int world, int64_t n, int64_t slot_floats) {
  dp_clip_adam_body(R, peer_bufs, peer_pads, rank, world, n, slot_floats, (int)blockIdx.x);
}

// The SAME body with all ranks emulated on ONE GPU (tests on a single-GPU box): rank r = blockIdx.x / DPF_CTAS works on
// ranks[r]; "peer" buffers and pads are ordinary device allocations.  Blocks of different ranks wait for each other, so the
// launch is cooperative (world * DPF_CTAS <= number of SMs).
__global__ void __launch_bounds__(DPF_THREADS) dp_clip_adam_emulated_kernel(const DpRank* ranks, float* const* peer_bufs,
                                                                            uint32_t* const* peer_pads, int world, int64_t n,
                                                                            int64_t slot_floats) {
  const int rank = (int)blockIdx.x / DPF_CTAS;
  dp_clip_adam_body(ranks[rank], peer_bufs, peer_pads, rank, world, n, slot_floats, (int)blockIdx.x % DPF_CTAS);
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

extern "C" int snnflow_dp_clip_adam_ctas(void) { return DPF_CTAS; }
extern "C" int64_t snnflow_dp_clip_adam_max_n(void) { return (int64_t)DPF_CTAS * DPF_THREADS * DPF_MAX_PER_THREAD; }

extern "C" int snnflow_dp_clip_adam(const float* grad_local, const void* peer_bufs, const void* peer_pads, unsigned int* counter,
                                    int rank, int world, int64_t n, int64_t slot_floats, float* params, float* exp_avg,
                                    float* exp_avg_sq, const float* hyper, int64_t* step, double* state, float* partials,
                                    unsigned int* grid_counter, float* grad_norm, const unsigned int* gate, float* reduced,
                                    snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(grad_local && counter && params && exp_avg && exp_avg_sq && hyper && step && state && partials && grid_counter,
                  "null pointer");
  SNNFLOW_REQUIRE(n > 0 && n <= snnflow_dp_clip_adam_max_n(), "n out of range for the one-launch update");
  SNNFLOW_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
  SNNFLOW_REQUIRE(world == 1 || (peer_bufs && peer_pads && slot_floats >= n + DPF_CTAS), "symmetric buffers required for world > 1");
  DpRank R{grad_local, counter, params, exp_avg, exp_avg_sq, hyper, reinterpret_cast<long long*>(step), state, partials,
           grid_counter, grad_norm, gate, reduced};
  prof_begin("dp_clip_adam", (cudaStream_t)stream, 4.0 * n * (world + 7));
  SNNFLOW_CUDA(launch_coop(dp_clip_adam_kernel, dim3(DPF_CTAS), dim3(DPF_THREADS), 0, (cudaStream_t)stream, R,
                           (float* const*)peer_bufs, (uint32_t* const*)peer_pads, rank, world, n, slot_floats));
  return check_launch("dp_clip_adam_kernel");
}

// All `world` ranks on one GPU in one cooperative launch (see dp_clip_adam_emulated_kernel).  rank_ptrs: device array of
// world x 13 pointers in the order of snnflow_dp_clip_adam's per-rank arguments: grad_local, counter, params, exp_avg,
// exp_avg_sq, hyper, step, state, partials, grid_counter, grad_norm, gate, reduced.
extern "C" int snnflow_dp_clip_adam_emulated(const void* rank_ptrs, const void* peer_bufs, const void* peer_pads, int world,
                                             int64_t n, int64_t slot_floats, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(rank_ptrs && peer_bufs && peer_pads, "null pointer");
  SNNFLOW_REQUIRE(world >= 2 && world <= AR_MAX_WORLD && world * DPF_CTAS <= sm_count(), "world size does not fit one GPU");
  SNNFLOW_REQUIRE(n > 0 && n <= snnflow_dp_clip_adam_max_n() && slot_floats >= n + DPF_CTAS, "bad sizes");
  static_assert(sizeof(DpRank) == 13 * sizeof(void*), "DpRank must be 13 pointers");
  prof_begin("dp_clip_adam_emulated", (cudaStream_t)stream, 4.0 * n * world * (world + 7));
  SNNFLOW_CUDA(launch_coop(dp_clip_adam_emulated_kernel, dim3(world * DPF_CTAS), dim3(DPF_THREADS), 0, (cudaStream_t)stream,
                           (const DpRank*)rank_ptrs, (float* const*)peer_bufs, (uint32_t* const*)peer_pads, world, n, slot_floats));
  return check_launch("dp_clip_adam_emulated_kernel");
}

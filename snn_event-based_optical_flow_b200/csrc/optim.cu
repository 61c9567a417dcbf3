// Gradient-norm clipping + Adam as two launches over ONE flat parameter buffer.
//
// Reference: train_flow.py:264-271 - torch.nn.utils.clip_grad.clip_grad_norm_(model.parameters(), clip) followed by
// torch.optim.Adam(lr).step() - which PyTorch executes as ~20 small kernels per step (per-tensor norms, stack, norm,
// clamp, foreach multiply, eight multi-tensor Adam kernels) over 75 k parameters: pure launch latency inside a 3 ms
// step.  Here: launch 1 = per-block sums of squares of the flat gradient (fixed order) and the step counter; launch
// 2 = every block re-reduces the partials in the same fixed order (identical norm in every block, deterministic),
// derives the clip coefficient and applies Adam to its slice.
//   total_norm = sqrt(sum g^2);  coef = min(1, max_norm / (total_norm + 1e-6));  g' = g * coef      (clip_grad_norm_)
//   m = m + (1 - b1) (g' - m);   v = b2 v + (1 - b2) g'^2                                            (Adam, no amsgrad,
//   p = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)                                  no weight decay)
#include "common.cuh"

namespace snnflow {

constexpr int OPT_THREADS = 256;
constexpr int OPT_PER_BLOCK = 1024;   // elements per block (4 per thread, all loads of a thread in flight together)
constexpr int OPT_PER_THREAD = OPT_PER_BLOCK / OPT_THREADS;

__device__ __forceinline__ float opt_block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
  }
  return t;   // valid in thread 0
}

// state (device, 4 doubles): {beta1^t, beta2^t, lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)}.  The powers are running
// products (one double multiply per step instead of a double-precision pow(), which costs ~20 us on this part's FP64
// rate); block 0 of the first launch advances them while the other blocks already sum their slices.
__global__ void __launch_bounds__(OPT_THREADS) opt_sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partials,
                                                                long long* __restrict__ step, const float* __restrict__ hyper,
                                                                double* __restrict__ state, const unsigned int* __restrict__ gate) {
  __shared__ float red[OPT_THREADS / 32];
  if (gate != nullptr && *gate != 0u) return;   // the step is vetoed (see snnflow_clip_adam): nothing advances
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const long long t = *step;
    const double p1 = (t == 0 ? 1.0 : state[0]) * (double)hyper[1], p2 = (t == 0 ? 1.0 : state[1]) * (double)hyper[2];
    state[0] = p1; state[1] = p2;
    state[2] = (double)hyper[0] / (1.0 - p1);
    state[3] = 1.0 / sqrt(1.0 - p2);
    *step = t + 1;   // Adam's step counter lives on the device (graph replays advance it)
  }
  const int64_t base = (int64_t)blockIdx.x * OPT_PER_BLOCK;
  float s = 0.f;
#pragma unroll
  for (int i = threadIdx.x; i < OPT_PER_BLOCK; i += OPT_THREADS) {
    const int64_t k = base + i;
    if (k < n) { const float x = g[k]; s = fmaf(x, x, s); }
  }
  const float t = opt_block_sum(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(OPT_THREADS) opt_clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                                    const float* __restrict__ hyper, const double* __restrict__ state,
                                                                    const float* __restrict__ partials, int n_part,
                                                                    float* __restrict__ norm_out, const unsigned int* __restrict__ gate) {
  __shared__ float s_coef;
  if (gate != nullptr && *gate != 0u) return;
  // this thread's slice first: the loads are in flight while warp 0 derives the clip coefficient
  const int64_t base = (int64_t)blockIdx.x * OPT_PER_BLOCK + threadIdx.x;
  float gk[OPT_PER_THREAD], mk[OPT_PER_THREAD], vk[OPT_PER_THREAD], pk[OPT_PER_THREAD];
#pragma unroll
  for (int i = 0; i < OPT_PER_THREAD; ++i) {
    const int64_t k = base + (int64_t)i * OPT_THREADS;
    const bool ok = k < n;
    gk[i] = ok ? g[k] : 0.f; mk[i] = ok ? m[k] : 0.f; vk[i] = ok ? v[k] : 0.f; pk[i] = ok ? p[k] : 0.f;
  }
  if (threadIdx.x < 32) {
    // fixed-order reduction of the block partials: lane-strided sums, then a butterfly - the same in every block
    float t = 0.f;
    for (int i = threadIdx.x; i < n_part; i += 32) t += partials[i];
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const float max_norm = hyper[4];
      const float total = sqrtf(t);
      float coef = 1.f;
      if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.0f);      // clip_grad.py: clamp(max=1.0)
      if (max_norm > 0.f && total != total) coef = total;                      // ... which propagates a NaN norm (fminf would not)
      s_coef = coef;
      if (blockIdx.x == 0 && norm_out) *norm_out = total;
    }
  }
  __syncthreads();
  const float coef = s_coef, step_size = (float)state[2], inv_bc2_sqrt = (float)state[3];
  const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
#pragma unroll
  for (int i = 0; i < OPT_PER_THREAD; ++i) {
    const int64_t k = base + (int64_t)i * OPT_THREADS;
    if (k >= n) break;
    const float gc = gk[i] * coef;
    const float mn = fmaf(1.0f - b1, gc - mk[i], mk[i]);
    const float vn = fmaf(1.0f - b2, gc * gc, vk[i] * b2);
    m[k] = mn; v[k] = vn;
    p[k] = pk[i] - step_size * (mn / (sqrtf(vn) * inv_bc2_sqrt + eps));
  }
}

}  // namespace snnflow
using namespace snnflow;

extern "C" int snnflow_clip_adam_partials(int64_t n) { return n <= 0 ? 0 : (int)ceil_div64(n, OPT_PER_BLOCK); }

extern "C" int snnflow_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 const float* hyper, int64_t* step, double* state, float* partials, float* grad_norm,
                                 const unsigned int* gate, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper && step && state && partials && n > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_part = snnflow_clip_adam_partials(n);
  prof_begin("opt_sumsq", st, 4.0 * n);
  opt_sumsq_kernel<<<n_part, OPT_THREADS, 0, st>>>(grads, n, partials, reinterpret_cast<long long*>(step), hyper, state, gate);
  int rc = check_launch("opt_sumsq_kernel");
  if (rc) return rc;
  prof_begin("opt_clip_adam", st, 28.0 * n);
  opt_clip_adam_kernel<<<n_part, OPT_THREADS, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, hyper, state, partials, n_part,
                                                     grad_norm, gate);
  return check_launch("opt_clip_adam_kernel");
}

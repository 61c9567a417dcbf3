// Event encodings: count image, "last event" image / mask, temporal-bilinear voxel grid.
// Reference: dataloader/encodings.py:30-85.  HBM-bound streaming reads of the event arrays (float4
// per thread) + L2 atomics into a small image; counts are integers, exact in fp32, so plain fp32
// reductions are order-independent (bit-exact + deterministic); fractional voxel weights go through a
// 64-bit fixed-point accumulator (2^-32 resolution) for run-to-run determinism.
#include "common.cuh"

namespace snnflow {

constexpr int EN_THREADS = 256;
constexpr double FIX_SCALE = 4294967296.0;       // 2^32
constexpr double FIX_INV = 1.0 / 4294967296.0;

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_fix(int64_t* addr, float v) {
  long long q = __double2ll_rn((double)v * FIX_SCALE);
  atomicAdd(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)q);
}

__device__ __forceinline__ void cnt_one(float x, float y, float p, float* out, int H, int W) {
  const int xi = (int)x, yi = (int)y;   // .long() truncation, encodings.py:39-42
  if (xi < 0 || xi >= W || yi < 0 || yi >= H || p == 0.f) return;
  // events_to_channels adds ps * (ps masked to its sign) = p*p to the channel of its sign (:77-83)
  red_add_f32(out + (size_t)(p > 0.f ? 0 : 1) * H * W + (size_t)yi * W + xi, p * p);
}

__global__ void __launch_bounds__(EN_THREADS) encode_cnt_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                                                const float* __restrict__ ps, float* __restrict__ out,
                                                                int64_t N, int H, int W, int vec_ok) {
  const int b = blockIdx.y;
  xs += (size_t)b * N; ys += (size_t)b * N; ps += (size_t)b * N;
  out += (size_t)b * 2 * H * W;
  const int64_t stride = (int64_t)gridDim.x * EN_THREADS;
  int64_t i = (int64_t)blockIdx.x * EN_THREADS + threadIdx.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = N >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    const float4* y4 = reinterpret_cast<const float4*>(ys);
    const float4* p4 = reinterpret_cast<const float4*>(ps);
    for (int64_t k = i; k < n4; k += stride) {
      float4 x = __ldg(x4 + k), y = __ldg(y4 + k), p = __ldg(p4 + k);
      cnt_one(x.x, y.x, p.x, out, H, W);
      cnt_one(x.y, y.y, p.y, out, H, W);
      cnt_one(x.z, y.z, p.z, out, H, W);
      cnt_one(x.w, y.w, p.w, out, H, W);
    }
    done = n4 << 2;
  }
  for (int64_t k = done + i; k < N; k += stride) cnt_one(xs[k], ys[k], ps[k], out, H, W);
}

__global__ void __launch_bounds__(EN_THREADS) encode_image_acc_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                                                      const float* __restrict__ ps, float* __restrict__ out,
                                                                      int64_t N, int H, int W) {
  const int64_t stride = (int64_t)gridDim.x * EN_THREADS;
  for (int64_t k = (int64_t)blockIdx.x * EN_THREADS + threadIdx.x; k < N; k += stride) {
    const int xi = (int)xs[k], yi = (int)ys[k];
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
    red_add_f32(out + (size_t)yi * W + xi, ps[k]);
  }
}

// accumulate=False: the value of the LAST event at each pixel survives (CPU index_put_ order).
__global__ void __launch_bounds__(EN_THREADS) encode_image_last_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                                                       int32_t* __restrict__ last, int64_t N, int H, int W) {
  const int64_t stride = (int64_t)gridDim.x * EN_THREADS;
  for (int64_t k = (int64_t)blockIdx.x * EN_THREADS + threadIdx.x; k < N; k += stride) {
    const int xi = (int)xs[k], yi = (int)ys[k];
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
    atomicMax(last + (size_t)yi * W + xi, (int32_t)(k + 1));
  }
}
__global__ void __launch_bounds__(EN_THREADS) encode_image_pick_kernel(const float* __restrict__ ps, const int32_t* __restrict__ last,
                                                                       float* __restrict__ out, int HW) {
  const int i = blockIdx.x * EN_THREADS + threadIdx.x;
  if (i >= HW) return;
  const int32_t k = last[i];
  out[i] = k > 0 ? ps[k - 1] : 0.f;
}

__global__ void __launch_bounds__(EN_THREADS) encode_voxel_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                                                  const float* __restrict__ ts, const float* __restrict__ ps,
                                                                  int64_t* __restrict__ acc, int64_t N, int nb, int H, int W,
                                                                  int round_ts) {
  const int64_t stride = (int64_t)gridDim.x * EN_THREADS;
  const float scale = (float)(nb - 1);
  for (int64_t k = (int64_t)blockIdx.x * EN_THREADS + threadIdx.x; k < N; k += stride) {
    const int xi = (int)xs[k], yi = (int)ys[k];
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
    float t = __fmul_rn(ts[k], scale);                      // encodings.py:56
    if (round_ts) t = rintf(t);                             // :58-59 (half to even)
    const float p = ps[k];
    const int b0 = (int)floorf(t);
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int bin = b0 + d;
      if (bin < 0 || bin >= nb) continue;
      const float w = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(t, (float)bin))));   // :63
      const float val = __fmul_rn(p, w);                                             // :64
      if (val != 0.f) red_add_fix(acc + ((size_t)bin * H + yi) * W + xi, val);
    }
  }
}

__global__ void __launch_bounds__(EN_THREADS) fix_to_float_kernel(const int64_t* __restrict__ acc, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * EN_THREADS + threadIdx.x;
  if (i < n) out[i] = (float)((double)acc[i] * FIX_INV);
}

int event_grid(int64_t N, int per_thread) {
  int64_t blocks = ceil_div64(N, (int64_t)EN_THREADS * per_thread);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace snnflow
using namespace snnflow;

extern "C" int snnflow_encode_cnt(const float* xs, const float* ys, const float* ps, float* out, int64_t N, int B,
                                  int H, int W, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(out && B > 0 && H > 0 && W > 0 && N >= 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SNNFLOW_CUDA(cudaMemsetAsync(out, 0, (size_t)B * 2 * H * W * sizeof(float), st));
  if (N == 0) return SNNFLOW_OK;
  SNNFLOW_REQUIRE(xs && ys && ps, "null event arrays");
  const int vec_ok = ((N & 3) == 0) && ((((uintptr_t)xs | (uintptr_t)ys | (uintptr_t)ps) & 15) == 0);
  prof_begin("encode_cnt", st, 12.0 * N * B + 8.0 * H * W * B);
  encode_cnt_kernel<<<dim3(event_grid(N, 4), B), EN_THREADS, 0, st>>>(xs, ys, ps, out, N, H, W, vec_ok);
  return check_launch("encode_cnt_kernel");
}

extern "C" int snnflow_encode_image(const float* xs, const float* ys, const float* ps, float* out, int32_t* scratch,
                                    int64_t N, int H, int W, int accumulate, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(out && H > 0 && W > 0 && N >= 0 && N < 2147483647LL, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SNNFLOW_CUDA(cudaMemsetAsync(out, 0, (size_t)H * W * sizeof(float), st));
  if (N == 0) return SNNFLOW_OK;
  SNNFLOW_REQUIRE(xs && ys && ps, "null event arrays");
  if (accumulate) {
    prof_begin("encode_image_acc", st, 12.0 * N + 4.0 * H * W);
    encode_image_acc_kernel<<<event_grid(N, 1), EN_THREADS, 0, st>>>(xs, ys, ps, out, N, H, W);
    return check_launch("encode_image_acc_kernel");
  }
  SNNFLOW_REQUIRE(scratch, "scratch required for accumulate=0");
  SNNFLOW_CUDA(cudaMemsetAsync(scratch, 0, (size_t)H * W * sizeof(int32_t), st));
  prof_begin("encode_image_last", st, 8.0 * N + 4.0 * H * W);
  encode_image_last_kernel<<<event_grid(N, 1), EN_THREADS, 0, st>>>(xs, ys, scratch, N, H, W);
  int rc = check_launch("encode_image_last_kernel");
  if (rc) return rc;
  prof_begin("encode_image_pick", st, 12.0 * H * W);
  encode_image_pick_kernel<<<ceil_div(H * W, EN_THREADS), EN_THREADS, 0, st>>>(ps, scratch, out, H * W);
  return check_launch("encode_image_pick_kernel");
}

extern "C" int snnflow_encode_voxel(const float* xs, const float* ys, const float* ts, const float* ps, float* out,
                                    int64_t* scratch, int64_t N, int num_bins, int H, int W, int round_ts,
                                    snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(out && scratch && num_bins > 0 && H > 0 && W > 0 && N >= 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)num_bins * H * W;
  SNNFLOW_CUDA(cudaMemsetAsync(scratch, 0, (size_t)n * sizeof(int64_t), st));
  if (N > 0) {
    SNNFLOW_REQUIRE(xs && ys && ts && ps, "null event arrays");
    prof_begin("encode_voxel", st, 16.0 * N + 4.0 * n);
    encode_voxel_kernel<<<event_grid(N, 1), EN_THREADS, 0, st>>>(xs, ys, ts, ps, scratch, N, num_bins, H, W, round_ts);
    int rc = check_launch("encode_voxel_kernel");
    if (rc) return rc;
  }
  prof_begin("fix_to_float", st, 12.0 * n);
  fix_to_float_kernel<<<(unsigned)ceil_div64(n, EN_THREADS), EN_THREADS, 0, st>>>(scratch, out, n);
  return check_launch("fix_to_float_kernel");
}

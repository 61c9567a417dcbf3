// Image of warped events: per-event flow gather, warp + bilinear/rounded splat, and their adjoints.
// Reference: utils/iwe.py:4-93 (purge_unfeasible, get_interpolation, interpolate), the gather of
// loss/flow.py:66-81 and compute_pol_iwe (utils/iwe.py:133-154).
// One thread per event: 16 B event + 8 B flow + 8 B polarity mask streamed from HBM, up to 8 L2
// atomics into small images.  The forward accumulates in 64-bit fixed point (2^-32), so the result is
// independent of the atomic order (deterministic) and equal to the correctly rounded fp32 sum up to
// 2^-33 per term.
#include "common.cuh"

namespace snnflow {

constexpr int IW_THREADS = 256;
constexpr double IW_FIX_SCALE = 4294967296.0;
constexpr double IW_FIX_INV = 1.0 / 4294967296.0;

__device__ __forceinline__ void fix_add(int64_t* addr, float v) {
  long long q = __double2ll_rn((double)v * IW_FIX_SCALE);
  atomicAdd(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)q);
}

// flat index y*W + x formed in fp32 and truncated (loss/flow.py:67-69,77; utils/iwe.py:110-116)
__device__ __forceinline__ long long flat_index(float y, float x, int W) {
  return (long long)__fadd_rn(__fmul_rn(y, (float)W), x);
}

__global__ void __launch_bounds__(IW_THREADS) flow_gather_fwd_kernel(const float* __restrict__ flow,
                                                                     const float4* __restrict__ events,
                                                                     float2* __restrict__ ev_flow, int64_t N, int H, int W) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * IW_THREADS + threadIdx.x;
  if (n >= N) return;
  const float4 e = __ldg(events + (size_t)b * N + n);   // (ts, y, x, p)
  const long long idx = flat_index(e.y, e.z, W);
  const size_t hw = (size_t)H * W;
  float2 f = make_float2(0.f, 0.f);
  if (idx >= 0 && idx < (long long)hw) {
    f.x = __ldg(flow + ((size_t)b * 2 + 1) * hw + idx);   // vertical component  (channel 1)
    f.y = __ldg(flow + ((size_t)b * 2 + 0) * hw + idx);   // horizontal component (channel 0)
  }
  ev_flow[(size_t)b * N + n] = f;
}

__global__ void __launch_bounds__(IW_THREADS) flow_gather_bwd_kernel(const float2* __restrict__ g_ev_flow,
                                                                     const float4* __restrict__ events,
                                                                     float* __restrict__ g_flow, int64_t N, int H, int W) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * IW_THREADS + threadIdx.x;
  if (n >= N) return;
  const float4 e = __ldg(events + (size_t)b * N + n);
  const long long idx = flat_index(e.y, e.z, W);
  const size_t hw = (size_t)H * W;
  if (idx < 0 || idx >= (long long)hw) return;
  const float2 g = g_ev_flow[(size_t)b * N + n];
  if (g.x != 0.f) atomicAdd(g_flow + ((size_t)b * 2 + 1) * hw + idx, g.x);
  if (g.y != 0.f) atomicAdd(g_flow + ((size_t)b * 2 + 0) * hw + idx, g.y);
}

struct Warp {
  float Y, X, dtS_y, dtS_x;
};

// warped location  (y, x) + (tref - ts) * flow * flow_scaling   (utils/iwe.py:37), rounded like torch
__device__ __forceinline__ void warp_event(const float4 e, const float2 f, float tref, float S, float& Y, float& X,
                                           float& dt) {
  dt = __fsub_rn(tref, e.x);
  Y = __fadd_rn(e.y, __fmul_rn(__fmul_rn(dt, f.x), S));
  X = __fadd_rn(e.z, __fmul_rn(__fmul_rn(dt, f.y), S));
}

__global__ void __launch_bounds__(IW_THREADS) iwe_splat_fwd_kernel(const float4* __restrict__ events,
                                                                   const float2* __restrict__ ev_flow,
                                                                   const float2* __restrict__ pol_mask,
                                                                   int64_t* __restrict__ acc, int64_t N, int H, int W,
                                                                   float tref, float S, int n_img, int ts_mode,
                                                                   float ts_ref, int round_idx) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * IW_THREADS + threadIdx.x;
  if (n >= N) return;
  const float4 e = __ldg(events + (size_t)b * N + n);
  const float2 f = __ldg(ev_flow + (size_t)b * N + n);
  const float2 pm = __ldg(pol_mask + (size_t)b * N + n);
  float Y, X, dt;
  warp_event(e, f, tref, S, Y, X, dt);
  const float tsw = ts_mode == 1 ? e.x : (ts_mode == 2 ? __fsub_rn(ts_ref, e.x) : 0.f);
  const size_t hw = (size_t)H * W;
  int64_t* img = acc + (size_t)b * n_img * hw;

  auto deposit = [&](float iy, float ix, float w) {
    // purge_unfeasible (utils/iwe.py:4-17): out-of-range corners get weight 0 (and index 0): no-op
    if (!(iy >= 0.f && iy < (float)H && ix >= 0.f && ix < (float)W) || w == 0.f) return;
    const size_t p = (size_t)((long long)__fadd_rn(__fmul_rn(iy, (float)W), ix));   // :68-69
    const float w0 = __fmul_rn(w, pm.x), w1 = __fmul_rn(w, pm.y);
    if (w0 != 0.f) fix_add(img + p, w0);
    if (w1 != 0.f) fix_add(img + hw + p, w1);
    if (n_img == 4) {
      const float wt = __fmul_rn(w, tsw);   // loss/flow.py:208-212: (weights * ts) * polarity_mask
      const float t0 = __fmul_rn(wt, pm.x), t1 = __fmul_rn(wt, pm.y);
      if (t0 != 0.f) fix_add(img + 2 * hw + p, t0);
      if (t1 != 0.f) fix_add(img + 3 * hw + p, t1);
    }
  };

  if (round_idx) {
    deposit(rintf(Y), rintf(X), 1.0f);   // torch.round = half to even (:41)
  } else {
    const float ty = floorf(Y), by = floorf(__fadd_rn(Y, 1.0f));   // :45-48
    const float lx = floorf(X), rx = floorf(__fadd_rn(X, 1.0f));
    const float wty = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(Y, ty))));   // :59
    const float wby = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(Y, by))));
    const float wlx = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(X, lx))));
    const float wrx = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(X, rx))));
    deposit(ty, lx, __fmul_rn(wty, wlx));   // :65 prod over (y, x)
    deposit(ty, rx, __fmul_rn(wty, wrx));
    deposit(by, lx, __fmul_rn(wby, wlx));
    deposit(by, rx, __fmul_rn(wby, wrx));
  }
}

__global__ void __launch_bounds__(IW_THREADS) iwe_fix_to_float_kernel(const int64_t* __restrict__ acc,
                                                                      float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * IW_THREADS + threadIdx.x;
  if (i < n) out[i] = (float)((double)acc[i] * IW_FIX_INV);
}

// d max(0, 1 - |d|) / d d with torch's tie rules: abs'(0) = 0; max(0, a) at a == 0 passes 1/2.
__device__ __forceinline__ float dtent(float d) {
  const float a = __fsub_rn(1.0f, fabsf(d));
  const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
  if (a > 0.f) return -sgn;
  if (a == 0.f) return -0.5f * sgn;
  return 0.f;
}

__global__ void __launch_bounds__(IW_THREADS) iwe_splat_bwd_kernel(const float4* __restrict__ events,
                                                                   const float2* __restrict__ ev_flow,
                                                                   const float2* __restrict__ pol_mask,
                                                                   const float* __restrict__ g_img,
                                                                   float2* __restrict__ g_ev_flow, int64_t N, int H, int W,
                                                                   float tref, float S, int n_img, int ts_mode,
                                                                   float ts_ref) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * IW_THREADS + threadIdx.x;
  if (n >= N) return;
  const float4 e = __ldg(events + (size_t)b * N + n);
  const float2 f = __ldg(ev_flow + (size_t)b * N + n);
  const float2 pm = __ldg(pol_mask + (size_t)b * N + n);
  float Y, X, dt;
  warp_event(e, f, tref, S, Y, X, dt);
  const float tsw = ts_mode == 1 ? e.x : (ts_mode == 2 ? __fsub_rn(ts_ref, e.x) : 0.f);
  const size_t hw = (size_t)H * W;
  const float* g = g_img + (size_t)b * n_img * hw;

  const float cy[2] = {floorf(Y), floorf(__fadd_rn(Y, 1.0f))};
  const float cx[2] = {floorf(X), floorf(__fadd_rn(X, 1.0f))};
  float wy[2], wx[2], dy[2], dx[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float ddy = __fsub_rn(Y, cy[k]), ddx = __fsub_rn(X, cx[k]);
    wy[k] = fmaxf(0.f, __fsub_rn(1.0f, fabsf(ddy)));
    wx[k] = fmaxf(0.f, __fsub_rn(1.0f, fabsf(ddx)));
    dy[k] = dtent(ddy);
    dx[k] = dtent(ddx);
  }
  float gY = 0.f, gX = 0.f;
#pragma unroll
  for (int iy = 0; iy < 2; ++iy)
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
      const float py = cy[iy], px = cx[ix];
      if (!(py >= 0.f && py < (float)H && px >= 0.f && px < (float)W)) continue;   // purged: weight * 0
      const size_t p = (size_t)((long long)__fadd_rn(__fmul_rn(py, (float)W), px));
      float G = g[p] * pm.x + g[hw + p] * pm.y;
      if (n_img == 4) G += tsw * (g[2 * hw + p] * pm.x + g[3 * hw + p] * pm.y);
      gY += G * dy[iy] * wx[ix];
      gX += G * wy[iy] * dx[ix];
    }
  const float k = dt * S;   // d warped / d flow
  g_ev_flow[(size_t)b * N + n] = make_float2(gY * k, gX * k);
}

}  // namespace snnflow
using namespace snnflow;

static int check_ev(const void* a, const void* b, int B, int64_t N, int H, int W) {
  SNNFLOW_REQUIRE(a && b, "null pointer");
  SNNFLOW_REQUIRE(B > 0 && N >= 0 && H > 0 && W > 0, "bad dims");
  SNNFLOW_REQUIRE((((uintptr_t)a) & 15) == 0, "events must be 16-byte aligned");
  SNNFLOW_REQUIRE((int64_t)H * W < (1 << 24), "H*W must stay below 2^24 (fp32 flat index)");
  return SNNFLOW_OK;
}

extern "C" int snnflow_flow_gather_fwd(const float* flow, const float* events, float* ev_flow, int B, int64_t N, int H,
                                       int W, snnflow_stream_t stream) {
  int rc = check_ev(events, flow, B, N, H, W);
  if (rc) return rc;
  SNNFLOW_REQUIRE(ev_flow || N == 0, "null output");
  if (N == 0) return SNNFLOW_OK;
  prof_begin("flow_gather_fwd", (cudaStream_t)stream, 32.0 * N * B);
  flow_gather_fwd_kernel<<<dim3((unsigned)ceil_div64(N, IW_THREADS), B), IW_THREADS, 0, (cudaStream_t)stream>>>(
      flow, (const float4*)events, (float2*)ev_flow, N, H, W);
  return check_launch("flow_gather_fwd_kernel");
}

extern "C" int snnflow_flow_gather_bwd(const float* g_ev_flow, const float* events, float* g_flow, int B, int64_t N,
                                       int H, int W, snnflow_stream_t stream) {
  int rc = check_ev(events, g_flow, B, N, H, W);
  if (rc) return rc;
  if (N == 0) return SNNFLOW_OK;
  SNNFLOW_REQUIRE(g_ev_flow, "null pointer");
  prof_begin("flow_gather_bwd", (cudaStream_t)stream, 32.0 * N * B);
  flow_gather_bwd_kernel<<<dim3((unsigned)ceil_div64(N, IW_THREADS), B), IW_THREADS, 0, (cudaStream_t)stream>>>(
      (const float2*)g_ev_flow, (const float4*)events, g_flow, N, H, W);
  return check_launch("flow_gather_bwd_kernel");
}

extern "C" int snnflow_iwe_splat_fwd(const float* events, const float* ev_flow, const float* pol_mask, float* out,
                                     int64_t* scratch, int B, int64_t N, int H, int W, float tref, float flow_scaling,
                                     int n_img, int ts_mode, float ts_ref, int round_idx, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(out && scratch, "null pointer");
  SNNFLOW_REQUIRE(n_img == 2 || n_img == 4, "n_img must be 2 or 4");
  SNNFLOW_REQUIRE(ts_mode >= 0 && ts_mode <= 2 && (n_img == 2 || ts_mode != 0), "bad ts_mode");
  SNNFLOW_REQUIRE(B > 0 && N >= 0 && H > 0 && W > 0 && (int64_t)H * W < (1 << 24), "bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)B * n_img * H * W;
  SNNFLOW_CUDA(cudaMemsetAsync(scratch, 0, (size_t)n * sizeof(int64_t), st));
  if (N > 0) {
    SNNFLOW_REQUIRE(events && ev_flow && pol_mask, "null pointer");
    SNNFLOW_REQUIRE((((uintptr_t)events) & 15) == 0 && (((uintptr_t)ev_flow | (uintptr_t)pol_mask) & 7) == 0, "misaligned");
    prof_begin("iwe_splat_fwd", st, 32.0 * N * B + 4.0 * n);
    iwe_splat_fwd_kernel<<<dim3((unsigned)ceil_div64(N, IW_THREADS), B), IW_THREADS, 0, st>>>(
        (const float4*)events, (const float2*)ev_flow, (const float2*)pol_mask, scratch, N, H, W, tref, flow_scaling,
        n_img, ts_mode, ts_ref, round_idx);
    int rc = check_launch("iwe_splat_fwd_kernel");
    if (rc) return rc;
  }
  prof_begin("iwe_fix_to_float", st, 12.0 * n);
  iwe_fix_to_float_kernel<<<(unsigned)ceil_div64(n, IW_THREADS), IW_THREADS, 0, st>>>(scratch, out, n);
  return check_launch("iwe_fix_to_float_kernel");
}

extern "C" int snnflow_iwe_splat_bwd(const float* events, const float* ev_flow, const float* pol_mask,
                                     const float* g_img, float* g_ev_flow, int B, int64_t N, int H, int W, float tref,
                                     float flow_scaling, int n_img, int ts_mode, float ts_ref, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(n_img == 2 || n_img == 4, "n_img must be 2 or 4");
  SNNFLOW_REQUIRE(B > 0 && N >= 0 && H > 0 && W > 0 && (int64_t)H * W < (1 << 24), "bad dims");
  if (N == 0) return SNNFLOW_OK;
  SNNFLOW_REQUIRE(events && ev_flow && pol_mask && g_img && g_ev_flow, "null pointer");
  SNNFLOW_REQUIRE((((uintptr_t)events) & 15) == 0 && (((uintptr_t)ev_flow | (uintptr_t)pol_mask | (uintptr_t)g_ev_flow) & 7) == 0,
                  "misaligned");
  prof_begin("iwe_splat_bwd", (cudaStream_t)stream, 40.0 * N * B + 4.0 * B * n_img * H * W);
  iwe_splat_bwd_kernel<<<dim3((unsigned)ceil_div64(N, IW_THREADS), B), IW_THREADS, 0, (cudaStream_t)stream>>>(
      (const float4*)events, (const float2*)ev_flow, (const float2*)pol_mask, g_img, (float2*)g_ev_flow, N, H, W, tref,
      flow_scaling, n_img, ts_mode, ts_ref);
  return check_launch("iwe_splat_bwd_kernel");
}

// Device functions shared by the image-of-warped-events kernels (iwe.cu) and the fused window loss (loss_window.cu):
// one event's warp + splat (utils/iwe.py:4-93) and its adjoint, with the reference's rounding and tie rules.
#pragma once
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

// warped location  (y, x) + (tref - ts) * flow * flow_scaling   (utils/iwe.py:37), rounded like torch
__device__ __forceinline__ void warp_event(const float4 e, const float2 f, float tref, float S, float& Y, float& X,
                                           float& dt) {
  dt = __fsub_rn(tref, e.x);
  Y = __fadd_rn(e.y, __fmul_rn(__fmul_rn(dt, f.x), S));
  X = __fadd_rn(e.z, __fmul_rn(__fmul_rn(dt, f.y), S));
}

// One event deposited into the images of one sample: img = [n_img][H*W] 64-bit fixed point (count+, count-[, ts+, ts-]).
__device__ __forceinline__ void splat_event(const float4 e, const float2 f, const float2 pm, int64_t* __restrict__ img, int H, int W,
                                            float tref, float S, int n_img, float tsw, int round_idx) {
  float Y, X, dt;
  warp_event(e, f, tref, S, Y, X, dt);
  const size_t hw = (size_t)H * W;
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

// d max(0, 1 - |d|) / d d with torch's tie rules: abs'(0) = 0; max(0, a) at a == 0 passes 1/2.
__device__ __forceinline__ float dtent(float d) {
  const float a = __fsub_rn(1.0f, fabsf(d));
  const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
  if (a > 0.f) return -sgn;
  if (a == 0.f) return -0.5f * sgn;
  return 0.f;
}

// d (sum over images of g_img * image) / d (per-event flow (fy, fx)) of one event: the adjoint of the bilinear splat.
__device__ __forceinline__ float2 splat_event_grad(const float4 e, const float2 f, const float2 pm, const float* __restrict__ g, int H,
                                                   int W, float tref, float S, int n_img, float tsw) {
  float Y, X, dt;
  warp_event(e, f, tref, S, Y, X, dt);
  const size_t hw = (size_t)H * W;
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
  return make_float2(gY * k, gX * k);
}

}  // namespace snnflow

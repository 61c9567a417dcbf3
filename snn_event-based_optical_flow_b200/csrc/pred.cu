// Flow prediction head: flow = tanh(conv1x1(x) + b), forward and backward.
// Reference: models/submodules.py:96-113 (ConvLayer, kernel_size 1, activation tanh) as used by
// LIFFireNet.pred (models/model.py:105-107,182).  HBM-bound: reads C planes, writes 2.
#include "common.cuh"

namespace snnflow {

constexpr int PR_THREADS = 256;
constexpr int PR_MAX_C = 64;

__global__ void __launch_bounds__(PR_THREADS) pred_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ flow,
                                                              int C, int HW) {
  __shared__ float sw[2 * PR_MAX_C + 2];
  for (int i = threadIdx.x; i < 2 * C; i += PR_THREADS) sw[i] = w[i];
  if (threadIdx.x < 2) sw[2 * C + threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int p = blockIdx.x * PR_THREADS + threadIdx.x;
  if (p >= HW) return;
  const float* xp = x + (size_t)b * C * HW + p;
  float a0 = 0.f, a1 = 0.f;
  for (int c = 0; c < C; ++c) {
    float v = xp[(size_t)c * HW];
    a0 = fmaf(v, sw[c], a0);
    a1 = fmaf(v, sw[C + c], a1);
  }
  flow[((size_t)b * 2 + 0) * HW + p] = tanhf(a0 + sw[2 * C]);
  flow[((size_t)b * 2 + 1) * HW + p] = tanhf(a1 + sw[2 * C + 1]);
}

// g_pre = g_flow * (1 - flow^2); g_x[c] = sum_o g_pre[o] * w[o][c]; dw[o][c] += sum g_pre[o]*x[c]; db[o] += sum g_pre[o]
// per-CTA partials [n_cta][2*C + 2] reduced in fixed order by pred_reduce_kernel.
__global__ void __launch_bounds__(PR_THREADS) pred_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ flow,
                                                              const float* __restrict__ g_flow, float* __restrict__ g_x,
                                                              float* __restrict__ part, int C, int HW) {
  __shared__ float sw[2 * PR_MAX_C];
  __shared__ float red[PR_THREADS / 32][2];
  for (int i = threadIdx.x; i < 2 * C; i += PR_THREADS) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int p = blockIdx.x * PR_THREADS + threadIdx.x;
  const bool ok = p < HW;
  float g0 = 0.f, g1 = 0.f;
  if (ok) {
    float f0 = flow[((size_t)b * 2 + 0) * HW + p], f1 = flow[((size_t)b * 2 + 1) * HW + p];
    g0 = g_flow[((size_t)b * 2 + 0) * HW + p] * (1.0f - f0 * f0);
    g1 = g_flow[((size_t)b * 2 + 1) * HW + p] * (1.0f - f1 * f1);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* mypart = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (2 * C + 2);
  const float* xp = x + (size_t)b * C * HW + p;
  float* gxp = g_x + (size_t)b * C * HW + p;
  for (int c = 0; c < C; ++c) {
    float xv = ok ? xp[(size_t)c * HW] : 0.f;
    if (ok) gxp[(size_t)c * HW] = g0 * sw[c] + g1 * sw[C + c];
    float s0 = warp_sum(g0 * xv), s1 = warp_sum(g1 * xv);
    if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t0 = 0.f, t1 = 0.f;
      for (int k = 0; k < PR_THREADS / 32; ++k) { t0 += red[k][0]; t1 += red[k][1]; }
      mypart[c] = t0; mypart[C + c] = t1;
    }
    __syncthreads();
  }
  float s0 = warp_sum(g0), s1 = warp_sum(g1);
  if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int k = 0; k < PR_THREADS / 32; ++k) { t0 += red[k][0]; t1 += red[k][1]; }
    mypart[2 * C] = t0; mypart[2 * C + 1] = t1;
  }
}

__global__ void pred_reduce_kernel(const float* __restrict__ part, float* dw, float* db, int C, int n_part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 2 * C + 2;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += part[(size_t)p * n + i];
  if (i < 2 * C) { if (dw) dw[i] += s; }
  else if (db) db[i - 2 * C] += s;
}

}  // namespace snnflow
using namespace snnflow;

extern "C" int snnflow_pred_fwd(const float* x, const float* w, const float* b, float* flow, int B, int C, int H,
                                int W, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(x && w && flow, "null pointer");
  SNNFLOW_REQUIRE(B > 0 && C > 0 && C <= PR_MAX_C && H > 0 && W > 0, "bad dims (C <= 64)");
  const int HW = H * W;
  prof_begin("pred_fwd", (cudaStream_t)stream, 4.0 * B * HW * (C + 2), 4.0 * B * HW * C);
  pred_fwd_kernel<<<dim3(ceil_div(HW, PR_THREADS), B), PR_THREADS, 0, (cudaStream_t)stream>>>(x, w, b, flow, C, HW);
  return check_launch("pred_fwd_kernel");
}

extern "C" size_t snnflow_pred_bwd_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return align_up((size_t)B * ceil_div(H * W, PR_THREADS) * (2 * C + 2) * sizeof(float), 256);
}

extern "C" int snnflow_pred_bwd(const float* x, const float* w, const float* flow, const float* g_flow, float* g_x,
                                float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int C, int H,
                                int W, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(x && w && flow && g_flow && g_x && workspace, "null pointer");
  SNNFLOW_REQUIRE(B > 0 && C > 0 && C <= PR_MAX_C && H > 0 && W > 0, "bad dims (C <= 64)");
  if (workspace_bytes < snnflow_pred_bwd_workspace_bytes(B, C, H, W)) {
    set_error("snnflow_pred_bwd: workspace too small");
    return SNNFLOW_EWORKSPACE;
  }
  const int HW = H * W, gx = ceil_div(HW, PR_THREADS);
  float* part = (float*)workspace;
  prof_begin("pred_bwd", (cudaStream_t)stream, 4.0 * B * HW * (2 * C + 4), 8.0 * B * HW * C);
  pred_bwd_kernel<<<dim3(gx, B), PR_THREADS, 0, (cudaStream_t)stream>>>(x, w, flow, g_flow, g_x, part, C, HW);
  int rc = check_launch("pred_bwd_kernel");
  if (rc) return rc;
  prof_begin("pred_reduce", (cudaStream_t)stream, 4.0 * gx * B * (2 * C + 2));
  pred_reduce_kernel<<<ceil_div(2 * C + 2, 128), 128, 0, (cudaStream_t)stream>>>(part, dw, db, C, gx * B);
  return check_launch("pred_reduce_kernel");
}

// Image of warped events: per-event flow gather, warp + bilinear/rounded splat, and their adjoints.
// Reference: utils/iwe.py:4-93 (purge_unfeasible, get_interpolation, interpolate), the gather of
// loss/flow.py:66-81 and compute_pol_iwe (utils/iwe.py:133-154).
// One thread per event: 16 B event + 8 B flow + 8 B polarity mask streamed from HBM, up to 8 L2
// atomics into small images.  The forward accumulates in 64-bit fixed point (2^-32), so the result is
// independent of the atomic order (deterministic) and equal to the correctly rounded fp32 sum up to
// 2^-33 per term.
#include "iwe.cuh"

namespace snnflow {

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
  const float tsw = ts_mode == 1 ? e.x : (ts_mode == 2 ? __fsub_rn(ts_ref, e.x) : 0.f);
  splat_event(e, f, pm, acc + (size_t)b * n_img * ((size_t)H * W), H, W, tref, S, n_img, tsw, round_idx);
}

__global__ void __launch_bounds__(IW_THREADS) iwe_fix_to_float_kernel(const int64_t* __restrict__ acc,
                                                                      float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * IW_THREADS + threadIdx.x;
  if (i < n) out[i] = (float)((double)acc[i] * IW_FIX_INV);
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
  const float tsw = ts_mode == 1 ? e.x : (ts_mode == 2 ? __fsub_rn(ts_ref, e.x) : 0.f);
  g_ev_flow[(size_t)b * N + n] = splat_event_grad(e, f, pm, g_img + (size_t)b * n_img * ((size_t)H * W), H, W, tref, S, n_img, tsw);
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

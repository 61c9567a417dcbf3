// Contrast-maximisation loss of one window and its gradient w.r.t. the flow maps, in one host call.
//
// Replaces T calls of EventWarping.event_flow_association (loss/flow.py:58-121), EventWarping.forward (:178-303) and
// the autograd graph torch builds behind them (about 250 small kernels and 8 host synchronisations per optimizer step
// in the reference) by nine launches:
//   gather      per-event flow lookup of all T bins, events concatenated per sample with the reference's
//               timestamp shift ts += bin index (:77-92)
//   splat x2    image of warped events at t_ref = T (forward) and 0 (backward): 4 images each (iwe.cu)
//   sums        per sample and direction: S = sum(ts_p^2 + ts_n^2), Z = number of pixels with events (:214-228)
//   g_img       d loss / d images
//   splat^T x2  d loss / d per-event flow (iwe.cu, torch's tie rules)
//   smooth      Charbonnier smoothness over dx, dy, both diagonals and dt (:264-296): value + gradient, which
//               also INITIALISES the flow-map gradient
//   scatter     per-event gradients added onto the flow maps (adjoint of the gather)
//   final       fixed-order sum of all partial sums -> loss
// All reductions run in a fixed order (deterministic) except the scatter's fp32 atomics, as in flow_gather_bwd.
#include "common.cuh"

namespace snnflow {

constexpr int WL_THREADS = 256;

__device__ __forceinline__ long long wl_flat_index(float y, float x, int W) {   // loss/flow.py:67-69: fp32 y*W + x, truncated
  return (long long)__fadd_rn(__fmul_rn(y, (float)W), x);
}

// events [T,B,N,4] -> ev_cat [B,T*N,4] (ts + t), pm_cat [B,T*N,2], ev_flow [B,T*N,2] = (fy, fx) at the event's pixel
__global__ void __launch_bounds__(WL_THREADS) wl_gather_kernel(const float* __restrict__ flow, const float4* __restrict__ events,
                                                               const float2* __restrict__ pol, float4* __restrict__ ev_cat,
                                                               float2* __restrict__ pm_cat, float2* __restrict__ ev_flow, int T,
                                                               int B, int64_t N, int H, int W) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * WL_THREADS + threadIdx.x;
  if (i >= (int64_t)T * N) return;
  const int t = (int)(i / N);
  const int64_t n = i - (int64_t)t * N;
  const size_t src = ((size_t)t * B + b) * N + n;
  float4 e = __ldg(events + src);
  if (t > 0) e.x = __fadd_rn(e.x, (float)t);   // event_list[:, :, 0:1] += self._passes  (:91)
  const size_t hw = (size_t)H * W;
  const long long idx = wl_flat_index(e.y, e.z, W);
  float2 f = make_float2(0.f, 0.f);
  if (idx >= 0 && idx < (long long)hw) {
    const float* fl = flow + ((size_t)t * B + b) * 2 * hw;
    f.x = __ldg(fl + hw + idx);   // vertical component (channel 1)
    f.y = __ldg(fl + idx);        // horizontal component (channel 0)
  }
  const size_t dst = (size_t)b * T * N + i;
  ev_cat[dst] = e;
  pm_cat[dst] = __ldg(pol + src);
  ev_flow[dst] = f;
}

// per-block partial sums of S and Z for (direction, sample); img [2][B][4][H*W]
__global__ void __launch_bounds__(WL_THREADS) wl_sums_kernel(const float* __restrict__ img, float* __restrict__ part, int B, int HW,
                                                             float max_ts) {
  const int b = blockIdx.y, dir = blockIdx.z;
  const float* im = img + ((size_t)dir * B + b) * 4 * HW;
  float s = 0.f, z = 0.f;
  for (int p = blockIdx.x * WL_THREADS + threadIdx.x; p < HW; p += gridDim.x * WL_THREADS) {
    const float cp = im[p], cn = im[HW + p];
    const float tp = im[2 * HW + p] / (cp + 1e-9f) / max_ts, tn = im[3 * HW + p] / (cn + 1e-9f) / max_ts;   // :214-217
    s += tp * tp + tn * tn;
    const float tot = cp + cn;
    z += tot > 0.f ? 1.f : tot;                                                                             // :224-227
  }
  __shared__ float red[2][WL_THREADS / 32];
  s = warp_sum(s);
  z = warp_sum(z);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = z; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < WL_THREADS / 32; ++w) { a += red[0][w]; c += red[1][w]; }
    float* o = part + (((size_t)dir * B + b) * gridDim.x + blockIdx.x) * 2;
    o[0] = a; o[1] = c;
  }
}

// sz[dir][b] = (S, Z): fixed-order sum of the block partials
__global__ void wl_sz_kernel(const float* __restrict__ part, float* __restrict__ sz, int n_db, int n_blk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_db) return;
  float s = 0.f, z = 0.f;
  for (int k = 0; k < n_blk; ++k) { s += part[((size_t)i * n_blk + k) * 2]; z += part[((size_t)i * n_blk + k) * 2 + 1]; }
  sz[2 * i] = s; sz[2 * i + 1] = z;
}

// d loss / d images: L = S / Z (or S without loss_scaling)
__global__ void __launch_bounds__(WL_THREADS) wl_gimg_kernel(const float* __restrict__ img, const float* __restrict__ sz,
                                                             float* __restrict__ g_img, int B, int HW, float max_ts, int loss_scaling) {
  const int b = blockIdx.y, dir = blockIdx.z;
  const int p = blockIdx.x * WL_THREADS + threadIdx.x;
  if (p >= HW) return;
  const size_t base = ((size_t)dir * B + b) * 4 * HW;
  const float S = sz[2 * (dir * B + b)], Z = loss_scaling ? sz[2 * (dir * B + b) + 1] : 1.f;
  const float cp = img[base + p], cn = img[base + HW + p];
  const float ap = cp + 1e-9f, an = cn + 1e-9f;
  const float tp = img[base + 2 * HW + p] / ap / max_ts, tn = img[base + 3 * HW + p] / an / max_ts;
  const float gz = (loss_scaling && !(cp + cn > 0.f)) ? -S / (Z * Z) : 0.f;   // the where(tot > 0, 1, tot) path
  g_img[base + p] = -2.f * tp * tp / (ap * Z) + gz;
  g_img[base + HW + p] = -2.f * tn * tn / (an * Z) + gz;
  g_img[base + 2 * HW + p] = 2.f * tp / (ap * max_ts * Z);
  g_img[base + 3 * HW + p] = 2.f * tn / (an * max_ts * Z);
}

// Charbonnier smoothness of the flow maps [T,B,2,H,W]: every pixel computes the (up to) ten pair terms it belongs to,
// sums the values of the five pairs it "owns" (it is their first pixel) and writes its gradient
//   g_fx = g_fy = scale * sum over pairs  +-(d / sqrt(d^2 + 1e-6)) [* mask_p * mask_q],  d = (fx_p - fx_q) + (fy_p - fy_q)
__global__ void __launch_bounds__(WL_THREADS) wl_smooth_kernel(const float* __restrict__ flow, const float* __restrict__ mask,
                                                               float* __restrict__ g_flow, float* __restrict__ part, int T, int B,
                                                               int H, int W, float scale) {
  const int HW = H * W;
  const int p = blockIdx.x * WL_THREADS + threadIdx.x;
  const int tb = blockIdx.y;   // t * B + b
  const int t = tb / B;
  float val = 0.f;
  if (p < HW) {
    const int y = p / W, x = p - y * W;
    const float* fx = flow + (size_t)tb * 2 * HW;
    const float* fy = fx + HW;
    const float* m = mask ? mask + (size_t)tb * HW : nullptr;
    const float s_p = fx[p] + fy[p];
    const float m_p = m ? m[p] : 1.f;
    float g = 0.f;
    // pair(p, q): returns d/charb * mq and accumulates the value when p owns the pair
    auto pair = [&](const float* qx, const float* qy, const float* qm, int q, bool own, float sign) {
      const float d = sign * (s_p - (qx[q] + qy[q]));          // first-minus-second of the pair
      const float mm = m_p * (qm ? qm[q] : 1.f);
      const float c = sqrtf(d * d + 1e-6f);
      if (own) val += c * mm;
      g += sign * (d / c) * mm;
    };
    if (x + 1 < W) pair(fx, fy, m, p + 1, true, 1.f);                               // dx
    if (x > 0) pair(fx, fy, m, p - 1, false, -1.f);
    if (y + 1 < H) pair(fx, fy, m, p + W, true, 1.f);                               // dy
    if (y > 0) pair(fx, fy, m, p - W, false, -1.f);
    if (y + 1 < H && x + 1 < W) pair(fx, fy, m, p + W + 1, true, 1.f);              // (y,x) - (y+1,x+1)
    if (y > 0 && x > 0) pair(fx, fy, m, p - W - 1, false, -1.f);
    if (y > 0 && x + 1 < W) pair(fx, fy, m, p - W + 1, true, 1.f);                  // (y,x) - (y-1,x+1): fx[1:, :-1] - fx[:-1, 1:]
    if (y + 1 < H && x > 0) pair(fx, fy, m, p + W - 1, false, -1.f);
    if (t + 1 < T) {                                                                // dt: frame t - frame t+1
      const float* nx = fx + (size_t)B * 2 * HW;
      pair(nx, nx + HW, m ? m + (size_t)B * HW : nullptr, p, true, 1.f);
    }
    if (t > 0) {
      const float* px = fx - (size_t)B * 2 * HW;
      pair(px, px + HW, m ? m - (size_t)B * HW : nullptr, p, false, -1.f);
    }
    g_flow[(size_t)tb * 2 * HW + p] = scale * g;
    g_flow[(size_t)tb * 2 * HW + HW + p] = scale * g;
  }
  __shared__ float red[WL_THREADS / 32];
  val = warp_sum(val);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < WL_THREADS / 32; ++w) a += red[w];
    part[(size_t)tb * gridDim.x + blockIdx.x] = a;
  }
}

// adjoint of the gather: g_flow[t][b][{1,0}][pixel] += (g_fw + g_bw)[b][i]
__global__ void __launch_bounds__(WL_THREADS) wl_scatter_kernel(const float4* __restrict__ ev_cat, const float2* __restrict__ g_fw,
                                                                const float2* __restrict__ g_bw, float* __restrict__ g_flow, int T,
                                                                int B, int64_t N, int H, int W) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * WL_THREADS + threadIdx.x;
  if (i >= (int64_t)T * N) return;
  const int t = (int)(i / N);
  const size_t src = (size_t)b * T * N + i;
  const float4 e = __ldg(ev_cat + src);
  const size_t hw = (size_t)H * W;
  const long long idx = wl_flat_index(e.y, e.z, W);
  if (idx < 0 || idx >= (long long)hw) return;
  const float2 a = g_fw[src], c = g_bw[src];
  const float gy = a.x + c.x, gx = a.y + c.y;
  float* g = g_flow + ((size_t)t * B + b) * 2 * hw;
  if (gy != 0.f) atomicAdd(g + hw + idx, gy);
  if (gx != 0.f) atomicAdd(g + idx, gx);
}

// loss = sum_b S_fw/Z_fw + sum_b S_bw/Z_bw + weight * smooth_sum / (5 T)
__global__ void wl_final_kernel(const float* __restrict__ sz, const float* __restrict__ spart, int n_spart, float* __restrict__ loss,
                                int B, int loss_scaling, float smooth_scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_spart; i += blockDim.x) s += spart[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sm = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sm += red[w];
    float l = 0.f;
    for (int i = 0; i < 2 * B; ++i) l += loss_scaling ? sz[2 * i] / sz[2 * i + 1] : sz[2 * i];
    loss[0] = l + smooth_scale * sm;
  }
}

struct WlLayout {
  size_t ev_cat, pm_cat, ev_flow, img, scratch, g_img, g_fw, g_bw, part, sz, spart, total;
  int n_blk, n_sblk;
};
static WlLayout wl_layout(int T, int B, int64_t N, int H, int W) {
  WlLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  const size_t ne = (size_t)B * T * N, hw = (size_t)H * W;
  L.ev_cat = take(ne * 16);
  L.pm_cat = take(ne * 8);
  L.ev_flow = take(ne * 8);
  L.img = take(2 * (size_t)B * 4 * hw * 4);
  L.scratch = take((size_t)B * 4 * hw * 8);
  L.g_img = take(2 * (size_t)B * 4 * hw * 4);
  L.g_fw = take(ne * 8);
  L.g_bw = take(ne * 8);
  L.n_blk = ceil_div((int)hw, WL_THREADS * 4);
  L.part = take((size_t)2 * B * L.n_blk * 2 * 4);
  L.sz = take((size_t)2 * B * 2 * 4);
  L.n_sblk = ceil_div((int)hw, WL_THREADS);
  L.spart = take((size_t)T * B * L.n_sblk * 4);
  L.total = o;
  return L;
}

}  // namespace snnflow
using namespace snnflow;

extern "C" size_t snnflow_window_loss_workspace_bytes(int T, int B, int64_t N, int H, int W) {
  if (T <= 0 || B <= 0 || N < 0 || H <= 0 || W <= 0) return 0;
  return wl_layout(T, B, N, H, W).total;
}

extern "C" int snnflow_window_loss(const float* flow, const float* events, const float* pol_mask, const float* event_mask,
                                   float* loss, float* g_flow, void* workspace, size_t workspace_bytes, int T, int B, int64_t N,
                                   int H, int W, float flow_scaling, float regul_weight, int loss_scaling,
                                   snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(flow && events && pol_mask && loss && g_flow && workspace, "null pointer");
  SNNFLOW_REQUIRE(T > 0 && B > 0 && N > 0 && H > 0 && W > 0 && (int64_t)H * W < (1 << 24), "bad dims");
  SNNFLOW_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)events & 15) == 0 && ((uintptr_t)pol_mask & 7) == 0, "misaligned");
  const WlLayout L = wl_layout(T, B, N, H, W);
  if (workspace_bytes < L.total) {
    set_error("snnflow_window_loss: workspace %zu < %zu", workspace_bytes, L.total);
    return SNNFLOW_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float4* ev_cat = (float4*)(ws + L.ev_cat);
  float2* pm_cat = (float2*)(ws + L.pm_cat);
  float2* ev_flow = (float2*)(ws + L.ev_flow);
  float* img = (float*)(ws + L.img);
  int64_t* scratch = (int64_t*)(ws + L.scratch);
  float* g_img = (float*)(ws + L.g_img);
  float2* g_fw = (float2*)(ws + L.g_fw);
  float2* g_bw = (float2*)(ws + L.g_bw);
  float* part = (float*)(ws + L.part);
  float* sz = (float*)(ws + L.sz);
  float* spart = (float*)(ws + L.spart);
  const int64_t TN = (int64_t)T * N;
  const int HW = H * W;
  const float max_ts = (float)T;
  const dim3 egrid((unsigned)ceil_div64(TN, WL_THREADS), B);

  prof_begin("loss_gather", st, 56.0 * TN * B);
  wl_gather_kernel<<<egrid, WL_THREADS, 0, st>>>(flow, (const float4*)events, (const float2*)pol_mask, ev_cat, pm_cat, ev_flow, T, B,
                                                 N, H, W);
  int rc = check_launch("wl_gather_kernel");
  if (rc) return rc;
  // forward-warped (t_ref = T, weight ts) and backward-warped (t_ref = 0, weight T - ts) images      loss/flow.py:199-246
  rc = snnflow_iwe_splat_fwd((const float*)ev_cat, (const float*)ev_flow, (const float*)pm_cat, img, scratch, B, TN, H, W, max_ts,
                             flow_scaling, 4, 1, max_ts, 0, stream);
  if (rc) return rc;
  rc = snnflow_iwe_splat_fwd((const float*)ev_cat, (const float*)ev_flow, (const float*)pm_cat, img + (size_t)B * 4 * HW, scratch, B,
                             TN, H, W, 0.f, flow_scaling, 4, 2, max_ts, 0, stream);
  if (rc) return rc;
  prof_begin("loss_sums", st, 32.0 * B * HW);
  wl_sums_kernel<<<dim3(L.n_blk, B, 2), WL_THREADS, 0, st>>>(img, part, B, HW, max_ts);
  rc = check_launch("wl_sums_kernel");
  if (rc) return rc;
  prof_begin("loss_sz", st, 0.0);
  wl_sz_kernel<<<ceil_div(2 * B, 64), 64, 0, st>>>(part, sz, 2 * B, L.n_blk);
  rc = check_launch("wl_sz_kernel");
  if (rc) return rc;
  prof_begin("loss_gimg", st, 64.0 * B * HW);
  wl_gimg_kernel<<<dim3(ceil_div(HW, WL_THREADS), B, 2), WL_THREADS, 0, st>>>(img, sz, g_img, B, HW, max_ts, loss_scaling);
  rc = check_launch("wl_gimg_kernel");
  if (rc) return rc;
  rc = snnflow_iwe_splat_bwd((const float*)ev_cat, (const float*)ev_flow, (const float*)pm_cat, g_img, (float*)g_fw, B, TN, H, W,
                             max_ts, flow_scaling, 4, 1, max_ts, stream);
  if (rc) return rc;
  rc = snnflow_iwe_splat_bwd((const float*)ev_cat, (const float*)ev_flow, (const float*)pm_cat, g_img + (size_t)B * 4 * HW,
                             (float*)g_bw, B, TN, H, W, 0.f, flow_scaling, 4, 2, max_ts, stream);
  if (rc) return rc;
  const float smooth_scale = regul_weight / (5.0f * (float)T);
  prof_begin("loss_smooth", st, 16.0 * T * B * HW);
  wl_smooth_kernel<<<dim3(L.n_sblk, T * B), WL_THREADS, 0, st>>>(flow, event_mask, g_flow, spart, T, B, H, W, smooth_scale);
  rc = check_launch("wl_smooth_kernel");
  if (rc) return rc;
  prof_begin("loss_scatter", st, 40.0 * TN * B);
  wl_scatter_kernel<<<egrid, WL_THREADS, 0, st>>>(ev_cat, g_fw, g_bw, g_flow, T, B, N, H, W);
  rc = check_launch("wl_scatter_kernel");
  if (rc) return rc;
  prof_begin("loss_final", st, 0.0);
  wl_final_kernel<<<1, 256, 0, st>>>(sz, spart, T * B * L.n_sblk, loss, B, loss_scaling, smooth_scale);
  return check_launch("wl_final_kernel");
}

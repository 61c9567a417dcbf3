// Contrast-maximisation loss of one window and its gradient w.r.t. the flow maps, in one host call.
//
// Replaces T calls of EventWarping.event_flow_association (loss/flow.py:58-121), EventWarping.forward (:178-303) and
// the autograd graph torch builds behind them (about 250 small kernels and 8 host synchronisations per optimizer step
// in the reference) by six launches and one memset (14 launches and two memsets before the kernels were merged):
//   gather+splat  per-event flow lookup of all T bins (events of a sample concatenated with the reference's timestamp
//                 shift ts += bin index, :77-92) and BOTH images of warped events - t_ref = T (forward) and 0 (backward),
//                 4 images each - accumulated in 64-bit fixed point (iwe.cuh)
//   sums          fixed point -> fp32 images; per sample and direction S = sum(ts_p^2 + ts_n^2), Z = number of pixels
//                 with events (:214-228), as block partials
//   g_img         d loss / d images (every block folds the partials itself, in a fixed order)
//   smooth        Charbonnier smoothness over dx, dy, both diagonals and dt (:264-296): value + gradient, which
//                 also INITIALISES the flow-map gradient
//   splat^T       d loss / d per-event flow of both directions (iwe.cuh, torch's tie rules), added onto the flow-map
//                 gradient at the event's pixel (adjoint of the gather)
//   final         fixed-order sum of all partial sums -> loss
// All reductions run in a fixed order (deterministic) except the fp32 atomics of the last scatter, as in flow_gather_bwd.
#include "iwe.cuh"

namespace snnflow {

constexpr int WL_THREADS = 256;

// one event of the window: (ts + bin index, y, x, p), its polarity mask, the flow (fy, fx) at its pixel in ITS bin's map
struct WlEvent {
  float4 e;
  float2 f, pm;
  long long idx;   // flat pixel index (loss/flow.py:67-69: fp32 y*W + x, truncated); outside [0, H*W): no flow, no gradient
};

__device__ __forceinline__ WlEvent wl_load_event(const float* __restrict__ flow, const float4* __restrict__ events,
                                                 const float2* __restrict__ pol, int t, int b, int64_t n, int B, int64_t N, int H,
                                                 int W) {
  WlEvent ev;
  const size_t src = ((size_t)t * B + b) * N + n;
  ev.e = __ldg(events + src);
  if (t > 0) ev.e.x = __fadd_rn(ev.e.x, (float)t);   // event_list[:, :, 0:1] += self._passes  (:91)
  const size_t hw = (size_t)H * W;
  ev.idx = flat_index(ev.e.y, ev.e.z, W);
  ev.f = make_float2(0.f, 0.f);
  if (ev.idx >= 0 && ev.idx < (long long)hw) {
    const float* fl = flow + ((size_t)t * B + b) * 2 * hw;
    ev.f.x = __ldg(fl + hw + ev.idx);   // vertical component (channel 1)
    ev.f.y = __ldg(fl + ev.idx);        // horizontal component (channel 0)
  }
  ev.pm = __ldg(pol + src);
  return ev;
}

// events [T,B,N,4] -> acc [2 directions][B][4][H*W] 64-bit fixed point (zeroed by the caller)
__global__ void __launch_bounds__(WL_THREADS) wl_gather_splat_kernel(const float* __restrict__ flow, const float4* __restrict__ events,
                                                                     const float2* __restrict__ pol, int64_t* __restrict__ acc, int T,
                                                                     int B, int64_t N, int H, int W, float max_ts, float S) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * WL_THREADS + threadIdx.x;
  if (i >= (int64_t)T * N) return;
  const int t = (int)(i / N);
  const WlEvent ev = wl_load_event(flow, events, pol, t, b, i - (int64_t)t * N, B, N, H, W);
  const size_t hw = (size_t)H * W;
  // forward-warped (t_ref = T, weight ts) and backward-warped (t_ref = 0, weight T - ts) images      loss/flow.py:199-246
  splat_event(ev.e, ev.f, ev.pm, acc + (size_t)b * 4 * hw, H, W, max_ts, S, 4, ev.e.x, 0);
  splat_event(ev.e, ev.f, ev.pm, acc + ((size_t)B + b) * 4 * hw, H, W, 0.f, S, 4, __fsub_rn(max_ts, ev.e.x), 0);
}

// fixed point -> fp32 images img [2][B][4][H*W], and per-block partial sums of S and Z for (direction, sample)
__global__ void __launch_bounds__(WL_THREADS) wl_sums_kernel(const int64_t* __restrict__ acc, float* __restrict__ img,
                                                             float* __restrict__ part, int B, int HW, float max_ts) {
  const int b = blockIdx.y, dir = blockIdx.z;
  const size_t base = ((size_t)dir * B + b) * 4 * HW;
  const int64_t* ac = acc + base;
  float* im = img + base;
  float s = 0.f, z = 0.f;
  for (int p = blockIdx.x * WL_THREADS + threadIdx.x; p < HW; p += gridDim.x * WL_THREADS) {
    const float cp = (float)((double)ac[p] * IW_FIX_INV), cn = (float)((double)ac[HW + p] * IW_FIX_INV);
    const float sp = (float)((double)ac[2 * HW + p] * IW_FIX_INV), sn = (float)((double)ac[3 * HW + p] * IW_FIX_INV);
    im[p] = cp; im[HW + p] = cn; im[2 * HW + p] = sp; im[3 * HW + p] = sn;
    const float tp = sp / (cp + 1e-9f) / max_ts, tn = sn / (cn + 1e-9f) / max_ts;   // :214-217
    s += tp * tp + tn * tn;
    const float tot = cp + cn;
    z += tot > 0.f ? 1.f : tot;                                                    // :224-227
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

// (S, Z) of one (direction, sample): fixed-order sum of the block partials - every consumer folds them itself
__device__ __forceinline__ float2 wl_fold(const float* __restrict__ part, int db, int n_blk) {
  float s = 0.f, z = 0.f;
  for (int k = 0; k < n_blk; ++k) { s += part[((size_t)db * n_blk + k) * 2]; z += part[((size_t)db * n_blk + k) * 2 + 1]; }
  return make_float2(s, z);
}

// d loss / d images: L = S / Z (or S without loss_scaling)
__global__ void __launch_bounds__(WL_THREADS) wl_gimg_kernel(const float* __restrict__ img, const float* __restrict__ part, int n_blk,
                                                             float* __restrict__ g_img, int B, int HW, float max_ts, int loss_scaling) {
  const int b = blockIdx.y, dir = blockIdx.z;
  __shared__ float2 sz_s;
  if (threadIdx.x == 0) sz_s = wl_fold(part, dir * B + b, n_blk);
  __syncthreads();
  const int p = blockIdx.x * WL_THREADS + threadIdx.x;
  if (p >= HW) return;
  const size_t base = ((size_t)dir * B + b) * 4 * HW;
  const float S = sz_s.x, Z = loss_scaling ? sz_s.y : 1.f;
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
      // sqrt(q2) = q2 * rsqrt(q2), d / sqrt(q2) = d * rsqrt(q2): one MUFU.RSQ (2 ulp) instead of an IEEE sqrt and a division
      // per pair - the kernel was instruction bound on those (37 -> 12 us at 10 x 8 x 128 x 128)
      const float q2 = fmaf(d, d, 1e-6f), r = rsqrtf(q2);
      if (own) val += (q2 * r) * mm;
      g += sign * (d * r) * mm;
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

// adjoint of the splats and of the gather: g_flow[t][b][{1,0}][pixel] += d loss / d (fy, fx) of the event, both directions
__global__ void __launch_bounds__(WL_THREADS) wl_splat_bwd_kernel(const float* __restrict__ flow, const float4* __restrict__ events,
                                                                  const float2* __restrict__ pol, const float* __restrict__ g_img,
                                                                  float* __restrict__ g_flow, int T, int B, int64_t N, int H, int W,
                                                                  float max_ts, float S) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * WL_THREADS + threadIdx.x;
  if (i >= (int64_t)T * N) return;
  const int t = (int)(i / N);
  const WlEvent ev = wl_load_event(flow, events, pol, t, b, i - (int64_t)t * N, B, N, H, W);
  const size_t hw = (size_t)H * W;
  if (ev.idx < 0 || ev.idx >= (long long)hw) return;   // no flow-map pixel receives this event's gradient
  const float2 a = splat_event_grad(ev.e, ev.f, ev.pm, g_img + (size_t)b * 4 * hw, H, W, max_ts, S, 4, ev.e.x);
  const float2 c = splat_event_grad(ev.e, ev.f, ev.pm, g_img + ((size_t)B + b) * 4 * hw, H, W, 0.f, S, 4, __fsub_rn(max_ts, ev.e.x));
  const float gy = a.x + c.x, gx = a.y + c.y;
  float* g = g_flow + ((size_t)t * B + b) * 2 * hw;
  if (gy != 0.f) atomicAdd(g + hw + ev.idx, gy);
  if (gx != 0.f) atomicAdd(g + ev.idx, gx);
}

// loss = sum_b S_fw/Z_fw + sum_b S_bw/Z_bw + weight * smooth_sum / (5 T)
__global__ void __launch_bounds__(256) wl_final_kernel(const float* __restrict__ part, int n_blk, const float* __restrict__ spart,
                                                       int n_spart, float* __restrict__ loss, int B, int loss_scaling,
                                                       float smooth_scale) {
  __shared__ float red[8];
  __shared__ float2 sz_s[64];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_spart; i += 256) s += spart[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  float l = 0.f;                                   // thread 0's running sum over (direction, sample), in index order
  for (int i0 = 0; i0 < 2 * B; i0 += 64) {
    __syncthreads();
    if (threadIdx.x < 64 && i0 + (int)threadIdx.x < 2 * B) sz_s[threadIdx.x] = wl_fold(part, i0 + threadIdx.x, n_blk);
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < 64 && i0 + i < 2 * B; ++i) l += loss_scaling ? sz_s[i].x / sz_s[i].y : sz_s[i].x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sm = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sm += red[w];
    loss[0] = l + smooth_scale * sm;
  }
}

struct WlLayout {
  size_t acc, img, g_img, part, spart, total;
  int n_blk, n_sblk;
};
static WlLayout wl_layout(int T, int B, int H, int W) {
  WlLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  const size_t hw = (size_t)H * W;
  L.acc = take(2 * (size_t)B * 4 * hw * 8);
  L.img = take(2 * (size_t)B * 4 * hw * 4);
  L.g_img = take(2 * (size_t)B * 4 * hw * 4);
  L.n_blk = ceil_div((int)hw, WL_THREADS * 4);
  L.part = take((size_t)2 * B * L.n_blk * 2 * 4);
  L.n_sblk = ceil_div((int)hw, WL_THREADS);
  L.spart = take((size_t)T * B * L.n_sblk * 4);
  L.total = o;
  return L;
}

}  // namespace snnflow
using namespace snnflow;

extern "C" size_t snnflow_window_loss_workspace_bytes(int T, int B, int64_t N, int H, int W) {
  if (T <= 0 || B <= 0 || N < 0 || H <= 0 || W <= 0) return 0;
  return wl_layout(T, B, H, W).total;
}

extern "C" int snnflow_window_loss(const float* flow, const float* events, const float* pol_mask, const float* event_mask,
                                   float* loss, float* g_flow, void* workspace, size_t workspace_bytes, int T, int B, int64_t N,
                                   int H, int W, float flow_scaling, float regul_weight, int loss_scaling,
                                   snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(flow && events && pol_mask && loss && g_flow && workspace, "null pointer");
  SNNFLOW_REQUIRE(T > 0 && B > 0 && N > 0 && H > 0 && W > 0 && (int64_t)H * W < (1 << 24), "bad dims");
  SNNFLOW_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)events & 15) == 0 && ((uintptr_t)pol_mask & 7) == 0, "misaligned");
  const WlLayout L = wl_layout(T, B, H, W);
  if (workspace_bytes < L.total) {
    set_error("snnflow_window_loss: workspace %zu < %zu", workspace_bytes, L.total);
    return SNNFLOW_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int64_t* acc = (int64_t*)(ws + L.acc);
  float* img = (float*)(ws + L.img);
  float* g_img = (float*)(ws + L.g_img);
  float* part = (float*)(ws + L.part);
  float* spart = (float*)(ws + L.spart);
  const int64_t TN = (int64_t)T * N;
  const int HW = H * W;
  const float max_ts = (float)T;
  const dim3 egrid((unsigned)ceil_div64(TN, WL_THREADS), B);

  SNNFLOW_CUDA(cudaMemsetAsync(acc, 0, 2 * (size_t)B * 4 * HW * sizeof(int64_t), st));
  prof_begin("loss_gather_splat", st, 32.0 * TN * B + 64.0 * B * HW);
  wl_gather_splat_kernel<<<egrid, WL_THREADS, 0, st>>>(flow, (const float4*)events, (const float2*)pol_mask, acc, T, B, N, H, W, max_ts,
                                                       flow_scaling);
  int rc = check_launch("wl_gather_splat_kernel");
  if (rc) return rc;
  prof_begin("loss_sums", st, 96.0 * B * HW);
  wl_sums_kernel<<<dim3(L.n_blk, B, 2), WL_THREADS, 0, st>>>(acc, img, part, B, HW, max_ts);
  rc = check_launch("wl_sums_kernel");
  if (rc) return rc;
  prof_begin("loss_gimg", st, 64.0 * B * HW);
  wl_gimg_kernel<<<dim3(ceil_div(HW, WL_THREADS), B, 2), WL_THREADS, 0, st>>>(img, part, L.n_blk, g_img, B, HW, max_ts, loss_scaling);
  rc = check_launch("wl_gimg_kernel");
  if (rc) return rc;
  const float smooth_scale = regul_weight / (5.0f * (float)T);
  prof_begin("loss_smooth", st, 16.0 * T * B * HW);
  wl_smooth_kernel<<<dim3(L.n_sblk, T * B), WL_THREADS, 0, st>>>(flow, event_mask, g_flow, spart, T, B, H, W, smooth_scale);
  rc = check_launch("wl_smooth_kernel");
  if (rc) return rc;
  prof_begin("loss_splat_bwd", st, 40.0 * TN * B + 32.0 * B * HW);
  wl_splat_bwd_kernel<<<egrid, WL_THREADS, 0, st>>>(flow, (const float4*)events, (const float2*)pol_mask, g_img, g_flow, T, B, N, H, W,
                                                    max_ts, flow_scaling);
  rc = check_launch("wl_splat_bwd_kernel");
  if (rc) return rc;
  prof_begin("loss_final", st, 0.0);
  wl_final_kernel<<<1, 256, 0, st>>>(part, L.n_blk, spart, T * B * L.n_sblk, loss, B, loss_scaling, smooth_scale);
  return check_launch("wl_final_kernel");
}

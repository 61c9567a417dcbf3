// Loader window formatter: one raw window of events per batch slot -> the tensors of one training / eval batch
// (event_cnt, event_voxel, event_mask, event_list, event_list_pol_mask) in five launches.
//
// Reference (per batch item, ~40 torch-CPU ops on the training thread, num_workers = 0):
//   event_formatting            dataloader/base.py:71-99     fp32 cast, p*2-1, min-max normalised timestamps
//   augment_events              dataloader/base.py:101-126   horizontal / vertical / polarity flips
//   create_*_encoding           dataloader/base.py:160-235   counts, any-event mask, voxel grid, event list, polarity mask
//   create_hot_mask             dataloader/base.py:237-256 + get_hot_event_mask dataloader/encodings.py:88-103
//   hot-pixel application, average-pool down-sampling and event-list rescaling   dataloader/h5.py:323-331,375-410
//   custom_collate              dataloader/base.py:261-278   [B,N,4] / [B,N,2] lists
//
// HBM-bound streaming pass over the events (one coalesced read of x, y, t, p; one float4 + one float2 write per event)
// plus L2 atomics into sensor-sized accumulators that stay L2 resident.  Counts / masks are exact integers in fp32;
// the voxel grid goes through the 64-bit fixed-point accumulator of encode.cu (order independent, deterministic).
// Per-sample reductions (timestamp range) use order-preserving integer atomics, so nothing depends on scheduling.
#include "common.cuh"

namespace snnflow {

constexpr int LD_THREADS = 256;
constexpr int HOT_THREADS = 1024;
constexpr double LD_FIX_SCALE = 4294967296.0;       // 2^32
constexpr double LD_FIX_INV = 1.0 / 4294967296.0;

// float <-> unsigned with the same ordering (for atomicMin / atomicMax on timestamps)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <bool TS64>
__device__ __forceinline__ float load_ts(const void* ts, int64_t k, double t0) {
  // get_events subtracts t0 in float64 (h5.py:129), event_formatting casts to fp32 (base.py:88)
  if (TS64) return (float)(reinterpret_cast<const double*>(ts)[k] - t0);
  return reinterpret_cast<const float*>(ts)[k];
}

struct LoaderWs {          // workspace carve-up (device pointers)
  float* cnt;              // [B,2,H,W]
  int32_t* last;           // [B,H,W]   1 + index of the last event at the pixel
  int64_t* vox;            // [B,nb,H,W] fixed point
  float* hotmask;          // [B,H,W]
  unsigned* tminmax;       // [B,2] ordered-uint min, max
};

__global__ void __launch_bounds__(LD_THREADS) ld_init_kernel(unsigned* tminmax, int B) {
  const int i = blockIdx.x * LD_THREADS + threadIdx.x;
  if (i < B) { tminmax[2 * i] = 0xffffffffu; tminmax[2 * i + 1] = 0u; }
}

template <bool TS64>
__global__ void __launch_bounds__(LD_THREADS) ld_minmax_kernel(const void* __restrict__ ts, const double* __restrict__ t0,
                                                               unsigned* __restrict__ tminmax, int64_t N) {
  const int b = blockIdx.y;
  const double t0b = t0 ? t0[b] : 0.0;
  unsigned lo = 0xffffffffu, hi = 0u;
  const int64_t stride = (int64_t)gridDim.x * LD_THREADS;
  // four independent loads per thread and trip: the reduction is latency bound otherwise (ncu: 1.6 TB/s with one)
  for (int64_t k0 = (int64_t)blockIdx.x * LD_THREADS + threadIdx.x; k0 < N; k0 += 4 * stride) {
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + j * stride;
      t[j] = load_ts<TS64>(ts, (int64_t)b * N + (k < N ? k : k0), t0b);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned o = f2ord(t[j]);
      lo = min(lo, o); hi = max(hi, o);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(tminmax + 2 * b, lo);
    atomicMax(tminmax + 2 * b + 1, hi);
  }
}

struct LdScatterArgs {
  const float* xs; const float* ys; const void* ts; const float* ps; const double* t0; const int32_t* flips;
  LoaderWs ws;
  float* ev_list; float* ev_pol;
  int64_t N; int H, W, nb, round_ts;
  int rescale; float sy, sx, ymax, xmax;   // down-sampled event list (h5.py:401-405)
};

template <bool TS64>
__global__ void __launch_bounds__(LD_THREADS) ld_scatter_kernel(const LdScatterArgs a) {
  const int b = blockIdx.y;
  const int H = a.H, W = a.W;
  const size_t HW = (size_t)H * W;
  const double t0b = a.t0 ? a.t0[b] : 0.0;
  const float tmin = ord2f(a.ws.tminmax[2 * b]), tmax = ord2f(a.ws.tminmax[2 * b + 1]);
  const float trange = __fsub_rn(tmax, tmin);                       // base.py:94
  const bool hflip = a.flips && a.flips[3 * b], vflip = a.flips && a.flips[3 * b + 1], pflip = a.flips && a.flips[3 * b + 2];
  float* cnt = a.ws.cnt + (size_t)b * 2 * HW;
  int32_t* last = a.ws.last + (size_t)b * HW;
  int64_t* vox = a.ws.vox + (size_t)b * a.nb * HW;
  const float scale = (float)(a.nb - 1);
  const int64_t stride = (int64_t)gridDim.x * LD_THREADS;
  for (int64_t k = (int64_t)blockIdx.x * LD_THREADS + threadIdx.x; k < a.N; k += stride) {
    const int64_t e = (int64_t)b * a.N + k;
    float x = a.xs[e], y = a.ys[e];
    float p = __fsub_rn(__fmul_rn(a.ps[e], 2.0f), 1.0f);             // base.py:89
    float t = load_ts<TS64>(a.ts, e, t0b);
    t = trange > 0.f ? __fdiv_rn(__fsub_rn(t, tmin), trange) : 0.f;  // base.py:95-98
    if (hflip) x = __fsub_rn((float)(W - 1), x);                     // base.py:114-116
    if (vflip) y = __fsub_rn((float)(H - 1), y);                     // base.py:118-120
    if (pflip) p = __fmul_rn(p, -1.0f);                              // base.py:122-124
    // event list [B,N,4] = (ts, y, x, p) (base.py:221, collate transpose :275-276) and polarity mask [B,N,2] (:231-235)
    float ly = y, lx = x;
    if (a.rescale) {
      ly = fminf(fmaxf(__fmul_rn(y, a.sy), 0.f), a.ymax);
      lx = fminf(fmaxf(__fmul_rn(x, a.sx), 0.f), a.xmax);
    }
    reinterpret_cast<float4*>(a.ev_list)[e] = make_float4(t, ly, lx, p);
    reinterpret_cast<float2*>(a.ev_pol)[e] = make_float2(p < 0.f ? 0.f : p, __fmul_rn(p > 0.f ? 0.f : p, -1.0f));
    const int xi = (int)x, yi = (int)y;                               // .long() truncation (encodings.py:39-42)
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
    const size_t px = (size_t)yi * W + xi;
    // events_to_channels (encodings.py:77-83): p * (p masked to its sign) on the plane of its sign
    if (p != 0.f) atomicAdd(cnt + (p > 0.f ? 0 : HW) + px, __fmul_rn(p, p));
    atomicMax(last + px, (int32_t)(k + 1));                           // accumulate=False: the last event wins
    float tb = __fmul_rn(t, scale);                                   // encodings.py:56
    if (a.round_ts) tb = rintf(tb);                                   // :58-59
    const int b0 = (int)floorf(tb);
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int bin = b0 + d;
      if (bin < 0 || bin >= a.nb) continue;
      const float w = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(tb, (float)bin))));   // :63
      const float val = __fmul_rn(p, w);                                             // :64
      if (val != 0.f) {
        const long long q = __double2ll_rn((double)val * LD_FIX_SCALE);
        atomicAdd(reinterpret_cast<unsigned long long*>(vox + (size_t)bin * HW + px), (unsigned long long)q);
      }
    }
  }
}

// create_hot_mask + get_hot_event_mask: one CTA per batch slot.  The reference removes, one argmax at a time, up to
// max_px pixels whose event rate exceeds max_rate (ties: lowest flat index first), i.e. the first max_px candidates in
// the order (rate descending, index ascending).  hot_events holds integer hit counts and every pixel is divided by the
// same hot_idx, so that order is (count descending, index ascending):
//   * no more than max_px candidates: all of them go, in one pass;
//   * otherwise a binary search over the integer count finds the cut c* (largest c with >= max_px candidates of count
//     >= c); everything above c* goes, and of the candidates AT c* the lowest-index ones fill the remaining places
//     (one ordered pass with a block-wide prefix count).  log2(hot_idx) + 2 passes instead of max_px argmax passes.
__global__ void __launch_bounds__(HOT_THREADS) ld_hot_kernel(const float* __restrict__ cnt, float* __restrict__ hot_events,
                                                             int32_t* __restrict__ hot_idx, float* __restrict__ hotmask,
                                                             int HW, int max_px, int min_obvs, float max_rate) {
  const int b = blockIdx.x;
  cnt += (size_t)b * 2 * HW; hot_events += (size_t)b * HW; hotmask += (size_t)b * HW;
  __shared__ int s_cand;
  __shared__ int s_warp[HOT_THREADS / 32];
  const int idx = hot_idx[b] + 1;                                    // base.py:249
  const float fidx = (float)idx;
  if (threadIdx.x == 0) s_cand = 0;
  __syncthreads();
  int cand = 0;
  // eight pixels per thread and trip with all their loads in flight (one CTA per slot: latency bound otherwise)
  for (int i0 = threadIdx.x; i0 < HW; i0 += 8 * HOT_THREADS) {
    float c0[8], c1[8], h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * HOT_THREADS;
      const bool ok = i < HW;
      c0[j] = ok ? cnt[i] : 0.f; c1[j] = ok ? cnt[HW + i] : 0.f; h[j] = ok ? hot_events[i] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * HOT_THREADS;
      if (i >= HW) break;
      const float s = __fadd_rn(c0[j], c1[j]);                         // torch.sum(event_cnt, dim=0)   :246
      const float he = __fadd_rn(h[j], s > 0.f ? 1.f : s);             // hot_update[hot_update > 0] = 1 :247-248
      hot_events[i] = he;
      hotmask[i] = 1.f;
      cand += (__fdiv_rn(he, fidx) > max_rate) ? 1 : 0;                // event_rate = hot_events / hot_idx :250
    }
  }
  if (cand) atomicAdd(&s_cand, cand);
  __syncthreads();
  if (threadIdx.x == 0) hot_idx[b] = idx;
  const int n_cand = s_cand;
  if (!(idx > min_obvs) || n_cand == 0 || max_px <= 0) return;       // encodings.py:94
  if (n_cand <= max_px) {
    for (int i = threadIdx.x; i < HW; i += HOT_THREADS)
      if (__fdiv_rn(hot_events[i], fidx) > max_rate) hotmask[i] = 0.f;
    return;
  }
  // more candidates than places: find the cut.  Invariant: #(cand, count >= lo) >= max_px > #(cand, count >= hi + 1)
  int lo = 0, hi = idx;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const float fm = (float)mid;
    int c = 0;
    for (int i = threadIdx.x; i < HW; i += HOT_THREADS) {
      const float he = hot_events[i];
      c += (he >= fm && __fdiv_rn(he, fidx) > max_rate) ? 1 : 0;
    }
    // per-thread counts -> block total (warp shuffle + shared), identical in every thread
    c = (int)warp_sum((float)c);                                      // < 2^24: exact in fp32
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < HOT_THREADS / 32; ++w) tot += s_warp[w];
    if (tot >= max_px) lo = mid; else hi = mid - 1;
  }
  const float cut = (float)lo;
  int above = 0;
  for (int i = threadIdx.x; i < HW; i += HOT_THREADS) {
    const float he = hot_events[i];
    above += (he > cut && __fdiv_rn(he, fidx) > max_rate) ? 1 : 0;
  }
  above = (int)warp_sum((float)above);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = above;
  __syncthreads();
  int n_above = 0;
#pragma unroll
  for (int w = 0; w < HOT_THREADS / 32; ++w) n_above += s_warp[w];
  const int places = max_px - n_above;                               // >= 1 by the invariant
  // ordered pass: pixels in flat-index order, 1024 at a time; rank of a pixel among the candidates AT the cut
  int running = 0;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < HW; base += HOT_THREADS) {
    const int i = base + threadIdx.x;
    float he = 0.f;
    bool is_cand = false;
    if (i < HW) { he = hot_events[i]; is_cand = __fdiv_rn(he, fidx) > max_rate; }
    const bool at_cut = is_cand && he == cut, over = is_cand && he > cut;
    const unsigned bal = __ballot_sync(0xffffffffu, at_cut);
    __syncthreads();
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = running, total = 0;
#pragma unroll
    for (int w = 0; w < HOT_THREADS / 32; ++w) {
      const int n = s_warp[w];
      before += (w < (int)warp) ? n : 0;
      total += n;
    }
    const int rank = before + __popc(bal & ((1u << lane) - 1u));
    if (over || (at_cut && rank < places)) hotmask[i] = 0.f;
    running += total;
  }
}

struct LdFinalArgs {
  LoaderWs ws;
  const float* ps;          // raw polarities (for the mask value |p| of the last event)
  float* out_cnt; float* out_vox; float* out_mask;
  int64_t N; int B, H, W, nb, ph, pw, h, w, use_hot;
};

// hot-pixel application (h5.py:323-331) and avg_pool2d(kernel = stride = (ph, pw)) (h5.py:390-399): the window is summed
// row-major in fp32 and divided by ph*pw once, like ATen's CPU kernel.
__global__ void __launch_bounds__(LD_THREADS) ld_finalize_kernel(const LdFinalArgs a) {
  const int b = blockIdx.y;
  const int o = blockIdx.x * LD_THREADS + threadIdx.x;
  if (o >= a.h * a.w) return;
  const int oy = o / a.w, ox = o - oy * a.w;
  const size_t HW = (size_t)a.H * a.W, hw = (size_t)a.h * a.w;
  const float* cnt = a.ws.cnt + (size_t)b * 2 * HW;
  const int32_t* last = a.ws.last + (size_t)b * HW;
  const int64_t* vox = a.ws.vox + (size_t)b * a.nb * HW;
  const float* hm = a.ws.hotmask + (size_t)b * HW;
  const float div = (float)(a.ph * a.pw);
  const bool pool = a.ph > 1 || a.pw > 1;
  float c0 = 0.f, c1 = 0.f, m = 0.f;
  for (int dy = 0; dy < a.ph; ++dy)
    for (int dx = 0; dx < a.pw; ++dx) {
      const size_t px = (size_t)(oy * a.ph + dy) * a.W + (ox * a.pw + dx);
      const float k = a.use_hot ? hm[px] : 1.f;
      const int32_t li = last[px];
      float mv = 0.f;
      if (li > 0) mv = fabsf(__fsub_rn(__fmul_rn(a.ps[(size_t)b * a.N + li - 1], 2.0f), 1.0f));
      c0 = __fadd_rn(c0, __fmul_rn(cnt[px], k));
      c1 = __fadd_rn(c1, __fmul_rn(cnt[HW + px], k));
      m = __fadd_rn(m, __fmul_rn(mv, k));
    }
  if (pool) { c0 = __fdiv_rn(c0, div); c1 = __fdiv_rn(c1, div); m = __fdiv_rn(m, div); }
  a.out_cnt[(size_t)b * 2 * hw + o] = c0;
  a.out_cnt[(size_t)b * 2 * hw + hw + o] = c1;
  a.out_mask[(size_t)b * hw + o] = m;
  for (int bin = 0; bin < a.nb; ++bin) {
    float v = 0.f;
    for (int dy = 0; dy < a.ph; ++dy)
      for (int dx = 0; dx < a.pw; ++dx) {
        const size_t px = (size_t)(oy * a.ph + dy) * a.W + (ox * a.pw + dx);
        const float k = a.use_hot ? hm[px] : 1.f;
        v = __fadd_rn(v, __fmul_rn((float)((double)vox[(size_t)bin * HW + px] * LD_FIX_INV), k));
      }
    if (pool) v = __fdiv_rn(v, div);
    a.out_vox[((size_t)b * a.nb + bin) * hw + o] = v;
  }
}

static size_t loader_carve(const snnflow_loader_desc* d, char* base, LoaderWs* ws) {
  const size_t HW = (size_t)d->H * d->W, B = (size_t)d->B;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return base ? base + o : nullptr; };
  char* p_cnt = take(B * 2 * HW * sizeof(float));
  char* p_last = take(B * HW * sizeof(int32_t));
  char* p_vox = take(B * (size_t)d->num_bins * HW * sizeof(int64_t));
  const size_t zero_bytes = off;                 // everything up to here is cleared by one memset
  char* p_hot = take(B * HW * sizeof(float));
  char* p_mm = take(B * 2 * sizeof(unsigned));
  if (ws) {
    ws->cnt = (float*)p_cnt; ws->last = (int32_t*)p_last; ws->vox = (int64_t*)p_vox;
    ws->hotmask = (float*)p_hot; ws->tminmax = (unsigned*)p_mm;
  }
  (void)zero_bytes;
  return off;
}

static int loader_grid(int64_t N, int B) {
  int64_t blocks = ceil_div64(N, (int64_t)LD_THREADS * 4);
  const int64_t cap = std::max<int64_t>(1, (int64_t)sm_count() * 8 / std::max(1, B));
  if (blocks > cap) blocks = cap;
  return (int)std::max<int64_t>(blocks, 1);
}

}  // namespace snnflow
using namespace snnflow;

static int loader_check_desc(const snnflow_loader_desc* d) {
  SNNFLOW_REQUIRE(d, "null descriptor");
  SNNFLOW_REQUIRE(d->B > 0 && d->B <= 65535 && d->N >= 0 && d->N < 2147483647LL, "bad batch size / event count");
  SNNFLOW_REQUIRE(d->H > 0 && d->W > 0 && d->num_bins > 0, "bad resolution / number of bins");
  SNNFLOW_REQUIRE(d->pool_h >= 1 && d->pool_w >= 1 && d->pool_h <= d->H && d->pool_w <= d->W, "bad pooling window");
  return SNNFLOW_OK;
}

extern "C" size_t snnflow_format_window_workspace_bytes(const snnflow_loader_desc* d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->num_bins <= 0) return 0;
  return loader_carve(d, nullptr, nullptr);
}

extern "C" int snnflow_format_window(const snnflow_loader_desc* d, const float* xs, const float* ys, const void* ts,
                                     int ts_is_f64, const double* t0, const float* ps, const int32_t* flips,
                                     float* hot_events, int32_t* hot_idx, float* event_cnt, float* event_voxel,
                                     float* event_mask, float* event_list, float* event_pol, void* workspace,
                                     size_t workspace_bytes, snnflow_stream_t stream) {
  int rc = loader_check_desc(d);
  if (rc) return rc;
  SNNFLOW_REQUIRE(event_cnt && event_voxel && event_mask && workspace, "null output / workspace");
  SNNFLOW_REQUIRE(!d->hot_enabled || (hot_events && hot_idx), "hot-pixel filter enabled without its state buffers");
  SNNFLOW_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  LoaderWs ws;
  const size_t need = loader_carve(d, (char*)workspace, &ws);
  SNNFLOW_REQUIRE(workspace_bytes >= need, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, H = d->H, W = d->W;
  const int64_t N = d->N;
  const size_t HW = (size_t)H * W;
  const size_t zero_bytes = (size_t)((char*)ws.hotmask - (char*)workspace);
  SNNFLOW_CUDA(cudaMemsetAsync(workspace, 0, zero_bytes, st));
  if (N > 0) {
    SNNFLOW_REQUIRE(xs && ys && ts && ps && event_list && event_pol, "null event arrays");
    SNNFLOW_REQUIRE((((uintptr_t)event_list & 15) | ((uintptr_t)event_pol & 7)) == 0, "event list outputs must be 16-byte aligned");
    const dim3 grid(loader_grid(N, B), B);
    ld_init_kernel<<<ceil_div(B, LD_THREADS), LD_THREADS, 0, st>>>(ws.tminmax, B);
    rc = check_launch("ld_init_kernel");
    if (rc) return rc;
    prof_begin("loader_minmax", st, (ts_is_f64 ? 8.0 : 4.0) * N * B);
    if (ts_is_f64) ld_minmax_kernel<true><<<grid, LD_THREADS, 0, st>>>(ts, t0, ws.tminmax, N);
    else ld_minmax_kernel<false><<<grid, LD_THREADS, 0, st>>>(ts, t0, ws.tminmax, N);
    rc = check_launch("ld_minmax_kernel");
    if (rc) return rc;
    LdScatterArgs a;
    a.xs = xs; a.ys = ys; a.ts = ts; a.ps = ps; a.t0 = t0; a.flips = flips; a.ws = ws;
    a.ev_list = event_list; a.ev_pol = event_pol;
    a.N = N; a.H = H; a.W = W; a.nb = d->num_bins; a.round_ts = d->round_ts;
    const int h = H / d->pool_h, w = W / d->pool_w;
    a.rescale = (d->pool_h > 1 || d->pool_w > 1) ? 1 : 0;
    // the reference multiplies by the Python float target/original (h5.py:401-402): a double rounded to fp32 by ATen
    a.sy = (float)((double)d->target_h / (double)H); a.sx = (float)((double)d->target_w / (double)W);
    a.ymax = (float)(d->target_h - 1); a.xmax = (float)(d->target_w - 1);
    (void)h; (void)w;
    prof_begin("loader_scatter", st, ((ts_is_f64 ? 20.0 : 16.0) + 24.0) * N * B);
    if (ts_is_f64) ld_scatter_kernel<true><<<grid, LD_THREADS, 0, st>>>(a);
    else ld_scatter_kernel<false><<<grid, LD_THREADS, 0, st>>>(a);
    rc = check_launch("ld_scatter_kernel");
    if (rc) return rc;
  }
  if (d->hot_enabled) {
    prof_begin("loader_hot", st, 16.0 * HW * B);
    ld_hot_kernel<<<B, HOT_THREADS, 0, st>>>(ws.cnt, hot_events, hot_idx, ws.hotmask, (int)HW, d->hot_max_px,
                                            d->hot_min_obvs, d->hot_max_rate);
    rc = check_launch("ld_hot_kernel");
    if (rc) return rc;
  }
  LdFinalArgs f;
  f.ws = ws; f.ps = ps; f.out_cnt = event_cnt; f.out_vox = event_voxel; f.out_mask = event_mask;
  f.N = N; f.B = B; f.H = H; f.W = W; f.nb = d->num_bins; f.ph = d->pool_h; f.pw = d->pool_w;
  f.h = H / d->pool_h; f.w = W / d->pool_w; f.use_hot = d->hot_enabled ? 1 : 0;
  prof_begin("loader_finalize", st, (double)B * HW * (12.0 + 8.0 * d->num_bins) + (double)B * f.h * f.w * 4.0 * (3 + d->num_bins));
  ld_finalize_kernel<<<dim3(ceil_div(f.h * f.w, LD_THREADS), B), LD_THREADS, 0, st>>>(f);
  return check_launch("ld_finalize_kernel");
}

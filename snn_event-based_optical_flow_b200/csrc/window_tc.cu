// Layer-major window engine: the tensor-core tile pipeline (sm_100a: TMA bulk copies + tcgen05 + TMEM).
//
// One skeleton, three kernels.  A CTA is 10 warps with fixed roles:
//   warp 8, lane 0 : TMA producer  - streams row tiles of bf16 planes (window.cuh) into an S-stage shared-memory
//                    ring with cp.async.bulk; a tile with its halo is one contiguous range per 8-channel chunk.
//   warp 9, lane 0 : MMA issuer    - for every 128-pixel segment of the tile, 9 shifted-descriptor taps x K/16
//                    k-steps x 3 bf16 term pairs of tcgen05.mma into one of two TMEM accumulator sets.
//   warps 0..7     : epilogue      - tcgen05.ld their 32 TMEM lanes (pixels) x 16 channels and finish the layer:
//       forward        : LIF update (leak, delayed reset, threshold, spike), membrane / current / spike planes out.
//                        In sequence mode (feed-forward ConvLIF) a CTA owns a row tile for ALL T time bins and the
//                        membrane state never leaves the registers (models/spiking_submodules.py:121-151 unrolled).
//       data gradient  : g_x = conv^T(g_I, W_ff) for all T*B images of a layer in one launch.
//       recurrent bwd  : g_z = conv^T(g_I[t+1], W_rec) fused with the surrogate / leak / reset chain of step t
//                        (the BPTT recursion of ConvLIFRecurrent, spiking_submodules.py:265-300).
// Stages and accumulators are handed over with mbarriers only (full/empty, acc_full/acc_empty); the MMAs of item
// k+1 run while the epilogue of item k drains, and the copies of items k+2.. are already in flight.
//
// Exactness: spikes / counts are exact in bf16; weights are split into three bf16 terms (24 mantissa bits: every
// product exact, fp32 accumulate), gradients into hi + lo (hi*hi + hi*lo + lo*hi).
#include "tcgen05.cuh"
#include "window.cuh"

#include <stdlib.h>

#include <map>

namespace snnflow {

constexpr int WT_EPI_WARPS = 8;
constexpr int WT_THREADS = (WT_EPI_WARPS + 2) * 32;
constexpr int WT_MAX_STAGES = 4;
constexpr int WT_HDR = 4096;   // barriers, TMEM slot, per-channel parameters, reduction scratch
constexpr int WT_TAIL = 4096;  // the last 128-pixel segment of a row may address up to 128 + 2 slots past its tile: keep
                               // that (discarded) operand read inside the CTA's shared memory

struct WtSmem {
  uint64_t *full, *empty, *acc_full, *acc_empty, *wbar;
  uint32_t* tmem_slot;
  float4* par;
  float* red;
  unsigned char *w, *stages;
};

__device__ __forceinline__ WtSmem wt_smem(unsigned char* smem, uint32_t wblob_bytes) {
  WtSmem s;
  s.full = reinterpret_cast<uint64_t*>(smem);
  s.empty = s.full + WT_MAX_STAGES;
  s.acc_full = s.empty + WT_MAX_STAGES;
  s.acc_empty = s.acc_full + 2;
  s.wbar = s.acc_empty + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  s.par = reinterpret_cast<float4*>(smem + 256);
  s.red = reinterpret_cast<float*>(smem + 1280);
  s.w = smem + WT_HDR;
  s.stages = s.w + ((wblob_bytes + 127u) & ~127u);
  return s;
}

struct ItemPos {
  int img, b, y0, t;
};

template <bool SEQ>
__device__ __forceinline__ int wt_n_items(const WtArgs& a) {
  const int n_tiles = a.n_outer * (a.H / a.R);
  const int mine = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  return SEQ ? mine * a.T : mine;
}

template <bool SEQ>
__device__ __forceinline__ ItemPos wt_item(const WtArgs& a, int k) {
  const int tpi = a.H / a.R;
  ItemPos p;
  if (SEQ) {
    const int tile = blockIdx.x + (k / a.T) * gridDim.x;
    p.t = k % a.T;
    p.b = tile / tpi;
    p.y0 = (tile - p.b * tpi) * a.R;
    p.img = p.t * a.B + p.b;
  } else {
    const int tile = blockIdx.x + k * gridDim.x;
    p.t = 0;
    p.img = tile / tpi;
    p.b = p.img;
    p.y0 = (tile - p.img * tpi) * a.R;
  }
  return p;
}

__device__ __forceinline__ uint32_t wt_tmem_cols(const WtArgs& a) {
  const uint32_t need = 2u * (uint32_t)(a.R * a.n_seg * a.N) * a.src[0].w_terms;
  uint32_t c = 32;
  while (c < need) c <<= 1;
  return c;
}

// ---- common prologue: barriers, TMEM, parameters ---------------------------------------------------------
__device__ __forceinline__ uint32_t wt_prologue(const WtArgs& a, const WtSmem& s) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < WT_MAX_STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], WT_EPI_WARPS * 32);
    }
    mbar_init(s.wbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(s.tmem_slot, wt_tmem_cols(a));
  if (a.par)
    for (int i = tid; i < a.N; i += WT_THREADS) s.par[i] = __ldg(reinterpret_cast<const float4*>(a.par) + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_slot;
}

// ---- producer: one thread ------------------------------------------------------------------------------------
template <bool SEQ>
__device__ void wt_producer(const WtArgs& a, const WtSmem& s) {
  if (a.wblob_bytes) {
    mbar_expect_tx(s.wbar, a.wblob_bytes);
    tma_bulk_g2s(s.w, a.wblob, a.wblob_bytes, s.wbar);
  }
  const int n_items = wt_n_items<SEQ>(a);
  const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16;
  uint32_t u = 0;
  for (int k = 0; k < n_items; ++k) {
    const ItemPos p = wt_item<SEQ>(a, k);
    for (int si = 0; si < a.n_src; ++si, ++u) {
      const uint32_t st = u % (uint32_t)a.S, use = u / (uint32_t)a.S;
      if (use > 0) mbar_wait(&s.empty[st], (use - 1) & 1);
      const WtSrc& S = a.src[si];
      unsigned char* dst = s.stages + (size_t)st * a.stage_bytes;
      mbar_expect_tx(&s.full[st], S.n_chunks * a.sub_bytes);
      const unsigned char* g = S.planes + (size_t)p.img * S.img_stride + (size_t)p.y0 * a.Wp * 16;
      for (uint32_t ch = 0; ch < S.n_chunks; ++ch)
        tma_bulk_g2s(dst + (size_t)ch * a.chunk_stride, g + ch * plane_bytes, a.sub_bytes, &s.full[st]);
    }
  }
}

// ---- MMA issuer: one thread ----------------------------------------------------------------------------------
template <bool SEQ>
__device__ void wt_mma(const WtArgs& a, const WtSmem& s, uint32_t tmem_base) {
  const int n_items = wt_n_items<SEQ>(a);
  if (n_items == 0) return;
  mbar_wait(s.wbar, 0);
  const int n_mt = a.R * a.n_seg;
  const uint32_t ncat = a.src[0].w_terms * (uint32_t)a.N;       // accumulator columns per 128-pixel segment
  const uint32_t acc_cols = (uint32_t)n_mt * ncat;
  const uint32_t stages16 = smem_u32(s.stages) >> 4, w16 = smem_u32(s.w) >> 4;
  const uint32_t cs16 = a.chunk_stride >> 4, b_lbo = (ncat >> 3) * 128u, blbo16 = b_lbo >> 4;
  const uint32_t a_lo_c = ((cs16 & 0x3FFF) << 16), b_lo_c = ((blbo16 & 0x3FFF) << 16), d_hi = desc_hi(128);
  uint32_t u = 0;
  for (int k = 0; k < n_items; ++k) {
    const uint32_t ab = (uint32_t)k & 1u;
    if (k >= 2) mbar_wait(&s.acc_empty[ab], (uint32_t)((k >> 1) - 1) & 1u);
    tc_fence_after();
    for (int si = 0; si < a.n_src; ++si, ++u) {
      const uint32_t st = u % (uint32_t)a.S, use = u / (uint32_t)a.S;
      mbar_wait(&s.full[st], use & 1);
      tc_fence_after();
      const WtSrc& S = a.src[si];
      const uint32_t idesc = make_idesc(128, (int)(S.w_used * (uint32_t)a.N), /*bf16*/ 1, 0, 0);
      const uint32_t base16 = stages16 + ((st * a.stage_bytes) >> 4);
      const uint32_t tile16 = (S.n_chunks * 8u * ncat * 2u) >> 4;   // one tap of this source's weights
      const uint32_t wsrc16 = w16 + (S.w_off >> 4);
      const uint32_t n_kk = S.n_chunks >> 1;
      int r = 0, seg = 0;
      for (int m = 0; m < n_mt; ++m) {
        const uint32_t d = tmem_base + ab * acc_cols + (uint32_t)m * ncat;
        const uint32_t slot0 = (uint32_t)(r * a.Wp + seg * 128);
        uint32_t accumulate = si > 0 ? 1u : 0u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a_t = base16 + slot0 + (uint32_t)((tap / 3) * a.Wp + tap % 3);
          const uint32_t w_t = wsrc16 + (uint32_t)tap * tile16;
          for (uint32_t kk = 0; kk < n_kk; ++kk) {
            umma_f16_split(d, a_lo_c | (a_t + 2u * kk * cs16), d_hi, b_lo_c | (w_t + 2u * kk * blbo16), d_hi, idesc, accumulate);
            accumulate = 1u;
          }
        }
        if (++seg == a.n_seg) { seg = 0; ++r; }
      }
      umma_commit(&s.empty[st]);
    }
    umma_commit(&s.acc_full[ab]);
  }
}

__device__ __forceinline__ uint32_t bf16_pair(uint32_t mask, int i) {
  return (((mask >> (2 * i)) & 1u) ? 0x3F80u : 0u) | (((mask >> (2 * i + 1)) & 1u) ? 0x3F800000u : 0u);
}
__device__ __forceinline__ uint32_t nz16_mask(const uint4& a, const uint4& b) {
  // bit c set when the c-th bf16 of (a, b) is non-zero
  uint32_t m = 0;
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) m |= ((w[i] & 0xFFFFu) ? 1u : 0u) << (2 * i) | ((w[i] >> 16) ? 1u : 0u) << (2 * i + 1);
  return m;
}

// =================================================================================================
// fp32 tensors of the engine (membranes, currents, spike gradients) use the "c8" layout
//   [image][chunk = C/8][H*W][8 channels]
// so that the epilogue thread that owns one pixel x 16 channels moves them as 16-byte vectors.
// =================================================================================================
__device__ __forceinline__ size_t c8_off(int img, int n_chunks, int chunk, size_t HW, size_t pix) {
  return (((size_t)img * n_chunks + chunk) * HW + pix) * 8;
}
__device__ __forceinline__ void ld16_c8(const float* base, int img, int n_chunks, int g, size_t HW, size_t pix, float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float4* p = reinterpret_cast<const float4*>(base + c8_off(img, n_chunks, g * 2 + j, HW, pix));
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[8 * j + 0] = a.x; v[8 * j + 1] = a.y; v[8 * j + 2] = a.z; v[8 * j + 3] = a.w;
    v[8 * j + 4] = b.x; v[8 * j + 5] = b.y; v[8 * j + 6] = b.z; v[8 * j + 7] = b.w;
  }
}
__device__ __forceinline__ void st16_c8(float* base, int img, int n_chunks, int g, size_t HW, size_t pix, const float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float4* p = reinterpret_cast<float4*>(base + c8_off(img, n_chunks, g * 2 + j, HW, pix));
    p[0] = make_float4(v[8 * j + 0], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]);
    p[1] = make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]);
  }
}

// =================================================================================================
// Forward
// =================================================================================================
template <bool SEQ, int NSEG>
__global__ void __launch_bounds__(WT_THREADS, 1) wt_fwd_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == WT_EPI_WARPS) {
    if (elect_one()) wt_producer<SEQ>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (elect_one()) wt_mma<SEQ>(a, s, tmem_base);
    __syncwarp();
  } else {
    const int q = warp & 3, g = warp >> 2;   // TMEM lane quarter ; 16-channel group of this warp
    const bool act = g * 16 < a.N;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16;
    const int nch = a.N >> 3;
    const uint32_t ncat = 3u * (uint32_t)a.N, acc_cols = (uint32_t)NSEG * ncat;   // three weight terms side by side
    const int n_items = wt_n_items<SEQ>(a);
    float vst[NSEG][16];
    uint32_t zm[NSEG];
    for (int k = 0; k < n_items; ++k) {
      const ItemPos p = wt_item<SEQ>(a, k);
      const uint32_t ab = (uint32_t)k & 1u;
      const bool load_state = SEQ ? (p.t == 0) : true;
      if (load_state && act) {
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const int y = p.y0 + m / a.n_seg, x = (m % a.n_seg) * 128 + q * 32 + lane;
          const bool ok = x < a.W;
          const size_t pix = (size_t)y * a.W + x;
          const size_t o = ((size_t)(p.b * a.N + g * 16)) * HW + pix;   // NCHW (state tensors of the caller)
          uint32_t zmask = 0;
#pragma unroll
          for (int c = 0; c < 16; ++c) vst[m][c] = 0.f;
          if (SEQ) {
            if (ok && a.v_init) {
#pragma unroll
              for (int c = 0; c < 16; ++c) vst[m][c] = __ldg(a.v_init + o + (size_t)c * HW);
            }
            if (ok && a.z_init) {
#pragma unroll
              for (int c = 0; c < 16; ++c) zmask |= (__ldg(a.z_init + o + (size_t)c * HW) != 0.f ? 1u : 0u) << c;
            }
          } else {
            if (ok && a.v_prev) {
              if (a.v_prev_nchw) {
#pragma unroll
                for (int c = 0; c < 16; ++c) vst[m][c] = __ldg(a.v_prev + o + (size_t)c * HW);
              } else {
                ld16_c8(a.v_prev, p.b, nch, g, HW, pix, vst[m]);
              }
            }
            if (ok && a.zin_planes) {
              const unsigned char* zp = a.zin_planes + (size_t)p.b * a.zin_img_stride + (size_t)(g * 2) * plane_bytes +
                                        ((size_t)(y + 1) * a.Wp + x + 1) * 16;
              const uint4 z0 = __ldg(reinterpret_cast<const uint4*>(zp));
              const uint4 z1 = __ldg(reinterpret_cast<const uint4*>(zp + plane_bytes));
              zmask = nz16_mask(z0, z1);
            }
          }
          zm[m] = zmask;
        }
      }
      const bool last = (SEQ ? (p.t == a.T - 1) : true) && (a.v_last != nullptr || a.z_last != nullptr);
      mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
      tc_fence_after();
      if (act) {
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const int y = p.y0 + m / a.n_seg, x = (m % a.n_seg) * 128 + q * 32 + lane;
          const bool ok = x < a.W;
          const size_t pix = (size_t)y * a.W + x;
          uint32_t u0[16], u1[16], u2[16];
          const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(g * 16);
          tmem_ld16_async(tcol, u0);
          tmem_ld16_async(tcol + (uint32_t)a.N, u1);
          tmem_ld16_async(tcol + 2u * (uint32_t)a.N, u2);
          tmem_ld_wait();
          float cur[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) cur[c] = (__uint_as_float(u0[c]) + __uint_as_float(u1[c])) + __uint_as_float(u2[c]);
          const uint32_t zin = zm[m];
          uint32_t nm = 0;
          if (a.hard_reset) {   // ((v*lam)*(1-z)) + ((1-lam)*I)      spiking_submodules.py:144
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float4 pr = s.par[g * 16 + c];   // lam, 1 - lam, theta
              const float omz = ((zin >> c) & 1u) ? 0.f : 1.f;
              const float vn = __fadd_rn(__fmul_rn(__fmul_rn(vst[m][c], pr.x), omz), __fmul_rn(pr.y, cur[c]));
              vst[m][c] = vn;
              nm |= (__fsub_rn(vn, pr.z) > 0.f ? 1u : 0u) << c;
            }
          } else {              // ((v*lam) + ((1-lam)*I)) - (z*theta)   spiking_submodules.py:146
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float4 pr = s.par[g * 16 + c];
              const float zt = ((zin >> c) & 1u) ? __fmul_rn(1.0f, pr.z) : 0.f;
              const float vn = __fsub_rn(__fadd_rn(__fmul_rn(vst[m][c], pr.x), __fmul_rn(pr.y, cur[c])), zt);
              vst[m][c] = vn;
              nm |= (__fsub_rn(vn, pr.z) > 0.f ? 1u : 0u) << c;
            }
          }
          zm[m] = nm;
          if (ok) {
            unsigned char* zp = a.zp_out + (size_t)p.img * a.zp_img_stride + (size_t)(g * 2) * plane_bytes +
                                ((size_t)(y + 1) * a.Wp + x + 1) * 16;
            *reinterpret_cast<uint4*>(zp) = make_uint4(bf16_pair(nm, 0), bf16_pair(nm, 1), bf16_pair(nm, 2), bf16_pair(nm, 3));
            *reinterpret_cast<uint4*>(zp + plane_bytes) =
                make_uint4(bf16_pair(nm, 4), bf16_pair(nm, 5), bf16_pair(nm, 6), bf16_pair(nm, 7));
            if (a.v_out) st16_c8(a.v_out, p.img, nch, g, HW, pix, vst[m]);
            if (a.cur_out) st16_c8(a.cur_out, p.img, nch, g, HW, pix, cur);
            if (last) {   // the caller-visible state [2,B,C,H,W] after the window
              const size_t o = ((size_t)(p.b * a.N + g * 16)) * HW + pix;
              if (a.v_last) {
#pragma unroll
                for (int c = 0; c < 16; ++c) a.v_last[o + (size_t)c * HW] = vst[m][c];
              }
              if (a.z_last) {
#pragma unroll
                for (int c = 0; c < 16; ++c) a.z_last[o + (size_t)c * HW] = ((nm >> c) & 1u) ? 1.f : 0.f;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&s.acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, wt_tmem_cols(a));
}

// =================================================================================================
// Data gradient: g_x (c8 layout) = accumulator columns [0,N) + [N,2N)
// =================================================================================================
__global__ void __launch_bounds__(WT_THREADS, 1) wt_dgrad_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    if (elect_one()) wt_producer<false>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (elect_one()) wt_mma<false>(a, s, tmem_base);
    __syncwarp();
  } else {
    const int q = warp & 3, g = warp >> 2;
    const bool act = g * 16 < a.N;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const int n_mt = a.R * a.n_seg, nch = a.N >> 3;
    const uint32_t ncat = 2u * (uint32_t)a.N, acc_cols = (uint32_t)n_mt * ncat;   // [g*w_hi | g_hi*w_lo]
    const int n_items = wt_n_items<false>(a);
    for (int k = 0; k < n_items; ++k) {
      const ItemPos p = wt_item<false>(a, k);
      const uint32_t ab = (uint32_t)k & 1u;
      mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
      tc_fence_after();
      if (act) {
        int r = 0, seg = 0;
        for (int m = 0; m < n_mt; ++m) {
          const int y = p.y0 + r, x = seg * 128 + q * 32 + lane;
          uint32_t u0[16], u1[16];
          const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(g * 16);
          tmem_ld16_async(tcol, u0);
          tmem_ld16_async(tcol + (uint32_t)a.N, u1);
          tmem_ld_wait();
          float acc[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = __uint_as_float(u0[c]) + __uint_as_float(u1[c]);
          if (x < a.W) st16_c8(a.g_x, p.img, nch, g, HW, (size_t)y * a.W + x, acc);
          if (++seg == a.n_seg) { seg = 0; ++r; }
        }
      }
      tc_fence_before();
      mbar_arrive(&s.acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, wt_tmem_cols(a));
}

// =================================================================================================
// Recurrent backward step t:  g_z = conv^T(g_I[t+1], W_rec)  fused with the pointwise BPTT chain of step t
//   gs = (g_out + g_z) * sg(v_t - theta);  gv = g_v + gs;  g_I = gv * (1 - lam)
//   hard: g_v' = gv*lam*(1-z_in); dlam += gv*(v_in*(1-z_in) - I); dtheta -= gs
//   soft: g_v' = gv*lam;          dlam += gv*(v_in - I);          dtheta -= gs + gv*z_in
// =================================================================================================
__global__ void __launch_bounds__(WT_THREADS, 1) wt_recbwd_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    if (a.has_gz && elect_one()) wt_producer<false>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (a.has_gz && elect_one()) wt_mma<false>(a, s, tmem_base);
    __syncwarp();
  } else {
    const int q = warp & 3, g = warp >> 2;
    const bool act = g * 16 < a.N;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16;
    const int n_mt = a.R * a.n_seg, nch = a.N >> 3;
    const uint32_t ncat = 2u * (uint32_t)a.N, acc_cols = (uint32_t)n_mt * ncat;
    const int n_items = wt_n_items<false>(a);
    float s_lam[16], s_th[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) s_lam[c] = s_th[c] = 0.f;
    for (int k = 0; k < n_items; ++k) {
      const ItemPos p = wt_item<false>(a, k);
      const uint32_t ab = (uint32_t)k & 1u;
      bool waited = false;
      int r = 0, seg = 0;
      for (int m = 0; m < n_mt && act; ++m) {
        const int y = p.y0 + r, x = seg * 128 + q * 32 + lane;
        const bool ok = x < a.W;
        const size_t pix = (size_t)y * a.W + x;
        const size_t o = ((size_t)(p.b * a.N + g * 16)) * HW + pix;   // NCHW (window-initial state of the caller)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {   // two 8-channel chunks (register pressure)
          const int chunk = g * 2 + hf;
          float4 go[2], vt[2], cu[2], vi[2], gv[2];
          go[0] = go[1] = vt[0] = vt[1] = cu[0] = cu[1] = vi[0] = vi[1] = gv[0] = gv[1] = make_float4(0.f, 0.f, 0.f, 0.f);
          const size_t co = c8_off(p.b, nch, chunk, HW, pix);
          if (ok) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              go[j] = __ldg(reinterpret_cast<const float4*>(a.g_out + co) + j);
              vt[j] = __ldg(reinterpret_cast<const float4*>(a.v_t + co) + j);
              cu[j] = __ldg(reinterpret_cast<const float4*>(a.cur_t + co) + j);
              if (!a.first_step) gv[j] = reinterpret_cast<const float4*>(a.g_v + co)[j];
              if (a.v_in && !a.v_in_nchw) vi[j] = __ldg(reinterpret_cast<const float4*>(a.v_in + co) + j);
            }
          }
          float vin[8] = {vi[0].x, vi[0].y, vi[0].z, vi[0].w, vi[1].x, vi[1].y, vi[1].z, vi[1].w};
          if (ok && a.v_in && a.v_in_nchw) {
#pragma unroll
            for (int c = 0; c < 8; ++c) vin[c] = __ldg(a.v_in + o + (size_t)(hf * 8 + c) * HW);
          }
          const float gof[8] = {go[0].x, go[0].y, go[0].z, go[0].w, go[1].x, go[1].y, go[1].z, go[1].w};
          const float vtf[8] = {vt[0].x, vt[0].y, vt[0].z, vt[0].w, vt[1].x, vt[1].y, vt[1].z, vt[1].w};
          const float cuf[8] = {cu[0].x, cu[0].y, cu[0].z, cu[0].w, cu[1].x, cu[1].y, cu[1].z, cu[1].w};
          const float gvf[8] = {gv[0].x, gv[0].y, gv[0].z, gv[0].w, gv[1].x, gv[1].y, gv[1].z, gv[1].w};
          float acc[8];
          if (a.has_gz) {
            if (!waited) {
              mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
              tc_fence_after();
              waited = true;
            }
            float acc1[8];
            const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(g * 16 + hf * 8);
            tmem_ld8(tcol, acc);
            tmem_ld8(tcol + (uint32_t)a.N, acc1);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = ok ? acc[c] + acc1[c] : 0.f;   // keep out-of-row garbage out of the sums
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.f;
          }
          uint32_t hi[4], lo[4];
          float gvn[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int cc = hf * 8 + c;
            const float4 pr = s.par[g * 16 + cc];
            float z_in;
            if (a.z_from_v) z_in = (__fsub_rn(vin[c], pr.z) > 0.f) ? 1.f : 0.f;
            else z_in = (ok && a.z_init) ? __ldg(a.z_init + o + (size_t)cc * HW) : 0.f;
            const float gz = gof[c] + acc[c];
            const float gs = gz * surrogate(vtf[c] - pr.z, a.width, a.surrogate);
            const float gvv = gvf[c] + gs;
            const float gi_ = gvv * pr.y;
            if (a.hard_reset) {
              gvn[c] = gvv * pr.x * (1.0f - z_in);
              s_lam[cc] += gvv * (vin[c] * (1.0f - z_in) - cuf[c]);
              s_th[cc] -= gs;
            } else {
              gvn[c] = gvv * pr.x;
              s_lam[cc] += gvv * (vin[c] - cuf[c]);
              s_th[cc] -= gs + gvv * z_in;
            }
            const __nv_bfloat16 bh = __float2bfloat16_rn(gi_);
            const __nv_bfloat16 bl = __float2bfloat16_rn(gi_ - __bfloat162float(bh));
            const uint32_t uh = (uint32_t)__bfloat16_as_ushort(bh), ul = (uint32_t)__bfloat16_as_ushort(bl);
            if (c & 1) { hi[c >> 1] |= uh << 16; lo[c >> 1] |= ul << 16; }
            else { hi[c >> 1] = uh; lo[c >> 1] = ul; }
          }
          if (ok) {
            float4* gvp = reinterpret_cast<float4*>(a.g_v + co);
            gvp[0] = make_float4(gvn[0], gvn[1], gvn[2], gvn[3]);
            gvp[1] = make_float4(gvn[4], gvn[5], gvn[6], gvn[7]);
            unsigned char* gp = a.gp_out + (size_t)p.b * a.gp_img_stride + (size_t)chunk * plane_bytes +
                                ((size_t)(y + 1) * a.Wp + x + 1) * 16;
            *reinterpret_cast<uint4*>(gp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(gp + a.gp_term_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        if (++seg == a.n_seg) { seg = 0; ++r; }
      }
      if (a.has_gz) {
        if (!waited) mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
        tc_fence_before();
        mbar_arrive(&s.acc_empty[ab]);
      }
    }
    // per-warp partial sums of dlam / dtheta -> shared scratch [warp][2][16]
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float l = warp_sum(s_lam[c]), t = warp_sum(s_th[c]);
      if (lane == 0) {
        s.red[(warp * 2 + 0) * 16 + c] = l;
        s.red[(warp * 2 + 1) * 16 + c] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 2 * a.N) {
    const int which = tid / a.N, co = tid % a.N;
    const int g = co >> 4, c = co & 15;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += s.red[((g * 4 + q) * 2 + which) * 16 + c];
    a.part[(size_t)blockIdx.x * 2 * a.N + tid] = t;
  }
  if (warp == 0) tmem_dealloc(tmem_base, wt_tmem_cols(a));
}

// =================================================================================================
// host side
// =================================================================================================
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

int wt_grid(int n_tiles) {
  const int sms = sm_count();
  return n_tiles < sms ? n_tiles : sms;
}

bool wt_plan(int H, int W, int max_chunks_per_stage, int N, uint32_t wblob_bytes, bool seq_state, int w_terms, int* R_out, int* S_out,
             uint32_t* sub_bytes, uint32_t* chunk_stride, uint32_t* stage_bytes) {
  const int Wp = W + 2, n_seg = ceil_div(W, 128);
  if (N > 32) return false;   // one 16-channel group per epilogue warp pair
  const size_t budget = (size_t)227 * 1024 - WT_HDR - WT_TAIL - align_up(wblob_bytes, 128);
  const int forced_R = env_int("SNNFLOW_WT_R", 0), forced_S = env_int("SNNFLOW_WT_S", 0);
  int best_R = 0, best_S = 0;
  for (int R = 4; R >= 1; R >>= 1) {
    if (H % R) continue;
    if (forced_R && R != forced_R) continue;
    const int n_mt = R * n_seg;
    if (2 * n_mt * N * w_terms > 512) continue;
    if (seq_state && n_mt > 4) continue;
    const size_t cs = align_up((size_t)(R + 2) * Wp * 16, 128);
    const size_t stage = cs * max_chunks_per_stage;
    int S = (int)(budget / stage);
    if (S > WT_MAX_STAGES) S = WT_MAX_STAGES;
    if (forced_S && S > forced_S) S = forced_S;
    if (S < 2) continue;
    // prefer the tallest tile that still leaves three stages; otherwise the most stages
    if (best_R == 0 || (best_S < 3 && S > best_S)) { best_R = R; best_S = S; }
    if (best_S >= 3) break;
  }
  if (best_R == 0) return false;
  *R_out = best_R; *S_out = best_S;
  *sub_bytes = (uint32_t)((best_R + 2) * Wp * 16);
  *chunk_stride = (uint32_t)align_up((size_t)*sub_bytes, 128);
  *stage_bytes = *chunk_stride * (uint32_t)max_chunks_per_stage;
  return true;
}

static size_t wt_smem_bytes(const WtArgs& a) {
  return (size_t)WT_HDR + align_up(a.wblob_bytes, 128) + (size_t)a.S * a.stage_bytes + WT_TAIL;
}

template <typename K>
static int wt_launch(K kernel, const void* key, const WtArgs& a, cudaStream_t st, const char* what) {
  const size_t smem = wt_smem_bytes(a);
  if (smem > (size_t)227 * 1024) {
    set_error("%s: shared memory %zu exceeds 227 KB", what, smem);
    return SNNFLOW_EINVAL;
  }
  static std::map<const void*, size_t> attr;   // largest dynamic shared memory size set per kernel
  size_t& have = attr[key];
  if (smem > have) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  const int n_tiles = a.n_outer * (a.H / a.R);
  kernel<<<wt_grid(n_tiles), WT_THREADS, smem, st>>>(a);
  return check_launch(what);
}

int launch_wt_fwd(const WtArgs& a, bool seq, cudaStream_t st, const char* prof_name, double bytes, double flops) {
  const int nseg = a.R * a.n_seg;
  prof_begin(prof_name, st, bytes, flops);
#define WT_FWD_CASE(SEQ, NS) \
  if (seq == SEQ && nseg == NS) return wt_launch(wt_fwd_kernel<SEQ, NS>, (const void*)wt_fwd_kernel<SEQ, NS>, a, st, "wt_fwd_kernel");
  WT_FWD_CASE(true, 1) WT_FWD_CASE(true, 2) WT_FWD_CASE(true, 3) WT_FWD_CASE(true, 4)
  WT_FWD_CASE(false, 1) WT_FWD_CASE(false, 2) WT_FWD_CASE(false, 3) WT_FWD_CASE(false, 4)
#undef WT_FWD_CASE
  set_error("launch_wt_fwd: no kernel for %d accumulator tiles", nseg);
  return SNNFLOW_EINVAL;
}

int launch_wt_dgrad(const WtArgs& a, cudaStream_t st, double bytes, double flops) {
  prof_begin("win_dgrad", st, bytes, flops);
  return wt_launch(wt_dgrad_kernel, (const void*)wt_dgrad_kernel, a, st, "wt_dgrad_kernel");
}

int launch_wt_recbwd(const WtArgs& a, cudaStream_t st, double bytes, double flops) {
  prof_begin("win_rec_bwd", st, bytes, flops);
  return wt_launch(wt_recbwd_kernel, (const void*)wt_recbwd_kernel, a, st, "wt_recbwd_kernel");
}

}  // namespace snnflow

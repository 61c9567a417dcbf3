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
  const uint32_t need = 2u * (uint32_t)(a.R * a.n_seg * a.N);
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
  const uint32_t idesc = make_idesc(128, a.N, /*bf16*/ 1, 0, 0);
  const int n_mt = a.R * a.n_seg;
  const uint32_t acc_cols = (uint32_t)(n_mt * a.N);
  const uint32_t stages_addr = smem_u32(s.stages), w_addr = smem_u32(s.w);
  const uint32_t b_lbo = (uint32_t)(a.N >> 3) * 128;
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
      const uint32_t base = stages_addr + st * a.stage_bytes;
      const uint32_t w_tile = S.n_chunks * 8u * (uint32_t)a.N * 2u;
      for (int m = 0; m < n_mt; ++m) {
        const int r = m / a.n_seg, seg = m - r * a.n_seg;
        const uint32_t d = tmem_base + ab * acc_cols + (uint32_t)(m * a.N);
        uint32_t accumulate = si > 0 ? 1u : 0u;
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t slot = (uint32_t)((r + tap / 3) * a.Wp + seg * 128 + tap % 3) * 16u;
          const uint32_t wt = w_addr + S.w_off + (uint32_t)tap * S.w_terms * w_tile;
          for (uint32_t kk = 0; kk < (S.n_chunks >> 1); ++kk) {
            const uint32_t a0 = base + 2u * kk * a.chunk_stride + slot;
            const uint32_t w0 = wt + 2u * kk * b_lbo;
            const uint64_t ad0 = make_desc(a0, a.chunk_stride, 128);
            for (uint32_t wi = 0; wi < S.w_used; ++wi) {
              umma_f16(d, ad0, make_desc(w0 + wi * w_tile, b_lbo, 128), idesc, accumulate);
              accumulate = 1u;
            }
          }
        }
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
// Forward
// =================================================================================================
template <bool SEQ, int NSEG, int NG>
__global__ void __launch_bounds__(WT_THREADS, 1) wt_fwd_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == WT_EPI_WARPS) {
    if (lane == 0) wt_producer<SEQ>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (lane == 0) wt_mma<SEQ>(a, s, tmem_base);
    __syncwarp();
  } else {
    const int q = warp & 3, h = warp >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16;
    const uint32_t acc_cols = (uint32_t)(NSEG * a.N);
    const int n_items = wt_n_items<SEQ>(a);
    float vst[NSEG][NG][16];
    uint32_t zm[NSEG][NG];
    for (int k = 0; k < n_items; ++k) {
      const ItemPos p = wt_item<SEQ>(a, k);
      const uint32_t ab = (uint32_t)k & 1u;
      const bool load_state = SEQ ? (p.t == 0) : true;
      if (load_state) {
        const float* vp = SEQ ? a.v_init : a.v_prev;
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const int y = p.y0 + m / a.n_seg, x = (m % a.n_seg) * 128 + q * 32 + lane;
          const bool ok = x < a.W;
#pragma unroll
          for (int gi = 0; gi < NG; ++gi) {
            const int g = h + 2 * gi;
            if (g * 16 >= a.N) continue;
            const size_t o = ((size_t)(p.b * a.N + g * 16)) * HW + (size_t)y * a.W + x;
#pragma unroll
            for (int c = 0; c < 16; ++c) vst[m][gi][c] = (ok && vp) ? __ldg(vp + o + (size_t)c * HW) : 0.f;
            uint32_t zmask = 0;
            if (SEQ) {
              if (ok && a.z_init) {
#pragma unroll
                for (int c = 0; c < 16; ++c) zmask |= (__ldg(a.z_init + o + (size_t)c * HW) != 0.f ? 1u : 0u) << c;
              }
            } else if (ok && a.zin_planes) {
              const unsigned char* zp = a.zin_planes + (size_t)p.b * a.zin_img_stride + (size_t)(g * 2) * plane_bytes +
                                        ((size_t)(y + 1) * a.Wp + x + 1) * 16;
              const uint4 z0 = __ldg(reinterpret_cast<const uint4*>(zp));
              const uint4 z1 = __ldg(reinterpret_cast<const uint4*>(zp + plane_bytes));
              zmask = nz16_mask(z0, z1);
            }
            zm[m][gi] = zmask;
          }
        }
      }
      const bool last = SEQ ? (p.t == a.T - 1) : true;
      mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < NSEG; ++m) {
        const int y = p.y0 + m / a.n_seg, x = (m % a.n_seg) * 128 + q * 32 + lane;
        const bool ok = x < a.W;
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const int g = h + 2 * gi;
          if (g * 16 >= a.N) continue;
          float acc[16];
          tmem_ld16(tmem_base + t_lane + ab * acc_cols + (uint32_t)(m * a.N + g * 16), acc);
          const size_t o = ((size_t)(p.img * a.N + g * 16)) * HW + (size_t)y * a.W + x;
          const size_t ob = ((size_t)(p.b * a.N + g * 16)) * HW + (size_t)y * a.W + x;
          const uint32_t zin = zm[m][gi];
          uint32_t nm = 0;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float4 pr = s.par[g * 16 + c];   // lam, 1 - lam, theta
            const float v = vst[m][gi][c], z = ((zin >> c) & 1u) ? 1.f : 0.f, cur = acc[c];
            const float t1 = __fmul_rn(v, pr.x), t3 = __fmul_rn(pr.y, cur);
            const float vn = a.hard_reset ? __fadd_rn(__fmul_rn(t1, __fsub_rn(1.0f, z)), t3)
                                          : __fsub_rn(__fadd_rn(t1, t3), __fmul_rn(z, pr.z));
            const bool sp = __fsub_rn(vn, pr.z) > 0.f;
            vst[m][gi][c] = vn;
            nm |= (sp ? 1u : 0u) << c;
            if (ok) {
              if (a.v_out) a.v_out[o + (size_t)c * HW] = vn;
              if (a.cur_out) a.cur_out[o + (size_t)c * HW] = cur;
              if (last) {
                if (a.v_last) a.v_last[ob + (size_t)c * HW] = vn;
                if (a.z_last) a.z_last[ob + (size_t)c * HW] = sp ? 1.f : 0.f;
              }
            }
          }
          zm[m][gi] = nm;
          if (ok) {
            unsigned char* zp = a.zp_out + (size_t)p.img * a.zp_img_stride + (size_t)(g * 2) * plane_bytes +
                                ((size_t)(y + 1) * a.Wp + x + 1) * 16;
            *reinterpret_cast<uint4*>(zp) = make_uint4(bf16_pair(nm, 0), bf16_pair(nm, 1), bf16_pair(nm, 2), bf16_pair(nm, 3));
            *reinterpret_cast<uint4*>(zp + plane_bytes) =
                make_uint4(bf16_pair(nm, 4), bf16_pair(nm, 5), bf16_pair(nm, 6), bf16_pair(nm, 7));
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
// Data gradient: g_x[img][n][y][x] = accumulator
// =================================================================================================
__global__ void __launch_bounds__(WT_THREADS, 1) wt_dgrad_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    if (lane == 0) wt_producer<false>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (lane == 0) wt_mma<false>(a, s, tmem_base);
    __syncwarp();
  } else {
    const int q = warp & 3, h = warp >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const int n_mt = a.R * a.n_seg;
    const uint32_t acc_cols = (uint32_t)(n_mt * a.N);
    const int n_items = wt_n_items<false>(a);
    for (int k = 0; k < n_items; ++k) {
      const ItemPos p = wt_item<false>(a, k);
      const uint32_t ab = (uint32_t)k & 1u;
      mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
      tc_fence_after();
      for (int m = 0; m < n_mt; ++m) {
        const int y = p.y0 + m / a.n_seg, x = (m % a.n_seg) * 128 + q * 32 + lane;
        const bool ok = x < a.W;
        for (int g = h; g * 16 < a.N; g += 2) {
          float acc[16];
          tmem_ld16(tmem_base + t_lane + ab * acc_cols + (uint32_t)(m * a.N + g * 16), acc);
          if (ok) {
            float* o = a.g_x + ((size_t)(p.img * a.N + g * 16)) * HW + (size_t)y * a.W + x;
#pragma unroll
            for (int c = 0; c < 16; ++c) o[(size_t)c * HW] = acc[c];
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
// Recurrent backward step t:  g_z = conv^T(g_I[t+1], W_rec)  fused with the pointwise BPTT chain of step t
//   gs = (g_out + g_z) * sg(v_t - theta);  gv = g_v + gs;  g_I = gv * (1 - lam)
//   hard: g_v' = gv*lam*(1-z_in); dlam += gv*(v_in*(1-z_in) - I); dtheta -= gs
//   soft: g_v' = gv*lam;          dlam += gv*(v_in - I);          dtheta -= gs + gv*z_in
// =================================================================================================
template <int NG>
__global__ void __launch_bounds__(WT_THREADS, 1) wt_recbwd_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    if (lane == 0 && a.has_gz) wt_producer<false>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (lane == 0 && a.has_gz) wt_mma<false>(a, s, tmem_base);
    __syncwarp();
  } else {
    const int q = warp & 3, h = warp >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16;
    const int n_mt = a.R * a.n_seg;
    const uint32_t acc_cols = (uint32_t)(n_mt * a.N);
    const int n_items = wt_n_items<false>(a);
    float s_lam[NG][16], s_th[NG][16];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi)
#pragma unroll
      for (int c = 0; c < 16; ++c) s_lam[gi][c] = s_th[gi][c] = 0.f;
    for (int k = 0; k < n_items; ++k) {
      const ItemPos p = wt_item<false>(a, k);
      const uint32_t ab = (uint32_t)k & 1u;
      bool waited = false;
      for (int m = 0; m < n_mt; ++m) {
        const int y = p.y0 + m / a.n_seg, x = (m % a.n_seg) * 128 + q * 32 + lane;
        const bool ok = x < a.W;
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const int g = h + 2 * gi;
          if (g * 16 >= a.N) continue;
          const size_t o = ((size_t)(p.b * a.N + g * 16)) * HW + (size_t)y * a.W + x;
          unsigned char* gp = a.gp_out + (size_t)p.b * a.gp_img_stride + (size_t)(g * 2) * plane_bytes +
                              ((size_t)(y + 1) * a.Wp + x + 1) * 16;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {   // two 8-channel chunks (register pressure)
            float go[8], vt[8], vi[8], cu[8], gv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const size_t i = o + (size_t)(hf * 8 + c) * HW;
              go[c] = ok ? __ldg(a.g_out + i) : 0.f;
              vt[c] = ok ? __ldg(a.v_t + i) : 0.f;
              cu[c] = ok ? __ldg(a.cur_t + i) : 0.f;
              vi[c] = (ok && a.v_in) ? __ldg(a.v_in + i) : 0.f;
              gv[c] = (ok && !a.first_step) ? a.g_v[i] : 0.f;
            }
            float acc[8];
            if (a.has_gz) {
              if (!waited) {
                mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
                tc_fence_after();
                waited = true;
              }
              tmem_ld8(tmem_base + t_lane + ab * acc_cols + (uint32_t)(m * a.N + g * 16 + hf * 8), acc);
              if (!ok) {   // pixels past the row end accumulate whatever the operand read found: keep them out of the sums
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = 0.f;
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[c] = 0.f;
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int cc = hf * 8 + c;
              const float4 pr = s.par[g * 16 + cc];
              float z_in;
              if (a.z_from_v) z_in = (__fsub_rn(vi[c], pr.z) > 0.f) ? 1.f : 0.f;
              else z_in = (ok && a.z_init) ? __ldg(a.z_init + o + (size_t)cc * HW) : 0.f;
              const float gz = go[c] + acc[c];
              const float gs = gz * surrogate(vt[c] - pr.z, a.width, a.surrogate);
              const float gvv = gv[c] + gs;
              const float gi_ = gvv * pr.y;
              float gvn;
              if (a.hard_reset) {
                gvn = gvv * pr.x * (1.0f - z_in);
                s_lam[gi][cc] += gvv * (vi[c] * (1.0f - z_in) - cu[c]);
                s_th[gi][cc] -= gs;
              } else {
                gvn = gvv * pr.x;
                s_lam[gi][cc] += gvv * (vi[c] - cu[c]);
                s_th[gi][cc] -= gs + gvv * z_in;
              }
              if (ok) a.g_v[o + (size_t)cc * HW] = gvn;
              const __nv_bfloat16 bh = __float2bfloat16_rn(gi_);
              const __nv_bfloat16 bl = __float2bfloat16_rn(gi_ - __bfloat162float(bh));
              const uint32_t uh = (uint32_t)__bfloat16_as_ushort(bh), ul = (uint32_t)__bfloat16_as_ushort(bl);
              if (c & 1) { hi[c >> 1] |= uh << 16; lo[c >> 1] |= ul << 16; }
              else { hi[c >> 1] = uh; lo[c >> 1] = ul; }
            }
            if (ok) {
              *reinterpret_cast<uint4*>(gp + (size_t)hf * plane_bytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(gp + a.gp_term_stride + (size_t)hf * plane_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      }
      if (a.has_gz) {
        if (!waited) mbar_wait(&s.acc_full[ab], (uint32_t)(k >> 1) & 1u);
        tc_fence_before();
        mbar_arrive(&s.acc_empty[ab]);
      }
    }
    // per-warp partial sums of dlam / dtheta -> shared scratch [warp][2][16*NG]
#pragma unroll
    for (int gi = 0; gi < NG; ++gi)
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float l = warp_sum(s_lam[gi][c]), t = warp_sum(s_th[gi][c]);
        if (lane == 0) {
          s.red[(warp * 2 + 0) * (16 * NG) + gi * 16 + c] = l;
          s.red[(warp * 2 + 1) * (16 * NG) + gi * 16 + c] = t;
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 2 * a.N) {
    const int which = tid / a.N, co = tid % a.N;
    const int g = co >> 4, hh = g & 1, gi = g >> 1, c = co & 15;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += s.red[((hh * 4 + q) * 2 + which) * (16 * NG) + gi * 16 + c];
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

bool wt_plan(int H, int W, int max_chunks_per_stage, int N, uint32_t wblob_bytes, bool seq_state, int* R_out, int* S_out,
             uint32_t* sub_bytes, uint32_t* chunk_stride, uint32_t* stage_bytes) {
  const int Wp = W + 2, n_seg = ceil_div(W, 128);
  const int NG = N > 32 ? 2 : 1;
  const size_t budget = (size_t)227 * 1024 - WT_HDR - WT_TAIL - align_up(wblob_bytes, 128);
  const int forced_R = env_int("SNNFLOW_WT_R", 0), forced_S = env_int("SNNFLOW_WT_S", 0);
  int best_R = 0, best_S = 0;
  for (int R = 4; R >= 1; R >>= 1) {
    if (H % R) continue;
    if (forced_R && R != forced_R) continue;
    const int n_mt = R * n_seg;
    if (2 * n_mt * N > 512) continue;
    if (seq_state && n_mt * NG > 4) continue;
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
static int wt_launch(K kernel, const WtArgs& a, cudaStream_t st, const char* what) {
  const size_t smem = wt_smem_bytes(a);
  if (smem > (size_t)227 * 1024) {
    set_error("%s: shared memory %zu exceeds 227 KB", what, smem);
    return SNNFLOW_EINVAL;
  }
  SNNFLOW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_tiles = a.n_outer * (a.H / a.R);
  kernel<<<wt_grid(n_tiles), WT_THREADS, smem, st>>>(a);
  return check_launch(what);
}

int launch_wt_fwd(const WtArgs& a, bool seq, cudaStream_t st, const char* prof_name, double bytes, double flops) {
  const int nseg = a.R * a.n_seg, ng = a.N > 32 ? 2 : 1;
  prof_begin(prof_name, st, bytes, flops);
#define WT_FWD_CASE(SEQ, NS, NGV) \
  if (seq == SEQ && nseg == NS && ng == NGV) return wt_launch(wt_fwd_kernel<SEQ, NS, NGV>, a, st, "wt_fwd_kernel");
  WT_FWD_CASE(true, 1, 1) WT_FWD_CASE(true, 2, 1) WT_FWD_CASE(true, 3, 1) WT_FWD_CASE(true, 4, 1) WT_FWD_CASE(true, 1, 2) WT_FWD_CASE(true, 2, 2)
  WT_FWD_CASE(false, 1, 1) WT_FWD_CASE(false, 2, 1) WT_FWD_CASE(false, 3, 1) WT_FWD_CASE(false, 4, 1) WT_FWD_CASE(false, 1, 2) WT_FWD_CASE(false, 2, 2)
#undef WT_FWD_CASE
  set_error("launch_wt_fwd: no kernel for %d accumulator tiles x %d channel groups", nseg, ng);
  return SNNFLOW_EINVAL;
}

int launch_wt_dgrad(const WtArgs& a, cudaStream_t st, double bytes, double flops) {
  prof_begin("win_dgrad", st, bytes, flops);
  return wt_launch(wt_dgrad_kernel, a, st, "wt_dgrad_kernel");
}

int launch_wt_recbwd(const WtArgs& a, cudaStream_t st, double bytes, double flops) {
  prof_begin("win_rec_bwd", st, bytes, flops);
  if (a.N > 32) return wt_launch(wt_recbwd_kernel<2>, a, st, "wt_recbwd_kernel");
  return wt_launch(wt_recbwd_kernel<1>, a, st, "wt_recbwd_kernel");
}

}  // namespace snnflow

// Layer-major window engine: the tensor-core tile pipeline (sm_100a: TMA bulk copies + tcgen05 + TMEM).
//
// One skeleton, three kernels.  A CTA is 18 warps with fixed roles:
//   warp 16 (one elected lane): TMA producer  - streams row tiles of bf16 planes (window.cuh) into an S-stage shared-memory
//                    ring with cp.async.bulk; a tile with its halo is one contiguous range per 8-channel chunk.
//   warp 17 (one elected lane): MMA issuer    - for every 128-pixel segment of the tile, 9 shifted-descriptor taps x K/16
//                    k-steps x 3 bf16 term pairs of tcgen05.mma into one of two TMEM accumulator sets.
//   warps 0..15    : epilogue      - tcgen05.ld their 32 TMEM lanes (pixels) x 8 channels and finish the layer:
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
#include <vector>

namespace snnflow {

constexpr int WT_EPI_WARPS = 16;   // 4 TMEM lane quarters x 4 eight-channel chunks
constexpr int WT_THREADS = (WT_EPI_WARPS + 4) * 32;   // + TMA producer, MMA issuer, tile-flag publisher, second TMA producer (20
                                                      // warps = 5 per scheduler: the register cap per thread stays 96)
constexpr int WT_PRODUCER2 = WT_EPI_WARPS + 3;        // warp index of the second producer
constexpr int WT_PUB_RING = 8;
constexpr int WT_MAX_STAGES = 4;
constexpr int WT_HDR = 4096;   // barriers, TMEM slot, per-channel parameters, reduction scratch
constexpr int WT_TAIL = 2304;  // the last 128-pixel segment of a row may address up to 128 + 2 slots past its tile: keep
                               // that (discarded) operand read inside the CTA's shared memory

struct WtSmem {
  uint64_t *full, *empty, *acc_full, *acc_empty, *wbar;
  uint32_t* tmem_slot;
  uint64_t* pub_bar;        // [WT_PUB_RING] "every epilogue thread has issued the stores of item g" (time-fused mode)
  volatile unsigned int* pub_count;   // items whose tile flag the publisher warp has raised
  uint64_t *aux_full, *aux_empty;     // [4] ring of staged epilogue inputs (WtArgs.aux)
  unsigned char* aux;
  float4* par;
  float* red;
  unsigned char *w, *stages;
};

__device__ __forceinline__ WtSmem wt_smem(unsigned char* smem, uint32_t wblob_bytes /* both blobs */, uint32_t stages_bytes = 0) {
  WtSmem s;
  s.full = reinterpret_cast<uint64_t*>(smem);
  s.empty = s.full + WT_MAX_STAGES;
  s.acc_full = s.empty + WT_MAX_STAGES;
  s.acc_empty = s.acc_full + 4;
  s.wbar = s.acc_empty + 4;
  s.tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  s.pub_count = reinterpret_cast<volatile unsigned int*>(smem + 208);
  s.pub_bar = reinterpret_cast<uint64_t*>(smem + 3072);   // behind the reduction scratch (1280 .. 2304)
  s.par = reinterpret_cast<float4*>(smem + 256);
  s.red = reinterpret_cast<float*>(smem + 1280);
  s.w = smem + WT_HDR;
  s.stages = s.w + ((wblob_bytes + 127u) & ~127u);
  s.aux_full = reinterpret_cast<uint64_t*>(smem + 3200);
  s.aux_empty = s.aux_full + 4;
  s.aux = s.stages + stages_bytes + WT_TAIL;   // behind the operand stages and their read-past tail
  return s;
}

template <bool SEQ>
__device__ __forceinline__ int wt_n_items(const WtArgs& a) {
  const int n_tiles = a.n_outer * (a.H / a.R) * a.n_col;
  const int mine = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  return SEQ ? mine * a.T : mine;
}

// Walks this CTA's pipeline items in launch order without per-item divisions.
//   sequence mode: tile (b, y0) fixed for T consecutive items (t = 0..T-1, image t*B + b), then the next tile;
//   step mode    : one item per tile, images and row blocks advance by gridDim.x tiles.
template <bool SEQ>
struct ItemIter {
  int img, b, y0, x0, t;
  int tile, tpi, T, B, R, H, grid, step_rows, n_col;
  // tile -> (b, y0, x0).  Column tiling (n_col > 1, forward kernels on rows wider than one 128-pixel segment): consecutive
  // tiles are the column tiles of one row block, so neighbours in launch order share their halo rows in L2.
  __device__ __forceinline__ void locate() {
    b = tile / tpi;
    const int r = tile - b * tpi;
    if (n_col > 1) {
      const int yb = r / n_col;
      x0 = (r - yb * n_col) * 128;
      y0 = yb * R;
    } else {
      x0 = 0;
      y0 = r * R;
    }
  }
  __device__ __forceinline__ void init(const WtArgs& a) {
    n_col = a.n_col;
    tpi = (a.H / a.R) * n_col; T = a.T; B = a.B; R = a.R; H = a.H; grid = (int)gridDim.x; step_rows = grid * a.R;
    tile = (int)blockIdx.x;
    locate();
    t = 0;
    img = b;
  }
  __device__ __forceinline__ void next() {
    if (SEQ) {
      if (++t < T) { img += B; return; }
      t = 0;
      tile += grid;
      locate();
      img = b;
    } else if (n_col > 1) {
      tile += grid;
      locate();
      img = b;
    } else {
      y0 += step_rows;
      while (y0 >= H) { y0 -= H; ++b; }
      img = b;
    }
  }
};

__device__ __forceinline__ uint32_t wt_tmem_cols(const WtArgs& a) {
  const uint32_t need = ((uint32_t)(a.R * a.n_seg * a.N) * a.src[0].w_terms) << a.acc_lg;
  uint32_t c = 32;
  while (c < need) c <<= 1;
  return c;
}

// ---- common prologue: barriers, TMEM, parameters ---------------------------------------------------------
__device__ __forceinline__ uint32_t wt_prologue(const WtArgs& a, const WtSmem& s) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long t_entry = clock64();
  pdl_launch_dependents();   // the next launch may start its own prologue as soon as this grid's CTAs retire
  if (tid == 0) {
    for (int i = 0; i < WT_MAX_STAGES; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], WT_EPI_WARPS * 32);
    }
    mbar_init(s.wbar, 1);
    for (int i = 0; i < WT_PUB_RING; ++i) mbar_init(&s.pub_bar[i], WT_EPI_WARPS * 32);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s.aux_full[i], 1);
      mbar_init(&s.aux_empty[i], WT_EPI_WARPS * 32);
    }
    *s.pub_count = 0u;
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(s.tmem_slot, wt_tmem_cols(a));
  if (a.par)
    for (int i = tid; i < a.N; i += WT_THREADS) s.par[i] = __ldg(reinterpret_cast<const float4*>(a.par) + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.dbg && tid == 0) a.dbg[blockIdx.x * 8 + 7] = clock64() - t_entry;
  return *s.tmem_slot;
}

// ---- producer: one warp; lane 0 owns the barriers, lanes 0..n_chunks-1 issue one bulk copy each -------------------
// Two producer warps share the stage ring: warp `pw` owns every second RING SLOT USE (slot use u = the u-th operand tile of
// this CTA, in the order the MMA thread consumes them; stage u % S).  One warp spends ~570 cycles of latency per slot (barrier
// wait -> expect_tx -> copies) plus ~50 per copy, which bounds a CTA at one tile per ~1200 cycles with 12 copies
// (profiles/micro/tma_rate.cu: 10 -> 20 B/clk/SM with two issuing warps); the slots of two warps overlap.
// Ownership by slot parity keeps the parity waits sound: with an even S a stage has ONE producer; with an odd S the owner of
// use u issued use u - 2 before, which waited for the MMA to pass use u - 2 - S >= u - 2S, so the stage's barrier is never
// more than one phase behind the waiter (ownership by ITEM breaks this when an item holds several slots: a waiter two
// phases ahead passes a parity wait at once).
template <bool SEQ>
__device__ void wt_producer(const WtArgs& a, const WtSmem& s, int pw = 0) {
  const int lane = threadIdx.x & 31;
  const int n_pw = (a.n_prod > 1 && a.aux == nullptr) ? 2 : 1;
  if (pw >= n_pw) return;
  if (pw == 0 && lane == 0 && a.wblob_bytes) {
    mbar_expect_tx(s.wbar, a.wblob_bytes + a.wblob2_bytes);
    tma_bulk_g2s(s.w, a.wblob, a.wblob_bytes, s.wbar);
    if (a.wblob2_bytes) tma_bulk_g2s(s.w + a.wblob_bytes, a.wblob2, a.wblob2_bytes, s.wbar);
  }
  pdl_wait();   // the planes are written by the preceding launches (the weights were packed at the start of the pass)
  const int n_items = wt_n_items<SEQ>(a), n_src = a.n_src, S = a.S;
  const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16, row_bytes = (size_t)a.Wp * 16;
  const uint32_t sub_bytes = a.sub_bytes, chunk_stride = a.chunk_stride, stage_bytes = a.stage_bytes;
  const bool t_rev = a.t_reverse != 0;
  const int n_col = a.n_col;
  const uint32_t col_row_bytes = (uint32_t)a.Wsm * 16u;
  // this lane's chunk of either source
  const unsigned char* base[WT_MAX_SRC];
  size_t img_stride[WT_MAX_SRC];
  uint32_t n_chunks[WT_MAX_SRC];
#pragma unroll
  for (int si = 0; si < WT_MAX_SRC; ++si) {
    const WtSrc& Sr = a.src[si < n_src ? si : 0];
    base[si] = Sr.planes + (size_t)lane * plane_bytes;
    img_stride[si] = Sr.img_stride;
    n_chunks[si] = Sr.n_chunks;
  }
  ItemIter<SEQ> it;
  uint32_t st = 0, use = 0, as = 0, aux_use = 0;
  long long t_wait = 0, t_flag = 0;
  const long long t_begin = clock64();
  const int n_bins = (!SEQ && a.n_bins > 1) ? a.n_bins : 1, dep_mask = a.bin_dep_mask;
  int turn = 0;   // whose ring slot is next (round robin over all operand tiles of all items and bins)
  for (int bin = 0; bin < n_bins; ++bin) {
    it.init(a);
    const int n_src_b = (bin == 0 && a.n_src_bin0 > 0) ? a.n_src_bin0 : n_src;
    for (int k = 0; k < n_items; ++k) {
#pragma unroll
      for (int si = 0; si < WT_MAX_SRC; ++si) {
        if (si >= n_src_b) break;
        const bool mine = turn == pw;
        if (++turn == n_pw) turn = 0;
        if (!mine) {   // the other producer's ring slot
          if (++st == (uint32_t)S) { st = 0; ++use; }
          continue;
        }
        if (bin > 0 && ((dep_mask >> si) & 1) && !(a.exp & 2)) {
          // rows y0-1 .. y0+R of this source were written by the epilogues of bin - 1 of this tile and of its two row
          // neighbours in the image (other CTAs): lanes 0..2 acquire one progress flag each
          const int tile = (int)blockIdx.x + k * (int)gridDim.x;
          const long long t0 = clock64();
          if (lane < 3) {
            const bool need = lane == 1 || (lane == 0 && it.y0 > 0) || (lane == 2 && it.y0 + it.R < it.H);
            if (need) tile_flag_wait(a.tile_flags + tile + lane - 1, (unsigned int)bin);
          }
          __syncwarp();
          fence_proxy_async_global();
          t_flag += clock64() - t0;
        }
        if (lane == 0) {
          if (use > 0) {
            const long long t0 = clock64();
            mbar_wait(&s.empty[st], (use - 1) & 1);
            t_wait += clock64() - t0;
          }
          mbar_expect_tx(&s.full[st], n_chunks[si] * sub_bytes);
        }
        __syncwarp();
        if (n_col > 1) {
          // column tile: (R + 2) rows of 128 + 2 pixels per chunk, one bulk copy per row and chunk (lane = chunk * rows + row)
          const int rows = it.R + 2, chunk = lane / rows, row = lane - chunk * rows;
          if ((uint32_t)chunk < n_chunks[si]) {
            const int img = (SEQ && t_rev) ? (it.T - 1 - it.t) * it.B + it.b : it.img;
            tma_bulk_g2s(s.stages + (size_t)st * stage_bytes + (size_t)chunk * chunk_stride + (size_t)row * col_row_bytes,
                         a.src[si].planes + (size_t)chunk * plane_bytes + (long long)bin * a.bin_src_stride[si] +
                             (size_t)img * img_stride[si] + (size_t)(it.y0 + row) * row_bytes + (size_t)it.x0 * 16,
                         col_row_bytes, &s.full[st]);
          }
        } else if ((uint32_t)lane < n_chunks[si]) {
          const int img = (SEQ && t_rev) ? (it.T - 1 - it.t) * it.B + it.b : it.img;
          tma_bulk_g2s(s.stages + (size_t)st * stage_bytes + (size_t)lane * chunk_stride,
                       base[si] + (long long)bin * a.bin_src_stride[si] + (size_t)img * img_stride[si] + (size_t)it.y0 * row_bytes,
                       sub_bytes, &s.full[st]);
        }
        if (++st == (uint32_t)S) { st = 0; ++use; }
      }
      if (SEQ && a.aux != nullptr) {
        // the epilogue input of this item: v of the bin BEFORE the one being processed (bins walk backwards), tile rows
        const int t = t_rev ? it.T - 1 - it.t : it.t;
        if (t > 0) {
          if (lane == 0) {
            if (aux_use > 0) mbar_wait(&s.aux_empty[as], (aux_use - 1) & 1);
            mbar_expect_tx(&s.aux_full[as], (uint32_t)(a.N >> 3) * a.aux_chunk_bytes);
          }
          __syncwarp();
          if (lane < (a.N >> 3)) {
            const size_t HWp = (size_t)a.H * a.W;
            const float* src = a.aux + ((((size_t)((t - 1) * it.B + it.b) * (a.N >> 3) + lane) * HWp + (size_t)it.y0 * a.W) << 3);
            tma_bulk_g2s(s.aux + (size_t)as * a.aux_stage_bytes + (size_t)lane * a.aux_chunk_bytes, src, a.aux_chunk_bytes, &s.aux_full[as]);
          }
          if (++as == (uint32_t)a.aux_slots) { as = 0; ++aux_use; }
        }
      }
      it.next();
    }
  }
  if (a.dbg && lane == 0 && pw == 0) {
    a.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;   // producer: total
    a.dbg[blockIdx.x * 8 + 1] = t_wait + (t_flag << 32);   // producer: waiting for a free stage | for neighbour tiles (high word)
  }
}

// ---- MMA issuer: one thread ----------------------------------------------------------------------------------
template <bool SEQ, int NS = 2>
__device__ void wt_mma(const WtArgs& a, const WtSmem& s, uint32_t tmem_base) {
  const int n_items = wt_n_items<SEQ>(a) * ((!SEQ && a.n_bins > 1) ? a.n_bins : 1);
  if (n_items == 0) return;
  mbar_wait(s.wbar, 0);
  const int n_mt = a.R * a.n_seg;
  const uint32_t ncat = a.src[0].w_terms * (uint32_t)a.N;       // accumulator columns per 128-pixel segment
  const uint32_t acc_cols = (uint32_t)n_mt * ncat;
  const uint32_t stages16 = smem_u32(s.stages) >> 4, w16 = smem_u32(s.w) >> 4;
  const uint32_t cs16 = a.chunk_stride >> 4, blbo16 = ((ncat >> 3) * 128u) >> 4;
  const uint32_t a_lo_c = ((cs16 & 0x3FFF) << 16), b_lo_c = ((blbo16 & 0x3FFF) << 16), d_hi = desc_hi(128);
  // The issuing thread shares its scheduler with four epilogue warps, so the per-MMA instruction count matters more
  // than anything else here: every descriptor low word is precomputed (18 = 9 taps x 2 k-steps per source), and one
  // integer add per MMA rebases the A descriptor onto the current stage / segment.
  // Kernels with up to two sources keep all 36 B descriptor words in registers; the four-source kernel (fused recurrent
  // backward) rebuilds them with one multiply-add per tap (a 72-word table would spill in this thread).
  uint32_t aoff[9][2], bdesc[NS <= 2 ? 2 : 1][9][2], idesc[NS], n_kk[NS], bbase[NS], tile16[NS];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) aoff[tap][kk] = a_lo_c | ((uint32_t)((tap / 3) * a.Wsm + tap % 3) + 2u * kk * cs16);
#pragma unroll
  for (int si = 0; si < NS; ++si) {
    const WtSrc& S = a.src[si < a.n_src ? si : 0];
    tile16[si] = (S.n_chunks * 8u * ncat * 2u) >> 4;   // one tap of this source's weights
    idesc[si] = make_idesc(128, (int)(S.w_used * (uint32_t)a.N), /*bf16*/ 1, 0, 0);
    n_kk[si] = S.n_chunks >> 1;
    bbase[si] = b_lo_c | (w16 + (S.w_off >> 4));
    if (NS <= 2) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) bdesc[si][tap][kk] = bbase[si] + (uint32_t)tap * tile16[si] + 2u * kk * blbo16;
    }
  }
  uint32_t st = 0, use = 0;
  const uint32_t n_stages = (uint32_t)a.S, acc_lg = a.acc_lg, acc_mask = (1u << acc_lg) - 1u;
  long long t_full = 0, t_acc = 0;
  const long long t_begin = clock64();
  const int items_bin0 = a.n_src_bin0 > 0 ? wt_n_items<SEQ>(a) : 0;   // items whose source list is the first bin's
  for (int k = 0; k < n_items; ++k) {
    const uint32_t ab = (uint32_t)k & acc_mask;
    const int n_src_k = k < items_bin0 ? a.n_src_bin0 : a.n_src;
    if (k > (int)acc_mask) {
      const long long t0 = clock64();
      mbar_wait(&s.acc_empty[ab], (uint32_t)((k >> acc_lg) - 1) & 1u);
      t_acc += clock64() - t0;
    }
    tc_fence_after();
#pragma unroll
    for (int si = 0; si < NS; ++si) {
      if (si >= n_src_k) break;
      {
        const long long t0 = clock64();
        mbar_wait(&s.full[st], use & 1);
        t_full += clock64() - t0;
      }
      tc_fence_after();
      uint32_t a_base = stages16 + ((st * a.stage_bytes) >> 4);
      uint32_t d = tmem_base + ab * acc_cols;
      int seg = 0;
      for (int m = 0; m < n_mt; ++m) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t b0 = NS <= 2 ? bdesc[NS <= 2 ? si : 0][tap][0] : bbase[si] + (uint32_t)tap * tile16[si];
          umma_f16_split(d, aoff[tap][0] + a_base, d_hi, b0, d_hi, idesc[si], (si > 0 || tap > 0) ? 1u : 0u);
          const uint32_t b1 = NS <= 2 ? bdesc[NS <= 2 ? si : 0][tap][1] : b0 + 2u * blbo16;
          if (n_kk[si] > 1) umma_f16_split(d, aoff[tap][1] + a_base, d_hi, b1, d_hi, idesc[si], 1u);
        }
        d += ncat;
        if (++seg == a.n_seg) { seg = 0; a_base += (uint32_t)(a.Wsm - (a.n_seg - 1) * 128); }
        else a_base += 128u;
      }
      umma_commit(&s.empty[st]);
      if (++st == n_stages) { st = 0; ++use; }
    }
    umma_commit(&s.acc_full[ab]);
  }
  if (a.dbg) {
    a.dbg[blockIdx.x * 8 + 2] = clock64() - t_begin;   // MMA issuer: total
    a.dbg[blockIdx.x * 8 + 3] = t_full;                 // ... waiting for operands
    a.dbg[blockIdx.x * 8 + 4] = t_acc;                  // ... waiting for a free accumulator
  }
}

// ---- tile-flag publisher (time-fused recurrent kernels): one thread ----------------------------------------------
// For every item g of this CTA, in processing order: wait until all epilogue threads have issued the item's stores
// (pub_bar, release.cta / acquire.cta), make them visible device-wide to the generic AND the async proxy (the readers are
// other CTAs' bulk copies), raise the tile's progress flag.  Cumulativity of the release carries the epilogue threads'
// stores; the only memory operations this thread ever has in flight are its own flag stores, so its fences are cheap.
__device__ void wt_publisher(const WtArgs& a, const WtSmem& s) {
  const int n_bins = a.n_bins > 1 ? a.n_bins : 1;
  if (n_bins <= 1 || (a.exp & 2)) return;
  const int n_items = wt_n_items<false>(a);
  int g = 0;
  for (int bin = 0; bin < n_bins; ++bin)
    for (int k = 0; k < n_items; ++k, ++g) {
      mbar_wait(&s.pub_bar[g & (WT_PUB_RING - 1)], (uint32_t)(g / WT_PUB_RING) & 1u);
      fence_proxy_async_global();
      __threadfence();
      tile_flag_set(a.tile_flags + ((int)blockIdx.x + k * (int)gridDim.x), (unsigned int)(bin + 1));
      *s.pub_count = (unsigned int)(g + 1);
    }
}

__device__ __forceinline__ uint32_t bf16_pair(uint32_t mask, int i) {
  return (((mask >> (2 * i)) & 1u) ? 0x3F80u : 0u) | (((mask >> (2 * i + 1)) & 1u) ? 0x3F800000u : 0u);
}
// =================================================================================================
// fp32 tensors of the engine (membranes, currents, spike gradients) use the "c8" layout
//   [image][chunk = C/8][H*W][8 channels]
// so that the epilogue thread that owns one pixel x 16 channels moves them as 16-byte vectors.
// =================================================================================================
__device__ __forceinline__ size_t c8_off(int img, int n_chunks, int chunk, size_t HW, size_t pix) {
  return (((size_t)img * n_chunks + chunk) * HW + pix) * 8;
}
__device__ __forceinline__ void ld8_c8(const float* p, float (&v)[8]) {
  ldg256(p, v);   // one 32-byte sector per lane, one instruction
}
__device__ __forceinline__ void st8_c8(float* p, const float (&v)[8]) {
  stg256(p, v);
}
__device__ __forceinline__ uint32_t nz8_mask(const uint4& a) {   // bit c set when the c-th bf16 is non-zero
  uint32_t m = 0;
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) m |= ((w[i] & 0xFFFFu) ? 1u : 0u) << (2 * i) | ((w[i] >> 16) ? 1u : 0u) << (2 * i + 1);
  return m;
}

// =================================================================================================
// Forward.  Epilogue warp w: TMEM lane quarter q = w & 3 (32 pixels of a segment), 8-channel chunk ch = w >> 2.
// =================================================================================================
template <bool SEQ, int NSEG, bool HARD>
__global__ void __launch_bounds__(WT_THREADS, 1) wt_fwd_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes + a.wblob2_bytes, (uint32_t)a.S * a.stage_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == WT_EPI_WARPS) {
    wt_producer<SEQ>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (elect_one()) wt_mma<SEQ>(a, s, tmem_base);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 2) {
    if (!SEQ && lane == 0) wt_publisher(a, s);
    __syncwarp();
  } else if (warp == WT_PRODUCER2) {
    wt_producer<SEQ>(a, s, 1);
    __syncwarp();
  } else {
    pdl_wait();
    const int q = warp & 3, ch = warp >> 2;
    // per-item parameters live in registers (re-reading the kernel parameter block per element stalls the epilogue)
    const int W = a.W, Wp = a.Wp, N = a.N, nch = N >> 3, n_seg = a.n_seg, T = a.T;
    const bool act = ch * 8 < N;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * W;
    const size_t plane_bytes = (size_t)(a.H + 2) * Wp * 16;
    const uint32_t ncat = 3u * (uint32_t)N, acc_cols = (uint32_t)NSEG * ncat;   // three weight terms side by side
    const int n_items = wt_n_items<SEQ>(a);
    float* const cur_out = a.cur_out;
    const size_t zp_img_stride = a.zp_img_stride;
    float lam[8], oml[8], th[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 pr = s.par[(act ? ch * 8 : 0) + c];
      lam[c] = pr.x; oml[c] = pr.y; th[c] = pr.z;
    }
    long long t_wait = 0;
    const long long t_begin = clock64();
    const uint32_t acc_lg = a.acc_lg, acc_mask = (1u << acc_lg) - 1u;
    if constexpr (SEQ) {
      // Time-fused feed-forward layer: a tile (b, y0) is walked through its T bins with the membrane and the previous
      // spikes in registers and RUNNING output pointers - one bin ahead is a constant step, nothing is re-derived per item.
      const int n_col = a.n_col, tpi = (a.H / a.R) * n_col, n_tiles = a.n_outer * tpi, B = a.B;
      const long long v_bin = (long long)B * nch * (long long)HW * 8;          // floats between bins (c8 membranes)
      const long long zp_bin = (long long)B * (long long)zp_img_stride;        // bytes between bins (spike planes)
      const bool have_v = a.v_out != nullptr, want_last = a.v_last != nullptr || a.z_last != nullptr;
      const bool state_c8 = a.state_c8 != 0;
      uint32_t k = 0;
      for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x) {
        const int b = tile / tpi, rt = tile - b * tpi;
        const int y0 = (n_col > 1 ? rt / n_col : rt) * a.R, x0 = n_col > 1 ? (rt % n_col) * 128 : 0;
        float* vp[NSEG];
        unsigned char* zp[NSEG];
        size_t o_state[NSEG], o_c8[NSEG];
        bool okm[NSEG];
        float vst[NSEG][8], zst[NSEG][8];
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const int y = y0 + m / n_seg, x = x0 + (m % n_seg) * 128 + q * 32 + lane;
          const size_t pix = (size_t)y * W + x;
          okm[m] = act && x < W;
          vp[m] = have_v ? a.v_out + c8_off(b, nch, ch, HW, pix) : nullptr;
          zp[m] = a.zp_out + (size_t)ch * plane_bytes + (size_t)b * zp_img_stride + ((size_t)(y + 1) * Wp + x + 1) * 16;
          o_state[m] = ((size_t)(b * N + ch * 8)) * HW + pix;   // NCHW (state tensors of the caller)
          o_c8[m] = c8_off(b, nch, ch, HW, pix);
#pragma unroll
          for (int c = 0; c < 8; ++c) vst[m][c] = zst[m][c] = 0.f;
          if (state_c8) {
            // streaming mode: the state lives in the engine's own layout between calls - membrane c8 (one 32-byte vector),
            // spikes = this layer's planes of the previous call (the slot this thread overwrites at the last bin)
            if (okm[m] && a.v_init) ld8_c8(a.v_init + c8_off(b, nch, ch, HW, pix), vst[m]);
            if (okm[m] && a.zin_planes) {
              const uint4 zz = *reinterpret_cast<const uint4*>(a.zin_planes + (size_t)b * a.zin_img_stride + (size_t)ch * plane_bytes +
                                                               ((size_t)(y + 1) * Wp + x + 1) * 16);
              const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
              for (int c = 0; c < 8; ++c) zst[m][c] = ((zw[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu) ? 1.f : 0.f;
            }
          } else {
            if (okm[m] && a.v_init) {
#pragma unroll
              for (int c = 0; c < 8; ++c) vst[m][c] = __ldg(a.v_init + o_state[m] + (size_t)c * HW);
            }
            if (okm[m] && a.z_init) {
#pragma unroll
              for (int c = 0; c < 8; ++c) zst[m][c] = __ldg(a.z_init + o_state[m] + (size_t)c * HW);
            }
          }
        }
        for (int t = 0; t < T; ++t, ++k) {
          const uint32_t ab = k & acc_mask;
          {
            const long long t0 = a.dbg ? clock64() : 0;
            mbar_wait(&s.acc_full[ab], (k >> acc_lg) & 1u);
            if (a.dbg) t_wait += clock64() - t0;
          }
          tc_fence_after();
          if (act) {
            uint32_t u0[NSEG][8], u1[NSEG][8], u2[NSEG][8];
#pragma unroll
            for (int m = 0; m < NSEG; ++m) {
              const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(ch * 8);
              tmem_ld8_async(tcol, u0[m]);
              tmem_ld8_async(tcol + (uint32_t)N, u1[m]);
              tmem_ld8_async(tcol + 2u * (uint32_t)N, u2[m]);
            }
            tmem_ld_wait();
#pragma unroll
            for (int m = 0; m < NSEG; ++m) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float cur = (__uint_as_float(u0[m][c]) + __uint_as_float(u1[m][c])) + __uint_as_float(u2[m][c]);   // hi + mid + lo terms
                const float t1 = __fmul_rn(vst[m][c], lam[c]), t3 = __fmul_rn(oml[c], cur);
                float vn;
                if (HARD) vn = __fadd_rn(__fmul_rn(t1, __fsub_rn(1.0f, zst[m][c])), t3);    // spiking_submodules.py:144
                else vn = __fsub_rn(__fadd_rn(t1, t3), __fmul_rn(zst[m][c], th[c]));         // spiking_submodules.py:146
                vst[m][c] = vn;
                zst[m][c] = __fsub_rn(vn, th[c]) > 0.f ? 1.f : 0.f;                          // spiking_util.py:21
              }
              if (okm[m]) {
                uint4 zz;   // {0, 1} are exact in bf16: one packed convert per channel pair
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz.x) : "f"(zst[m][1]), "f"(zst[m][0]));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz.y) : "f"(zst[m][3]), "f"(zst[m][2]));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz.z) : "f"(zst[m][5]), "f"(zst[m][4]));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz.w) : "f"(zst[m][7]), "f"(zst[m][6]));
                *reinterpret_cast<uint4*>(zp[m]) = zz;
                if (have_v) st8_c8(vp[m], vst[m]);
                if (t == T - 1 && want_last) {   // the caller-visible state [2,B,C,H,W] after the window
                  if (state_c8) {
                    st8_c8(a.v_last + (o_c8[m]), vst[m]);   // (the spikes of the last bin are already in the planes)
                  } else if (a.v_last) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) a.v_last[o_state[m] + (size_t)c * HW] = vst[m][c];
                  }
                  if (a.z_last && !state_c8) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) a.z_last[o_state[m] + (size_t)c * HW] = zst[m][c];
                  }
                }
              }
              zp[m] += zp_bin;
              if (have_v) vp[m] += v_bin;
            }
          }
          tc_fence_before();
          mbar_arrive(&s.acc_empty[ab]);
        }
      }
    } else {
      // Step mode (ConvLIFRecurrent): one item per tile and bin.  The membrane and the spikes entering the bin come from
      // memory (written by this very thread one bin earlier, or by the previous launch); they are requested ONE ITEM AHEAD,
      // before the wait for the current accumulator, so their latency hides behind the current item's arithmetic.
      // Time-fused launches (n_bins > 1) walk all bins of the window and publish a per-tile progress flag after every item.
      const int n_bins = a.n_bins > 1 ? a.n_bins : 1;
      const bool multi = n_bins > 1;
      const int total = n_bins * n_items;
      const bool ahead_ok = (!multi || n_items >= 2) && !(a.exp & 1);   // else the next item's inputs are this item's outputs
      const float4* const par = s.par + (act ? ch * 8 : 0);
      const long long bin_v_stride = a.bin_v_stride, bin_zp_stride = a.bin_zp_stride;
      const int bin_v_mask = a.bin_v_mask;
      const size_t zin_img_stride = a.zin_img_stride;

      float vst[NSEG][8], vnx[NSEG][8];
      uint4 zq[NSEG], znx[NSEG];
      // inputs of item (b, y0) of bin `bin` -> v, z  (zeros where there is no state)
      auto request = [&](int b, int y0, int x0, int bin, float (&v)[NSEG][8], uint4 (&z)[NSEG]) {
        const float* v_prev = bin == 0 ? a.v_prev : a.v_out + (size_t)((bin - 1) & bin_v_mask) * bin_v_stride;
        const bool nchw = bin == 0 && a.v_prev_nchw;
        const unsigned char* zin = a.zin_planes ? a.zin_planes + (long long)bin * bin_zp_stride : nullptr;
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const int y = y0 + m / n_seg, x = x0 + (m % n_seg) * 128 + q * 32 + lane;
          const bool ok = act && x < W;
          const size_t pix = (size_t)y * W + x;
#pragma unroll
          for (int c = 0; c < 8; ++c) v[m][c] = 0.f;
          z[m] = make_uint4(0u, 0u, 0u, 0u);
          if (ok && v_prev) {
            if (nchw) {
              const size_t o = ((size_t)(b * N + ch * 8)) * HW + pix;   // the caller's NCHW state tensor
#pragma unroll
              for (int c = 0; c < 8; ++c) v[m][c] = __ldg(v_prev + o + (size_t)c * HW);
            } else if (multi) {
              ldg256_coherent(v_prev + c8_off(b, nch, ch, HW, pix), v[m]);   // written by this thread one bin ago
            } else {
              ld8_c8(v_prev + c8_off(b, nch, ch, HW, pix), v[m]);
            }
          }
          if (ok && zin) {
            const uint4* zsrc = reinterpret_cast<const uint4*>(zin + (size_t)b * zin_img_stride + (size_t)ch * plane_bytes +
                                                               ((size_t)(y + 1) * Wp + x + 1) * 16);
            z[m] = multi ? ldg128_coherent(zsrc) : __ldg(zsrc);
          }
        }
      };

      // Time-fused launches: after the stores of an item every epilogue thread arrives on pub_bar[g % 8]; the publisher warp
      // (wt_publisher) raises the tile's progress flag from there.  The epilogue threads themselves never execute a fence:
      // a membar / proxy fence waits for the thread's outstanding loads and stores, i.e. for the inputs requested one item
      // ahead (measured: +1300 cycles per item with the fences here).
      const bool no_flags = (a.exp & 2) != 0;   // timing experiments only: results are wrong
      ItemIter<false> it, nx;
      it.init(a);
      nx.init(a);
      int bin = 0, nbin = 0, nk = 0;
      // one item: accumulator + (v, z) of set `vs / zs` -> new membrane and spikes, stored; the accumulator is handed back
      auto process = [&](int g, float (&vs)[NSEG][8], uint4 (&zs)[NSEG]) {
        const uint32_t ab = (uint32_t)g & acc_mask;
        const bool last = bin == n_bins - 1 && (a.v_last != nullptr || a.z_last != nullptr) && !a.state_c8;
        float* const v_out = a.v_out ? a.v_out + (size_t)(bin & bin_v_mask) * bin_v_stride : nullptr;
        unsigned char* const zp_out = a.zp_out + (long long)bin * bin_zp_stride + (size_t)ch * plane_bytes;
        {
          const long long t0 = a.dbg ? clock64() : 0;
          mbar_wait(&s.acc_full[ab], (uint32_t)(g >> acc_lg) & 1u);
          if (a.dbg) t_wait += clock64() - t0;
        }
        tc_fence_after();
        uint4 zz[NSEG];   // the new spikes as packed bf16 ({0, 1} are exact: one packed convert per channel pair)
        if (act) {
          uint32_t u0[NSEG][8], u1[NSEG][8], u2[NSEG][8];
#pragma unroll
          for (int m = 0; m < NSEG; ++m) {
            const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(ch * 8);
            tmem_ld8_async(tcol, u0[m]);
            tmem_ld8_async(tcol + (uint32_t)N, u1[m]);
            tmem_ld8_async(tcol + 2u * (uint32_t)N, u2[m]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int m = 0; m < NSEG; ++m) {
            const uint32_t zw[4] = {zs[m].x, zs[m].y, zs[m].z, zs[m].w};
            float zn[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 pr = par[c];   // (lam, 1 - lam, theta, .): broadcast read, keeps 24 registers free
              const float zin_c = ((zw[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu) ? 1.f : 0.f;
              const float cur = (__uint_as_float(u0[m][c]) + __uint_as_float(u1[m][c])) + __uint_as_float(u2[m][c]);   // hi + mid + lo terms
              const float t1 = __fmul_rn(vs[m][c], pr.x), t3 = __fmul_rn(pr.y, cur);
              float vn;
              if (HARD) vn = __fadd_rn(__fmul_rn(t1, __fsub_rn(1.0f, zin_c)), t3);    // spiking_submodules.py:144
              else vn = __fsub_rn(__fadd_rn(t1, t3), __fmul_rn(zin_c, pr.z));         // spiking_submodules.py:146
              vs[m][c] = vn;
              zn[c] = __fsub_rn(vn, pr.z) > 0.f ? 1.f : 0.f;                          // spiking_util.py:21
            }
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz[m].x) : "f"(zn[1]), "f"(zn[0]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz[m].y) : "f"(zn[3]), "f"(zn[2]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz[m].z) : "f"(zn[5]), "f"(zn[4]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(zz[m].w) : "f"(zn[7]), "f"(zn[6]));
          }
#pragma unroll
          for (int m = 0; m < NSEG; ++m) {
            const int y = it.y0 + m / n_seg, x = it.x0 + (m % n_seg) * 128 + q * 32 + lane;
            if (x < W) {
              const size_t pix = (size_t)y * W + x;
              *reinterpret_cast<uint4*>(zp_out + (size_t)it.img * zp_img_stride + ((size_t)(y + 1) * Wp + x + 1) * 16) = zz[m];
              if (v_out) st8_c8(v_out + c8_off(it.img, nch, ch, HW, pix), vs[m]);
              if (last) {   // the caller-visible state [2,B,C,H,W] after the window
                const size_t o = ((size_t)(it.b * N + ch * 8)) * HW + pix;
                if (a.v_last) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) a.v_last[o + (size_t)c * HW] = vs[m][c];
                }
                if (a.z_last) {
                  const uint32_t zw[4] = {zz[m].x, zz[m].y, zz[m].z, zz[m].w};
#pragma unroll
                  for (int c = 0; c < 8; ++c) a.z_last[o + (size_t)c * HW] = ((zw[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu) ? 1.f : 0.f;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&s.acc_empty[ab]);
        if (multi && !no_flags) {
          // back-pressure: the ring slot of item g was last used by item g - 8, which the publisher must have consumed
          if (tid == 0 && g >= WT_PUB_RING) {
            while (*s.pub_count < (unsigned int)(g - WT_PUB_RING + 1)) __nanosleep(20);
          }
          mbar_arrive(&s.pub_bar[g & (WT_PUB_RING - 1)]);   // release.cta: orders this thread's plane / membrane stores
        }
      };
      // position of the item after `it`, and its inputs on their way into the OTHER register set
      auto advance = [&](int g, float (&vs)[NSEG][8], uint4 (&zs)[NSEG]) -> bool {
        if (g + 1 >= total) return false;
        if (++nk == n_items) { nk = 0; ++nbin; nx.init(a); } else nx.next();
        if (ahead_ok) request(nx.b, nx.y0, nx.x0, nbin, vs, zs);
        return true;
      };
      // Two items per trip with the register sets swapping roles (no copies: a copy would be the first use of the loads in
      // flight and stall on them at the end of every item - what the first version of this loop did).
      if (total > 0) request(it.b, it.y0, it.x0, 0, vst, zq);
      for (int g = 0; g < total;) {
        bool have_next = advance(g, vnx, znx);
        process(g, vst, zq);
        if (!have_next) break;
        if (!ahead_ok) request(nx.b, nx.y0, nx.x0, nbin, vnx, znx);   // (the next item's inputs are this item's outputs)
        it = nx; bin = nbin; ++g;
        have_next = advance(g, vst, zq);
        process(g, vnx, znx);
        if (!have_next) break;
        if (!ahead_ok) request(nx.b, nx.y0, nx.x0, nbin, vst, zq);
        it = nx; bin = nbin; ++g;
      }
    }
    if (a.dbg && tid == 0) {
      a.dbg[blockIdx.x * 8 + 5] = clock64() - t_begin;   // epilogue warp 0: total
      a.dbg[blockIdx.x * 8 + 6] = t_wait;                 // ... waiting for the accumulator
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
  const WtSmem s = wt_smem(smem, a.wblob_bytes + a.wblob2_bytes, (uint32_t)a.S * a.stage_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    wt_producer<false>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (elect_one()) wt_mma<false>(a, s, tmem_base);
    __syncwarp();
  } else if (warp == WT_PRODUCER2) {
    wt_producer<false>(a, s, 1);
    __syncwarp();
  } else if (warp >= WT_EPI_WARPS + 2) {
    // (the publisher warp has no work in this kernel)
  } else {
    pdl_wait();
    const int q = warp & 3, ch = warp >> 2;
    const bool act = ch * 8 < a.N;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * a.W;
    const int n_mt = a.R * a.n_seg, nch = a.N >> 3;
    const uint32_t ncat = 2u * (uint32_t)a.N, acc_cols = (uint32_t)n_mt * ncat;   // [g*w_hi | g_hi*w_lo]
    const int n_items = wt_n_items<false>(a);
    const int W = a.W, N = a.N, n_seg = a.n_seg;
    float* const g_x = a.g_x;
    long long t_wait = 0;
    const long long t_begin = clock64();
    const uint32_t acc_lg = a.acc_lg, acc_mask = (1u << acc_lg) - 1u;
    ItemIter<false> p;
    p.init(a);
    for (int k = 0; k < n_items; ++k, p.next()) {
      const uint32_t ab = (uint32_t)k & acc_mask;
      {
        const long long t0 = clock64();
        mbar_wait(&s.acc_full[ab], (uint32_t)(k >> acc_lg) & 1u);
        t_wait += clock64() - t0;
      }
      tc_fence_after();
      if (act) {
        int r = 0, seg = 0;
        for (int m = 0; m < n_mt; ++m) {
          const int y = p.y0 + r, x = seg * 128 + q * 32 + lane;
          uint32_t u0[8], u1[8];
          const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(ch * 8);
          tmem_ld8_async(tcol, u0);
          tmem_ld8_async(tcol + (uint32_t)N, u1);
          tmem_ld_wait();
          float acc[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = __uint_as_float(u0[c]) + __uint_as_float(u1[c]);
          if (x < W) st8_c8(g_x + c8_off(p.img, nch, ch, HW, (size_t)y * W + x), acc);
          if (++seg == n_seg) { seg = 0; ++r; }
        }
      }
      tc_fence_before();
      mbar_arrive(&s.acc_empty[ab]);
    }
    if (a.dbg && tid == 0) {
      a.dbg[blockIdx.x * 8 + 5] = clock64() - t_begin;
      a.dbg[blockIdx.x * 8 + 6] = t_wait;
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
// (I is recovered from v_t, v_in and z_in as in pw_seq_kernel: the input current is not stored.)
// =================================================================================================
// walks the 128-pixel segments of this CTA's tiles in launch order without divisions (step-mode kernels)
struct SegIter {
  int b, y0, r, seg, k, H, R, n_seg, step_rows;
  __device__ __forceinline__ void init(const WtArgs& a) {
    const int tpi = a.H / a.R;
    b = (int)blockIdx.x / tpi;
    y0 = ((int)blockIdx.x - b * tpi) * a.R;
    r = seg = k = 0;
    H = a.H; R = a.R; n_seg = a.n_seg; step_rows = (int)gridDim.x * a.R;
  }
  __device__ __forceinline__ void next() {
    if (++seg < n_seg) return;
    seg = 0;
    if (++r < R) return;
    r = 0;
    ++k;
    y0 += step_rows;
    while (y0 >= H) { y0 -= H; ++b; }
  }
};

// MODE 0: one bin per launch.  MODE 1: one bin per launch, the FIRST bin of the window (t = 0) - v_in / z_in come from the
// caller's NCHW state tensors (or are zero).  MODE 2: time-fused - ONE cooperative launch walks the window backwards
// (bin j of the launch is t = T-1-j; per-bin strides in WtArgs), the g_I planes of bin j being the recurrent operand of
// bin j+1 under the per-tile progress flags (wt_publisher); d lam / d theta partial sums are carried over all bins.
template <int SG, bool HARD, int MODE>
__global__ void __launch_bounds__(WT_THREADS, 1) wt_recbwd_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes + a.wblob2_bytes, (uint32_t)a.S * a.stage_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    if (a.has_gz) wt_producer<false>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (a.has_gz && elect_one()) wt_mma<false, WT_MAX_SRC>(a, s, tmem_base);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 2) {
    if (MODE == 2 && lane == 0) wt_publisher(a, s);
    __syncwarp();
  } else if (warp == WT_PRODUCER2) {
    if (a.has_gz) wt_producer<false>(a, s, 1);
    __syncwarp();
  } else {
    pdl_wait();
    const int q = warp & 3, ch = warp >> 2;
    // kernel parameters used per segment live in registers (re-reading them from the constant bank stalls the epilogue)
    const int W = a.W, Wp = a.Wp, N = a.N, nch = N >> 3;
    const bool act = ch * 8 < N, has_gz = a.has_gz != 0;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * W;
    const size_t plane_bytes = (size_t)(a.H + 2) * Wp * 16;
    const int n_mt = a.R * a.n_seg;
    const uint32_t ncat = 2u * (uint32_t)N, acc_cols = (uint32_t)n_mt * ncat;
    const int n_sub = wt_n_items<false>(a) * n_mt;   // 128-pixel segments this CTA processes, in order
    const float width = a.width;
    const float* const g_out = a.g_out;
    const bool l2_pf = a.l2_prefetch != 0;
    float* g_v = a.g_v;
    const size_t gp_img_stride = a.gp_img_stride, gp_term_stride = a.gp_term_stride;
    const float4* par = s.par + (act ? ch * 8 : 0);   // (lam, 1 - lam, theta, 1 / (1 - lam)) per channel, read where used
    float s_lam[8], s_th[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) s_lam[c] = s_th[c] = 0.f;
    long long t_wait = 0;
    const long long t_begin = clock64();
    const uint32_t acc_lg = a.acc_lg, acc_mask = (1u << acc_lg) - 1u;
    // Inputs of a segment (g_out, v_t, v_in, g_v: 128 B per thread) are plain 256-bit loads at the top of the segment; a
    // cp.async prefetch through thread-private shared-memory slots was measured slower (38 vs 35 us per bin) and is gone.
    const int n_bins = (MODE == 2 && a.n_bins > 1) ? a.n_bins : 1;
    const int items_per_bin = wt_n_items<false>(a);
    for (int bin = 0; bin < n_bins; ++bin) {
    // this bin's views (bin 0 = the launch arguments; the strides are negative: the window is walked backwards)
    const float* const v_t = a.v_t + (long long)bin * a.bin_v_stride;
    const bool t0 = MODE == 1 || (MODE == 2 && bin == n_bins - 1);   // bin t = 0: state entering the window is the caller's
    const float* const v_in = (MODE == 2) ? (t0 ? a.v_init : a.v_in + (long long)bin * a.bin_v_stride) : a.v_in;
    const bool first_step = MODE == 2 ? bin == 0 : a.first_step != 0;
    unsigned char* const gp_out = a.gp_out + (long long)bin * a.bin_zp_stride;
    SegIter cur;
    cur.init(a);
    for (int j = 0; j < n_sub; ++j) {
      const int b = cur.b, y = cur.y0 + cur.r, x = cur.seg * 128 + q * 32 + lane, m = cur.r * cur.n_seg + cur.seg;
      const int k = bin * items_per_bin + cur.k;   // running item number of this CTA (accumulator ring, publisher ring)
      const uint32_t ab = (uint32_t)k & acc_mask;
      const bool ok = x < W;
      const size_t pix = (size_t)y * W + x;
      const size_t co = c8_off(b, nch, ch, HW, pix);
      float go[8], vt[8], vin[8], gv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) go[c] = vt[c] = vin[c] = gv[c] = 0.f;
      if (l2_pf && act && j + 1 < n_sub) {   // the next segment's streams (DRAM -> L2) while this one is processed
        SegIter nxs = cur;
        nxs.next();
        const int xn = nxs.seg * 128 + q * 32 + lane;
        if (xn < W) {
          const size_t con = c8_off(nxs.b, nch, ch, HW, (size_t)(nxs.y0 + nxs.r) * W + xn);
          if (g_out) prefetch_l2(g_out + con);
          prefetch_l2(v_t + con);
          if (!t0) prefetch_l2(v_in + con);
        }
      }
      if (act && ok) {
        if (g_out) ld8_c8(g_out + co, go);   // NULL: the spike gradient from the layer above arrives in the accumulator
        ld8_c8(v_t + co, vt);
        if (!t0) ld8_c8(v_in + co, vin);
        if (!first_step) {   // written by the previous launch / bin, rewritten below by this thread only
          if (MODE == 2) ldg256_coherent(g_v + co, gv);
          else ld8_c8(g_v + co, gv);
        }
      }
      if (act) {
        float zin[8];
        if (t0) {   // window-initial state of the caller (NCHW) or zeros
          const size_t o = ((size_t)(b * N + ch * 8)) * HW + pix;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            vin[c] = (ok && v_in) ? __ldg(v_in + o + (size_t)c * HW) : 0.f;
            zin[c] = (ok && a.z_init) ? __ldg(a.z_init + o + (size_t)c * HW) : 0.f;
          }
        }
        float acc[8];
        if (has_gz) {
          if (m == 0) {
            const long long t0c = a.dbg ? clock64() : 0;
            mbar_wait(&s.acc_full[ab], (uint32_t)(k >> acc_lg) & 1u);
            if (a.dbg) t_wait += clock64() - t0c;
            tc_fence_after();
          }
          uint32_t u0[8], u1[8];
          const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(ch * 8);
          tmem_ld8_async(tcol, u0);
          tmem_ld8_async(tcol + (uint32_t)N, u1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)   // pixels past the row end accumulate whatever the operand read found: mask them
            acc[c] = ok ? __uint_as_float(u0[c]) + __uint_as_float(u1[c]) : 0.f;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        }
        uint32_t hi[4], lo[4];
        float gvn[8], gI[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 pr = par[c];
          const float z_in = t0 ? zin[c] : ((__fsub_rn(vin[c], pr.z) > 0.f) ? 1.f : 0.f);
          const float gs = (go[c] + acc[c]) * surrogate_fast<SG>(vt[c] - pr.z, width);
          const float gvv = gv[c] + gs;
          gI[c] = gvv * pr.y;
          if (HARD) {
            const float omz = 1.0f - z_in;
            gvn[c] = gvv * pr.x * omz;
            s_lam[c] += gvv * (vin[c] * omz - vt[c]);   // (a - I) (1 - lam), see pw_seq_kernel
            s_th[c] -= gs;
          } else {
            gvn[c] = gvv * pr.x;
            s_lam[c] += gvv * (vin[c] - vt[c] - z_in * pr.z);
            s_th[c] -= gs + gvv * z_in;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) split_bf16_pair(gI[2 * c], gI[2 * c + 1], hi[c], lo[c]);
        if (ok) {
          st8_c8(g_v + co, gvn);
          unsigned char* gp = gp_out + (size_t)b * gp_img_stride + (size_t)ch * plane_bytes + ((size_t)(y + 1) * Wp + x + 1) * 16;
          *reinterpret_cast<uint4*>(gp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(gp + gp_term_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      } else if (has_gz && m == 0) {
        mbar_wait(&s.acc_full[ab], (uint32_t)(k >> acc_lg) & 1u);   // idle channel groups still follow the accumulator phases
      }
      if (has_gz && m == n_mt - 1) {
        tc_fence_before();
        mbar_arrive(&s.acc_empty[ab]);
        if (MODE == 2 && n_bins > 1 && !(a.exp & 2)) {   // hand the finished item to the publisher warp (see wt_fwd_kernel)
          if (tid == 0 && k >= WT_PUB_RING) {
            while (*s.pub_count < (unsigned int)(k - WT_PUB_RING + 1)) __nanosleep(20);
          }
          mbar_arrive(&s.pub_bar[k & (WT_PUB_RING - 1)]);
        }
      }
      cur.next();
    }
    }
    if (a.dbg && tid == 0) {
      a.dbg[blockIdx.x * 8 + 5] = clock64() - t_begin;
      a.dbg[blockIdx.x * 8 + 6] = t_wait;
    }
    // per-warp partial sums of dlam / dtheta -> shared scratch [warp][2][8]
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float l = warp_sum(s_lam[c] * par[c].w), t = warp_sum(s_th[c]);
      if (lane == 0) {
        s.red[(warp * 2 + 0) * 8 + c] = l;
        s.red[(warp * 2 + 1) * 8 + c] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 2 * a.N) {
    const int which = tid / a.N, co = tid % a.N;
    const int ch = co >> 3, c = co & 7;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += s.red[((ch * 4 + q) * 2 + which) * 8 + c];
    a.part[(size_t)blockIdx.x * 2 * a.N + tid] = t;
  }
  if (warp == 0) tmem_dealloc(tmem_base, wt_tmem_cols(a));
}

// =================================================================================================
// Data gradient of layer l+1 fused with the time-fused pointwise BPTT of the feed-forward layer l below it.
// A CTA owns a row tile for ALL T bins and walks them backwards (t = T-1 .. 0):
//   MMA      : g_out[t] = conv^T(g_I^{l+1}[t], W_ff^{l+1})                     (hi planes x [w_hi | w_lo], lo planes x w_hi)
//   epilogue : gs = g_out * sg(v[t] - theta);  gv = carry + gs;  g_I^{l}[t] = gv * (1 - lam)  -> bf16 hi/lo planes
//              carry = gv * lam * (1 - z_in)  (hard reset) / gv * lam  (soft)  stays in registers across the bins,
//              d lam / d theta partial sums as in pw_seq_kernel.
// The spike gradient g_out of layer l ([T*B,C,H,W] fp32) is never written to or read from memory.
// Measured (B200, C=32, 128x128, B=8, T=10): this plain epilogue (v[t-1] requested at the top of the item that consumes it)
// runs in 143 us.  Requesting the membranes one item ahead - a rolling window v[t], v[t-1], v[t-2] in registers with running
// pointers - was tried twice (round 1: 161 us, 180 B of spills; round 2: 178 us, 48 B of spills): slower both times, so the
// exposed load is not what bounds this epilogue (profiles/r2_dgpw_source.md has the per-instruction stall picture).
// =================================================================================================
template <int SG, bool HARD, int NSEG>
__global__ void __launch_bounds__(WT_THREADS, 1) wt_dgpw_kernel(const __grid_constant__ WtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const WtSmem s = wt_smem(smem, a.wblob_bytes + a.wblob2_bytes, (uint32_t)a.S * a.stage_bytes);
  const uint32_t tmem_base = wt_prologue(a, s);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == WT_EPI_WARPS) {
    wt_producer<true>(a, s);
    __syncwarp();
  } else if (warp == WT_EPI_WARPS + 1) {
    if (elect_one()) wt_mma<true>(a, s, tmem_base);
    __syncwarp();
  } else if (warp == WT_PRODUCER2) {
    wt_producer<true>(a, s, 1);
    __syncwarp();
  } else if (warp >= WT_EPI_WARPS + 2) {
    // (the publisher warp has no work in this kernel)
  } else {
    pdl_wait();

    const int q = warp & 3, ch = warp >> 2;
    const int W = a.W, Wp = a.Wp, N = a.N, nch = N >> 3, n_seg = a.n_seg, T = a.T, B = a.B;
    const bool act = ch * 8 < N;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const size_t HW = (size_t)a.H * W;
    const size_t plane_bytes = (size_t)(a.H + 2) * Wp * 16;
    const uint32_t ncat = 2u * (uint32_t)N, acc_cols = (uint32_t)NSEG * ncat;
    const int n_items = wt_n_items<true>(a);
    const float width = a.width;
    const float* const v = a.v_t;          // membranes of layer l, all bins (c8)
    const bool use_aux = a.aux != nullptr;
    const bool l2_pf = a.l2_prefetch != 0 && !use_aux;
    uint32_t aux_slot = 0, aux_phase = 0;
    unsigned char* const gp_out = a.gp_out + (size_t)ch * plane_bytes;
    const size_t gp_img_stride = a.gp_img_stride, gp_term_stride = a.gp_term_stride;
    const float4* par = s.par + (act ? ch * 8 : 0);
    float s_lam[8], s_th[8], carry[NSEG][8], v_cur[NSEG][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) s_lam[c] = s_th[c] = 0.f;
    const uint32_t acc_lg = a.acc_lg, acc_mask = (1u << acc_lg) - 1u;
    ItemIter<true> it;
    it.init(a);
    for (int k = 0; k < n_items; ++k, it.next()) {
      const uint32_t ab = (uint32_t)k & acc_mask;
      const int t = T - 1 - it.t;            // bins are walked backwards
      const int img = t * B + it.b;
      // this bin's inputs of the epilogue: v[t-1] (or the window-initial state); v[t] is carried from the previous item
      float vin[NSEG][8], zin[NSEG][8];
#pragma unroll
      for (int m = 0; m < NSEG; ++m) {
        const int y = it.y0 + m / n_seg, x = (m % n_seg) * 128 + q * 32 + lane;
        const bool ok = act && x < W;
        const size_t pix = (size_t)y * W + x;
#pragma unroll
        for (int c = 0; c < 8; ++c) vin[m][c] = zin[m][c] = 0.f;
        if (it.t == 0) {   // first item of a tile: last bin of the window
#pragma unroll
          for (int c = 0; c < 8; ++c) carry[m][c] = v_cur[m][c] = 0.f;
          if (ok) ld8_c8(v + c8_off(img, nch, ch, HW, pix), v_cur[m]);
        }
        if (ok) {
          // the membrane the NEXT item will ask for (v[t-2]) starts its way from DRAM to L2 now
          if (l2_pf && t > 1) prefetch_l2(v + c8_off(img - 2 * B, nch, ch, HW, pix));
          if (t > 0) {
            if (!use_aux) ld8_c8(v + c8_off(img - B, nch, ch, HW, pix), vin[m]);
          } else {
            const size_t o = ((size_t)(it.b * N + ch * 8)) * HW + pix;   // NCHW state tensors of the caller
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              vin[m][c] = a.v_init ? __ldg(a.v_init + o + (size_t)c * HW) : 0.f;
              zin[m][c] = a.z_init ? __ldg(a.z_init + o + (size_t)c * HW) : 0.f;
            }
          }
        }
      }
      mbar_wait(&s.acc_full[ab], (uint32_t)(k >> acc_lg) & 1u);
      tc_fence_after();
      if (use_aux && t > 0) {   // v[t-1] of this tile was staged by the producer (all epilogue threads keep the ring in step)
        mbar_wait(&s.aux_full[aux_slot], aux_phase);
        if (act) {
          const unsigned char* base = s.aux + (size_t)aux_slot * a.aux_stage_bytes + (size_t)ch * a.aux_chunk_bytes;
#pragma unroll
          for (int m = 0; m < NSEG; ++m) {
            const int x = (m % n_seg) * 128 + q * 32 + lane;
            if (x < W) {
              const float4* p4 = reinterpret_cast<const float4*>(base + ((size_t)(m / n_seg) * W + x) * 32);
              const float4 lo4 = p4[0], hi4 = p4[1];
              vin[m][0] = lo4.x; vin[m][1] = lo4.y; vin[m][2] = lo4.z; vin[m][3] = lo4.w;
              vin[m][4] = hi4.x; vin[m][5] = hi4.y; vin[m][6] = hi4.z; vin[m][7] = hi4.w;
            }
          }
        }
        mbar_arrive(&s.aux_empty[aux_slot]);
        if (++aux_slot == (uint32_t)a.aux_slots) { aux_slot = 0; aux_phase ^= 1u; }
      }
      if (act) {
        uint32_t u0[NSEG][8], u1[NSEG][8];
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const uint32_t tcol = tmem_base + t_lane + ab * acc_cols + (uint32_t)m * ncat + (uint32_t)(ch * 8);
          tmem_ld8_async(tcol, u0[m]);
          tmem_ld8_async(tcol + (uint32_t)N, u1[m]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int m = 0; m < NSEG; ++m) {
          const int y = it.y0 + m / n_seg, x = (m % n_seg) * 128 + q * 32 + lane;
          const bool ok = x < W;
          uint32_t hi[4], lo[4];
          float gI[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 pr = par[c];
            const float go = ok ? __uint_as_float(u0[m][c]) + __uint_as_float(u1[m][c]) : 0.f;   // mask out-of-row garbage
            const float z_in = t > 0 ? ((__fsub_rn(vin[m][c], pr.z) > 0.f) ? 1.f : 0.f) : zin[m][c];
            const float gs = go * surrogate_fast<SG>(v_cur[m][c] - pr.z, width);
            const float gv = carry[m][c] + gs;
            gI[c] = gv * pr.y;
            if (HARD) {
              const float omz = 1.0f - z_in;
              carry[m][c] = gv * pr.x * omz;
              s_lam[c] += gv * (vin[m][c] * omz - v_cur[m][c]);
              s_th[c] -= gs;
            } else {
              carry[m][c] = gv * pr.x;
              s_lam[c] += gv * (vin[m][c] - v_cur[m][c] - z_in * pr.z);
              s_th[c] -= gs + gv * z_in;
            }
            v_cur[m][c] = vin[m][c];
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) split_bf16_pair(gI[2 * c], gI[2 * c + 1], hi[c], lo[c]);
          if (ok) {
            unsigned char* gp = gp_out + (size_t)img * gp_img_stride + ((size_t)(y + 1) * Wp + x + 1) * 16;
            *reinterpret_cast<uint4*>(gp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(gp + gp_term_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&s.acc_empty[ab]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float l = warp_sum(s_lam[c] * par[c].w), t = warp_sum(s_th[c]);
      if (lane == 0) {
        s.red[(warp * 2 + 0) * 8 + c] = l;
        s.red[(warp * 2 + 1) * 8 + c] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 2 * a.N) {
    const int which = tid / a.N, co = tid % a.N;
    const int ch = co >> 3, c = co & 7;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += s.red[((ch * 4 + q) * 2 + which) * 8 + c];
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


int wt_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

bool wt_plan(int H, int W, int max_chunks_per_stage, int N, uint32_t wblob_bytes, bool seq_state, int w_terms, bool tall, int* R_out,
             int* S_out,
             uint32_t* sub_bytes, uint32_t* chunk_stride, uint32_t* stage_bytes, int only_R, bool column_tiles) {
  // column_tiles: the tile is (R rows) x (one 128-pixel segment) with a one-pixel halo, loaded row by row (W = n_col * 128)
  const int Wp = column_tiles ? 130 : W + 2, n_seg = column_tiles ? 1 : ceil_div(W, 128);
  if (column_tiles && (W % 128 != 0 || W <= 128)) return false;
  if (N > 32) return false;   // one 16-channel group per epilogue warp pair
  const size_t budget = (size_t)227 * 1024 - WT_HDR - WT_TAIL - align_up(wblob_bytes, 128);
  const int forced_R = only_R ? only_R : env_int("SNNFLOW_WT_R", 0), forced_S = env_int("SNNFLOW_WT_S", 0);
  // measured on B200 (profiles/): the LIF epilogues run best on one-row tiles (more, shorter pipeline items per CTA),
  // the data gradient (cheap epilogue, MMA-bound) on the tallest tile that still leaves three stages
  int best_R = 0, best_S = 0;
  for (int i = 0; i < 3; ++i) {
    const int R = tall ? (4 >> i) : (1 << i);
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

static size_t wt_smem_bytes(const WtArgs& a, size_t extra = 0) {
  return (size_t)WT_HDR + align_up((size_t)a.wblob_bytes + a.wblob2_bytes, 128) + (size_t)a.S * a.stage_bytes + WT_TAIL +
         (a.aux ? (size_t)a.aux_slots * a.aux_stage_bytes : 0) + extra;
}

template <typename K>
static int wt_launch(K kernel, const void* key, const WtArgs& a_in, cudaStream_t st, const char* what, size_t extra_smem = 0) {
  const size_t smem = wt_smem_bytes(a_in, extra_smem);
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
  WtArgs a = a_in;
  static const int l2_pf = env_int("SNNFLOW_L2_PREFETCH", 1);
  a.l2_prefetch = l2_pf;
  static const int exp_bits = env_int("SNNFLOW_EXP", 0);   // experiment switches (see wt_fwd_kernel); 0 in production
  a.exp = exp_bits;
  // Two producer warps where the producer is the bottleneck: the sequence-mode forward on column tiles (12 row copies per
  // tile: eval at 256x256, 31.8k -> 34.7k frames/s).  Elsewhere one warp keeps up and the second one costs 1 - 10 % (measured
  // per kernel at the training shape, profiles/r2_experiments.md).  SNNFLOW_PRODUCERS=1|2 forces either (read per launch).
  {
    const int forced = env_int("SNNFLOW_PRODUCERS", 0);
    a.n_prod = forced ? forced : (a_in.n_prod > 0 ? a_in.n_prod : 1);
  }
  if (a.n_col < 1) a.n_col = 1;
  a.Wsm = a.n_col > 1 ? 130 : a.Wp;   // shared-memory row pitch of an operand tile, in pixel slots
  const int n_tiles = a.n_outer * (a.H / a.R) * a.n_col;
  {   // accumulator ring: as many buffers (2 or 4) as the 512 TMEM columns hold
    const uint32_t acc_cols = (uint32_t)(a.R * a.n_seg * a.N) * a.src[0].w_terms;
    static const int max_lg = env_int("SNNFLOW_WT_ACC_LG", 2);
    a.acc_lg = (max_lg >= 2 && 4u * acc_cols <= 512u) ? 2u : 1u;
  }
  if (env_int("SNNFLOW_WT_TIMING", 0)) {   // debug: where does each role of the pipeline wait? (synchronises)
    static long long* dbg = nullptr;
    const int grid = wt_grid(n_tiles);
    if (!dbg) cudaMalloc(&dbg, sizeof(long long) * 8 * 1024);
    cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * grid, st);
    WtArgs b = a;
    b.dbg = dbg;
    if (b.n_bins > 1) cudaMemsetAsync(b.tile_flags, 0, sizeof(unsigned int) * (size_t)n_tiles, st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    kernel<<<grid, WT_THREADS, smem, st>>>(b);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ev_ms = 0.f;
    cudaEventElapsedTime(&ev_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::vector<long long> h(8 * grid);
    cudaMemcpy(h.data(), dbg, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost);
    double avg[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) avg[j] += (double)h[i * 8 + j] / grid;
    fprintf(stderr, "[wt-timing] %s R=%d S=%d n_src=%d N=%d items/CTA=%.1f | producer %.0f (wait-empty %.0f) | mma %.0f (wait-full %.0f, wait-acc %.0f) | epi %.0f (wait-acc-full %.0f) | prologue %.0f | %.1f us\n",
            what, a.R, a.S, a.n_src, a.N, (double)n_tiles * (a.T > 1 ? a.T : 1) / grid, avg[0], avg[1], avg[2], avg[3], avg[4], avg[5], avg[6], avg[7], ev_ms * 1e3);
    return check_launch(what);
  }
  if (a.n_bins > 1) {   // time-fused over the bins: CTAs wait for each other's tiles, so every CTA must be resident
    SNNFLOW_REQUIRE(a.tile_flags != nullptr, "multi-bin launch without tile progress flags");
    SNNFLOW_CUDA(cudaMemsetAsync(a.tile_flags, 0, sizeof(unsigned int) * (size_t)n_tiles, st));
    static const int coop = env_int("SNNFLOW_COOP", 1);
    if (coop) SNNFLOW_CUDA(launch_coop(kernel, dim3(wt_grid(n_tiles)), dim3(WT_THREADS), smem, st, a));
    else SNNFLOW_CUDA(launch_pdl(kernel, dim3(wt_grid(n_tiles)), dim3(WT_THREADS), smem, st, a));
    return check_launch(what);
  }
  SNNFLOW_CUDA(launch_pdl(kernel, dim3(wt_grid(n_tiles)), dim3(WT_THREADS), smem, st, a));
  return check_launch(what);
}

int launch_wt_fwd(const WtArgs& a_in, bool seq, cudaStream_t st, const char* prof_name, double bytes, double flops) {
  WtArgs a = a_in;
  a.n_prod = (seq && a.n_col > 1 && a.n_src == 1) ? 2 : 1;
  const int nseg = a.R * a.n_seg;
  prof_begin(prof_name, st, bytes, flops);
  const bool hard = a.hard_reset != 0;
#define WT_FWD_CASE(SEQ, NS) \
  if (seq == SEQ && nseg == NS) { \
    if (hard) return wt_launch(wt_fwd_kernel<SEQ, NS, true>, (const void*)wt_fwd_kernel<SEQ, NS, true>, a, st, "wt_fwd_kernel"); \
    return wt_launch(wt_fwd_kernel<SEQ, NS, false>, (const void*)wt_fwd_kernel<SEQ, NS, false>, a, st, "wt_fwd_kernel"); \
  }
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

int launch_wt_dgpw(const WtArgs& a, cudaStream_t st, double bytes, double flops) {
  prof_begin("win_dgrad_pw", st, bytes, flops);
  const int nseg = a.R * a.n_seg;
#define WT_DP_CASE(SGV, HARDV, NS) \
  if (a.surrogate == SGV && (a.hard_reset != 0) == HARDV && nseg == NS) \
    return wt_launch(wt_dgpw_kernel<SGV, HARDV, NS>, (const void*)wt_dgpw_kernel<SGV, HARDV, NS>, a, st, "wt_dgpw_kernel");
  WT_DP_CASE(0, true, 1) WT_DP_CASE(0, false, 1) WT_DP_CASE(1, true, 1) WT_DP_CASE(1, false, 1) WT_DP_CASE(2, true, 1) WT_DP_CASE(2, false, 1)
  WT_DP_CASE(0, true, 2) WT_DP_CASE(0, false, 2) WT_DP_CASE(1, true, 2) WT_DP_CASE(1, false, 2) WT_DP_CASE(2, true, 2) WT_DP_CASE(2, false, 2)
#undef WT_DP_CASE
  set_error("launch_wt_dgpw: no kernel for surrogate %d / %d accumulator tiles", a.surrogate, nseg);
  return SNNFLOW_EINVAL;
}

int launch_wt_recbwd(const WtArgs& a, cudaStream_t st, double bytes, double flops) {
  prof_begin("win_rec_bwd", st, bytes, flops);
  // first bin of the window (t = 0): v_in / z_in are the caller's NCHW state (or zero); n_bins > 1: time-fused launch
  const int mode = a.n_bins > 1 ? 2 : (!a.z_from_v ? 1 : 0);
#define WT_RB_CASE(SGV, HARDV) \
  if (a.surrogate == SGV && (a.hard_reset != 0) == HARDV) { \
    if (mode == 2) return wt_launch(wt_recbwd_kernel<SGV, HARDV, 2>, (const void*)wt_recbwd_kernel<SGV, HARDV, 2>, a, st, "wt_recbwd_kernel"); \
    if (mode == 1) return wt_launch(wt_recbwd_kernel<SGV, HARDV, 1>, (const void*)wt_recbwd_kernel<SGV, HARDV, 1>, a, st, "wt_recbwd_kernel"); \
    return wt_launch(wt_recbwd_kernel<SGV, HARDV, 0>, (const void*)wt_recbwd_kernel<SGV, HARDV, 0>, a, st, "wt_recbwd_kernel"); \
  }
  WT_RB_CASE(0, true) WT_RB_CASE(0, false) WT_RB_CASE(1, true) WT_RB_CASE(1, false) WT_RB_CASE(2, true) WT_RB_CASE(2, false)
#undef WT_RB_CASE
  set_error("launch_wt_recbwd: unknown surrogate %d", a.surrogate);
  return SNNFLOW_EINVAL;
}

}  // namespace snnflow

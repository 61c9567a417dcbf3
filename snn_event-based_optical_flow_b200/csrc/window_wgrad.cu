// Layer-major window engine: weight gradient of one layer over ALL T*B images of a window in one launch
// (tcgen05, accumulators resident in TMEM for the whole kernel).
//
//   dW[co][ci][ky][kx] = sum over images, pixels  g_I[co][y][x] * X[ci][y+ky-1][x+kx-1]        (X = layer input; z_prev)
//
// Per image row this is a GEMM with K = pixels:  D[m][n = co] += A[m][k] * B[n][k],  both operands MN-major
// (8 channels = 16 B contiguous per pixel slot).  The trick that fills the 128-row MMA: shared memory holds the input
// rows interleaved as [row][chunk][P slots], so the M index (16 chunk-groups of 8 channels, a fixed byte stride
// apart) runs first over the chunks of row r and then over the chunks of rows r+1, r+2, ...: ONE tcgen05.mma
// computes several vertical taps (ky) at once - all three for a 32-channel feed-forward layer, two for a
// recurrent layer (x and z_prev chunks side by side).  The horizontal tap kx is a one-slot shift of the start
// address.  g_I comes as bf16 hi + lo planes (two MMAs per k-step); its zero border column and the zeroed pad slots
// make the k-steps that run past the end of a row contribute nothing.
// Each persistent CTA writes one partial block [tap][ci][co]; window_reduce sums them in a fixed order.
//
// Paired mode (feed-forward layers, two-row tiles): the kernel is bound by SHARED-MEMORY bandwidth - every K = 16 MMA
// re-reads its 4 KB A tile and its B tile, 288 KB per two-row item next to the 66 KB the bulk copies write, at 128 B/clk -
// so both gradient rows of the tile ride in ONE MMA (N = 4C: [row0 hi | row0 lo | row1 hi | row1 lo], contiguous chunk groups
// of the gradient tile) against the four input rows y0 .. y0+3: 8 KB per two rows instead of 12.  Accumulator block
// (input row i, gradient row c) holds vertical tap ky = i - c; the two blocks of a tap go to two partial blocks.
#include "tcgen05.cuh"
#include "window.cuh"

#include <stdlib.h>

#include <vector>

namespace snnflow {

constexpr int WG_EPI_WARPS = 8;
constexpr int WG_THREADS = (WG_EPI_WARPS + 3) * 32;   // + two TMA producer warps (alternating items, window_tc.cu) and the MMA issuer
constexpr int WG_MAX_STAGES = 4;
constexpr int WG_HDR = 1024;

static int wg_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

__device__ __forceinline__ int wg_n_items(const WgArgs& a) {
  const int n_tiles = a.n_img * (a.H / a.R);
  return ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
}

__device__ __forceinline__ uint32_t wg_tmem_cols(const WgArgs& a) {
  const uint32_t need = a.pair ? (uint32_t)(3 * 4 * a.C) : (uint32_t)(a.n_kyg * 3 * 2 * a.C);   // [x*g_hi | x*g_lo] per (tap group, kx)
  uint32_t c = 32;
  while (c < need) c <<= 1;
  return c;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wg_planes_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + WG_MAX_STAGES;
  uint64_t* done = empty + WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  unsigned char* stages = smem + WG_HDR;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < WG_MAX_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, wg_tmem_cols(a));
  // pad slots, spare rows and unused chunk groups must read as zeros: clear the whole ring once
  {
    uint4* p = reinterpret_cast<uint4*>(stages);
    const size_t n = (size_t)a.S * a.stage_bytes / 16;
    for (size_t i = tid; i < n; i += WG_THREADS) p[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = wg_n_items(a);
  const int tpi = a.H / a.R;
  const size_t plane_bytes = (size_t)(a.H + 2) * a.Wp * 16;
  const uint32_t row_bytes = (uint32_t)a.Wp * 16, pitch = (uint32_t)a.P * 16;
  const int g_chunks = a.C >> 3;

  if (warp == WG_EPI_WARPS || warp == WG_EPI_WARPS + 2) {
    const int pw = warp == WG_EPI_WARPS ? 0 : 1, n_pw = a.n_prod > 1 ? 2 : 1;
    // Producer warp.  A tile is (R+2)*n_cg input-row pieces + R*2*C/8 gradient-row pieces of one padded row each;
    // the lanes compute the piece addresses in parallel and each issues its own bulk copies (a single thread spends
    // ~100 issue cycles per copy next to four busy epilogue warps - measured, profiles/), lane 0 owns the barriers.
    pdl_wait();   // planes and gradient planes come from the preceding launches
    const int n_xp = (a.R + 2) * a.n_cg, n_gp = a.R * 2 * g_chunks;
    long long t_wait = 0;
    const long long t_begin = clock64();
    for (int k = pw; pw < n_pw && k < n_items; k += n_pw) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int img = tile / tpi, y0 = (tile - img * tpi) * a.R;
      const uint32_t st = (uint32_t)k % (uint32_t)a.S, use = (uint32_t)k / (uint32_t)a.S;
      unsigned char* xs = stages + (size_t)st * a.stage_bytes;
      unsigned char* gs = xs + a.g_off;
      if (lane == 0) {
        if (use > 0) {
          const long long t0 = clock64();
          mbar_wait(&empty[st], (use - 1) & 1);
          t_wait += clock64() - t0;
        }
        mbar_expect_tx(&full[st], (uint32_t)(n_xp + n_gp) * row_bytes);
      }
      __syncwarp();
      for (int i = lane; i < n_xp + n_gp; i += 32) {
        const unsigned char* src;
        unsigned char* dst;
        if (i < n_xp) {
          const int row = i / a.n_cg, cg = i - row * a.n_cg;
          const int si = cg < a.x_chunks[0] ? 0 : 1, ch = si ? cg - a.x_chunks[0] : cg;
          src = a.xp[si] + (size_t)img * a.x_img_stride[si] + (size_t)(y0 + row) * row_bytes + (size_t)ch * plane_bytes;
          dst = xs + (size_t)i * pitch;
        } else {
          const int j = i - n_xp;                      // (r, term, ch)
          const int ch = j % g_chunks, rt = j / g_chunks, term = rt & 1, r = rt >> 1;
          src = a.gp + (size_t)term * a.g_term_stride + (size_t)img * a.g_img_stride + (size_t)(y0 + 1 + r) * row_bytes +
                (size_t)ch * plane_bytes;
          dst = gs + (size_t)j * pitch;
        }
        tma_bulk_g2s(dst, src, row_bytes, &full[st]);
      }
    }
    if (a.dbg && lane == 0 && pw == 0) { a.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin; a.dbg[blockIdx.x * 8 + 1] = t_wait; }
    __syncwarp();
  } else if (warp == WG_EPI_WARPS + 1) {
    if (n_items > 0 && elect_one()) {
      // N = 2C: the hi and lo planes of g_I are adjacent chunk groups of one row, so ONE MMA per k-step yields both
      // partial products (columns [0,C) and [C,2C)); the read-out adds them.
      const uint32_t idesc = make_idesc(128, 2 * a.C, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
      const uint32_t idesc_pair = make_idesc(128, 4 * a.C, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
      const uint32_t stages16 = smem_u32(stages) >> 4, pitch16 = pitch >> 4;
      const uint32_t lo_c = ((128u >> 4) << 16), d_hi = desc_hi(pitch);   // LBO = 128 B (k groups), SBO = pitch (chunk groups)
      long long t_wait = 0;
      const long long t_begin = clock64();
      for (int k = 0; k < n_items; ++k) {
        const uint32_t st = (uint32_t)k % (uint32_t)a.S, use = (uint32_t)k / (uint32_t)a.S;
        {
          const long long t0 = clock64();
          mbar_wait(&full[st], use & 1);
          t_wait += clock64() - t0;
        }
        tc_fence_after();
        const uint32_t xs16 = stages16 + ((st * a.stage_bytes) >> 4), gs16 = xs16 + (a.g_off >> 4);
        if (a.pair) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t d = tmem_base + (uint32_t)(kx * 4 * a.C);
            uint32_t a_lo = lo_c | (xs16 + (uint32_t)kx), b_lo = lo_c | (gs16 + 1u);
            umma_f16_split(d, a_lo, d_hi, b_lo, d_hi, idesc_pair, k > 0 ? 1u : 0u);
#pragma unroll 4
            for (int kk = 1; kk < a.ksteps; ++kk) {
              a_lo += 16u;
              b_lo += 16u;
              umma_f16_split(d, a_lo, d_hi, b_lo, d_hi, idesc_pair, 1u);
            }
          }
        } else
        for (int r = 0; r < a.R; ++r) {
          const uint32_t grow16 = gs16 + (uint32_t)(r * 2 * g_chunks) * pitch16 + 1u;
          for (int j = 0; j < a.n_kyg; ++j) {
            const uint32_t arow16 = xs16 + (uint32_t)((r + j * a.rpm) * a.n_cg) * pitch16;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t d = tmem_base + (uint32_t)((j * 3 + kx) * 2 * a.C);
              // g slot c (padded column) pairs with input slot c + kx - 1; the k range starts at c = 1.
              // One add per operand and MMA: a k-step advances both descriptors by 16 slots.
              uint32_t a_lo = lo_c | (arow16 + (uint32_t)kx), b_lo = lo_c | grow16;
              umma_f16_split(d, a_lo, d_hi, b_lo, d_hi, idesc, (k > 0 || r > 0) ? 1u : 0u);
#pragma unroll 4
              for (int kk = 1; kk < a.ksteps; ++kk) {
                a_lo += 16u;
                b_lo += 16u;
                umma_f16_split(d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
              }
            }
          }
        }
        umma_commit(&empty[st]);
      }
      umma_commit(done);
      if (a.dbg) { a.dbg[blockIdx.x * 8 + 2] = clock64() - t_begin; a.dbg[blockIdx.x * 8 + 3] = t_wait; }
    }
    __syncwarp();
  } else if (n_items > 0) {
    // ---- read-out: D[(j, kx)] row M = (ky - j*rpm) * n_cg*8 + cg*8 + c  ->  part[tap][ci][co] ----
    pdl_wait();   // the partial buffers are read by the preceding layer's reduction
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3, par = warp >> 2;   // (epilogue warps 0..7)
    const int M = q * 32 + lane;
    const int i = M / (a.n_cg * 8), cg = (M >> 3) % a.n_cg, c = M & 7;
    const int si = (cg < a.x_chunks[0]) ? 0 : 1;
    const int ci = (si == 0 ? cg : cg - a.x_chunks[0]) * 8 + c;
    const bool row_used = i < a.rpm;
    if (a.pair) {
      // (kx, c): accumulator columns kx*4C + c*2C + [hi | lo]; this thread's input row i carries tap ky = i - c of gradient
      // row c; partial block 2*blockIdx.x + c (every tap of a block is written by exactly one thread)
      for (int id = par; id < 6; id += 2) {
        const int kx = id >> 1, cb = id & 1;
        const int ky = i - cb;
        const bool ok = ky >= 0 && ky < 3 && ci < a.cin_real[si];
        float* dst = a.part[si] + ((size_t)(2 * blockIdx.x + cb) * 9 + (ky * 3 + kx)) * a.cin_alloc[si] * a.C + (size_t)ci * a.C;
        for (int g = 0; g < (a.C >> 4); ++g) {
          float acc[16], acc1[16];
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kx * 4 * a.C + cb * 2 * a.C + g * 16);
          tmem_ld16(tcol, acc);
          tmem_ld16(tcol + (uint32_t)a.C, acc1);
#pragma unroll
          for (int v = 0; v < 16; ++v) acc[v] += acc1[v];
          if (ok) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              reinterpret_cast<float4*>(dst + g * 16)[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
          }
        }
      }
    } else
    for (int acc_id = par; acc_id < a.n_kyg * 3; acc_id += 2) {
      const int j = acc_id / 3, kx = acc_id - j * 3;
      const int ky = j * a.rpm + i;
      const bool ok = row_used && ky < 3 && ci < a.cin_real[si];
      float* dst = a.part[si] + ((size_t)blockIdx.x * 9 + (ky * 3 + kx)) * a.cin_alloc[si] * a.C + (size_t)ci * a.C;
      for (int g = 0; g < (a.C >> 4); ++g) {
        float acc[16], acc1[16];
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_id * 2 * a.C + g * 16);
        tmem_ld16(tcol, acc);
        tmem_ld16(tcol + (uint32_t)a.C, acc1);
#pragma unroll
        for (int v = 0; v < 16; ++v) acc[v] += acc1[v];
        if (ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            reinterpret_cast<float4*>(dst + g * 16)[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, wg_tmem_cols(a));
}

// ---- host ---------------------------------------------------------------------------------------------------
struct WgPlan {
  int n_cg, rpm, n_kyg, ksteps, P, x_rows, R, S, pair;
  uint32_t stage_bytes, g_off;
  bool ok;
};

static WgPlan wg_plan(int C, int cin_chunks, int rec_chunks, int H, int W) {
  WgPlan p{};
  p.ok = false;
  p.n_cg = cin_chunks + rec_chunks;
  if (p.n_cg <= 0 || p.n_cg > 16 || (16 % p.n_cg) != 0) return p;
  if ((C % 16) || C < 16 || C > 64) return p;
  p.rpm = 16 / p.n_cg;
  p.n_kyg = (3 + p.rpm - 1) / p.rpm;
  if (p.n_kyg * 3 * 2 * C > 512) return p;
  p.ksteps = ceil_div(W, 16);
  p.P = (int)align_up((size_t)(16 * p.ksteps + 3 > W + 2 ? 16 * p.ksteps + 3 : W + 2), 8);
  const int forced_R = wg_env_int("SNNFLOW_WG_R", 0), forced_S = wg_env_int("SNNFLOW_WG_S", 0);
  for (int R = 2; R >= 1; --R) {
    if (H % R) continue;
    if (forced_R && R != forced_R) continue;
    const int x_rows = R - 1 + p.n_kyg * p.rpm;   // rows addressed by the last tap group of the last output row
    // Only the R + 2 loaded rows get their own slots.  The MMA's spare vertical taps (ky >= 3) address rows past them:
    // those reads fall into the gradient tile of the same stage (finite bf16 data) and only feed accumulator rows
    // that the read-out ignores - this keeps a stage small enough for three of them.
    const size_t gb = (size_t)R * 2 * (C / 8) * p.P * 16;
    const int rows_in_g = (int)(gb / ((size_t)p.n_cg * p.P * 16));   // spare rows that the gradient tile can absorb
    const int own_rows = x_rows - rows_in_g > R + 2 ? x_rows - rows_in_g : R + 2;
    const size_t xb = (size_t)own_rows * p.n_cg * p.P * 16;
    const size_t stage = align_up(xb + gb, 128);
    int S = (int)(((size_t)227 * 1024 - WG_HDR) / stage);
    if (S > WG_MAX_STAGES) S = WG_MAX_STAGES;
    if (forced_S && S > forced_S) S = forced_S;
    if (S < 2) continue;
    p.R = R; p.S = S; p.x_rows = x_rows;
    p.pair = (R == 2 && p.n_kyg == 1 && p.rpm >= 4 && 3 * 4 * C <= 512 && wg_env_int("SNNFLOW_WG_PAIR", 1)) ? 1 : 0;
    p.g_off = (uint32_t)xb;
    p.stage_bytes = (uint32_t)stage;
    p.ok = true;
    break;
  }
  return p;
}

bool wg_supported(int C, int cin_chunks, int rec_chunks, int H, int W) { return wg_plan(C, cin_chunks, rec_chunks, H, W).ok; }

int wg_grid(int n_img, int H, int W, int C, int cin_chunks, int rec_chunks) {
  const WgPlan p = wg_plan(C, cin_chunks, rec_chunks, H, W);
  if (!p.ok) return 0;
  const int n_tiles = n_img * (H / p.R);
  return n_tiles < sm_count() ? n_tiles : sm_count();
}

int wg_parts(int n_img, int H, int W, int C, int cin_chunks, int rec_chunks) {
  const WgPlan p = wg_plan(C, cin_chunks, rec_chunks, H, W);
  return wg_grid(n_img, H, W, C, cin_chunks, rec_chunks) * (p.ok && p.pair ? 2 : 1);
}

int launch_wgrad_planes(WgArgs a, cudaStream_t st, double bytes, double flops) {
  const WgPlan p = wg_plan(a.C, a.x_chunks[0], a.n_xsrc > 1 ? a.x_chunks[1] : 0, a.H, a.W);
  if (!p.ok) {
    set_error("launch_wgrad_planes: shape not covered");
    return SNNFLOW_EINVAL;
  }
  a.n_cg = p.n_cg; a.rpm = p.rpm; a.n_kyg = p.n_kyg; a.ksteps = p.ksteps; a.P = p.P; a.x_rows = p.x_rows;
  a.R = p.R; a.S = p.S; a.stage_bytes = p.stage_bytes; a.g_off = p.g_off; a.Wp = a.W + 2; a.pair = p.pair;
  a.n_prod = wg_env_int("SNNFLOW_WG_PRODUCERS", wg_env_int("SNNFLOW_PRODUCERS", 2));
  const size_t smem = WG_HDR + (size_t)a.S * a.stage_bytes;
  static size_t attr = 0;
  if (smem > attr) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(wg_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_tiles = a.n_img * (a.H / a.R);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  prof_begin("win_wgrad", st, bytes, flops);
  if (wg_env_int("SNNFLOW_WT_TIMING", 0)) {   // debug: where do the producer and the MMA issuer wait? (synchronises)
    static long long* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, sizeof(long long) * 8 * 1024);
    cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * grid, st);
    a.dbg = dbg;
    launch_pdl(wg_planes_kernel, dim3(grid), dim3(WG_THREADS), smem, st, a);
    cudaStreamSynchronize(st);
    std::vector<long long> h(8 * grid);
    cudaMemcpy(h.data(), dbg, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost);
    double avg[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) avg[j] += (double)h[i * 8 + j] / grid;
    fprintf(stderr, "[wt-timing] wg_planes_kernel R=%d S=%d n_cg=%d pair=%d items/CTA=%.1f | producer %.0f (wait-empty %.0f) | mma %.0f (wait-full %.0f)\n",
            a.R, a.S, a.n_cg, a.pair, (double)n_tiles / grid, avg[0], avg[1], avg[2], avg[3]);
    return check_launch("wg_planes_kernel");
  }
  SNNFLOW_CUDA(launch_pdl(wg_planes_kernel, dim3(grid), dim3(WG_THREADS), smem, st, a));
  return check_launch("wg_planes_kernel");
}

}  // namespace snnflow

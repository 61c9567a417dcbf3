// ConvLIF backward on the tensor cores (tcgen05 + TMEM): weight gradient and data gradient.
//
//   dW[co][ci][ky][kx] = sum over pixels  g_I[b,co,y,x] * X[b,ci,y+ky-1,x+kx-1]        (X = x for ff, z_prev for rec)
//
// is, per tap, a GEMM  D_tap[m = ci (x channels, then z_prev channels)] [n = co] += A[m][k = pixel] * B[n][k]
// with K running over the pixels of the image.  Both operands live in shared memory in the same "slot" layout
// as the forward kernel (slot = pixel, 8 channels = 16 B contiguous, one plane per 8-channel chunk), which is the
// canonical MN-major UMMA layout: the tap (ky,kx) is again just a shifted start address.  The input side is
// spikes / counts (exact in bf16, one term); g_I is arbitrary fp32 and is split into bf16 hi + lo terms
// (16 mantissa bits, rel 2^-17: inside the 1e-4 gradient tolerance).  The nine accumulators D_tap (64 lanes x
// C columns each) stay resident in TMEM for the whole kernel: each persistent CTA accumulates all its pixel tiles
// and writes ONE partial [C][Cin][9] block at the end (fixed-order reduction by bwd_reduce_kernel: deterministic).
// Staging of tile i+1 (global loads, bf16 conversion) overlaps the asynchronous MMAs of tile i (2 smem stages).
#include "tcgen05.cuh"

namespace snnflow {

constexpr int WG_TC_STAGES = 2;
constexpr int WG_TC_XCHUNKS = 8;                                  // M = 64 rows = 8 chunks of 8 channels
constexpr int WG_TC_XBYTES = WG_TC_XCHUNKS * TC_SLOTS * 16;       // 50176
__host__ __device__ inline int wg_tc_gbytes(int C) { return 2 * (C >> 3) * TC_TW * 16; }   // hi + lo planes

struct TcWgradArgs {
  const float *g_cur, *x, *z;
  float *part_ff, *part_rec;
  int B, Cin, C, H, W, n_z;   // n_z = C/8 if the recurrent source is present else 0
};

// g_I rows -> bf16 hi/lo slot planes: plane(term, chunk) = [128 slots][8 channels]
__device__ __forceinline__ void stage_grad(const float* __restrict__ g, int C, unsigned char* s_g, int W, size_t plane,
                                           size_t row_off, int x0, bool vec_ok) {
  const int tid = threadIdx.x;
  const int n_chunks = C >> 3;
  auto split8 = [&](const float (&f)[8], uint4& hi, uint4& lo) {
    uint32_t uh[4], ul[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * c], f[2 * c + 1]);
      const float2 back = __bfloat1622float2(h);
      __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * c] - back.x, f[2 * c + 1] - back.y);
      uh[c] = *reinterpret_cast<uint32_t*>(&h);
      ul[c] = *reinterpret_cast<uint32_t*>(&l);
    }
    hi = make_uint4(uh[0], uh[1], uh[2], uh[3]);
    lo = make_uint4(ul[0], ul[1], ul[2], ul[3]);
  };
  uint4* hi_base = reinterpret_cast<uint4*>(s_g);
  uint4* lo_base = reinterpret_cast<uint4*>(s_g + (size_t)n_chunks * TC_TW * 16);
  if (vec_ok) {
    for (int task = tid; task < n_chunks * 32; task += TC_THREADS) {
      const int q = task & 31, j = task >> 5;
      const int xx = x0 + 4 * q;
      const bool ok = xx < W;
      const float* p = g + (size_t)j * 8 * plane + row_off + (ok ? xx : 0);
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = ok ? __ldg(reinterpret_cast<const float4*>(p + (size_t)c * plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float f0[8] = {v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x};
      const float f1[8] = {v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y};
      const float f2[8] = {v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z};
      const float f3[8] = {v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, v[6].w, v[7].w};
      uint4* dh = hi_base + j * TC_TW + 4 * q;
      uint4* dl = lo_base + j * TC_TW + 4 * q;
      split8(f0, dh[0], dl[0]); split8(f1, dh[1], dl[1]); split8(f2, dh[2], dl[2]); split8(f3, dh[3], dl[3]);
    }
  } else {
    for (int task = tid; task < n_chunks * TC_TW; task += TC_THREADS) {
      const int s = task % TC_TW, j = task / TC_TW;
      const int xx = x0 + s;
      const bool ok = xx < W;
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = ok ? __ldg(g + ((size_t)j * 8 + c) * plane + row_off + xx) : 0.f;
      split8(f, hi_base[j * TC_TW + s], lo_base[j * TC_TW + s]);
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_tc_kernel(TcWgradArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);   // bar[s]: MMAs reading stage s complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int gbytes = wg_tc_gbytes(a.C);
  const int stage_bytes = WG_TC_XBYTES + gbytes;
  unsigned char* stages = smem + 1024;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_x = a.Cin >> 3;

  if (tid == 0) {
    for (int s = 0; s < WG_TC_STAGES; ++s) mbar_init(&bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // unused M rows (chunks beyond x and z_prev) must read zeros, not stale shared memory
  for (int s = 0; s < WG_TC_STAGES; ++s) {
    uint4* xb = reinterpret_cast<uint4*>(stages + (size_t)s * stage_bytes);
    for (int i = (n_x + a.n_z) * TC_SLOTS + tid; i < WG_TC_XCHUNKS * TC_SLOTS; i += TC_THREADS) xb[i] = make_uint4(0, 0, 0, 0);
    // slots 390, 391 of the used chunks are never written by the staging code but are never read either
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_x = (a.W + TC_TW - 1) / TC_TW;
  const int n_tiles = a.B * a.H * tiles_x;
  const size_t plane = (size_t)a.H * a.W;
  const uint32_t idesc = make_idesc(64, a.C, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
  const bool vec_ok = ((a.W & 3) == 0) && ((((uintptr_t)a.x | (uintptr_t)a.g_cur) & 15) == 0) &&
                      (a.z == nullptr || (((uintptr_t)a.z) & 15) == 0);
  unsigned int inexact = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int b = tile / (a.H * tiles_x);
    const int rem = tile - b * (a.H * tiles_x);
    const int y0 = rem / tiles_x, x0 = (rem - y0 * tiles_x) * TC_TW;
    unsigned char* xb = stages + (size_t)s * stage_bytes;
    unsigned char* gb = xb + WG_TC_XBYTES;
    if (it >= WG_TC_STAGES) mbar_wait(&bar[s], (uint32_t)(((it >> 1) - 1) & 1));   // stage s free again
    stage_source<1>(a.x + (size_t)b * a.Cin * plane, n_x, xb, a.H, a.W, y0, x0, vec_ok, inexact);
    if (a.n_z) stage_source<1>(a.z + (size_t)b * a.C * plane, a.n_z, xb + (size_t)n_x * TC_SLOTS * 16, a.H, a.W, y0, x0, vec_ok, inexact);
    stage_grad(a.g_cur + (size_t)b * a.C * plane, a.C, gb, a.W, plane, (size_t)y0 * a.W, x0, vec_ok);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t xaddr = smem_u32(xb), gaddr = smem_u32(gb);
      const uint32_t g_term = (uint32_t)(a.C >> 3) * TC_TW * 16;
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t shift = (uint32_t)((tap / 3) * TC_P + (tap % 3)) * 16;
        const uint32_t d = tmem_base + (uint32_t)(tap * a.C);
        for (int kk = 0; kk < TC_TW / 16; ++kk) {
          // A[m = channel][k = slot]: chunks TC_SLOTS*16 B apart, 8-slot groups 128 B apart
          const uint64_t adesc = make_desc_mn(xaddr + shift + (uint32_t)kk * 256, 128, TC_SLOTS * 16);
#pragma unroll
          for (int term = 0; term < 2; ++term) {
            const uint64_t bdesc = make_desc_mn(gaddr + term * g_term + (uint32_t)kk * 256, 128, TC_TW * 16);
            umma_f16(d, adesc, bdesc, idesc, (it > 0 || kk > 0 || term > 0) ? 1u : 0u);
          }
        }
      }
      umma_commit(&bar[s]);
    }
  }
  // ---- drain: the last commit covers every earlier MMA (they complete in order) ----
  {
    const int last = it - 1, s = last & 1;
    mbar_wait(&bar[s], (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }
  // D_tap rows: row r -> TMEM lane (r % 16) + 32 * (r / 16)  (M = 64 accumulator layout); r = channel index
  const int quarter = warp & 3, tap_par = warp >> 2;
  const int row = quarter * 16 + lane;            // valid for lane < 16
  const bool is_ff = row < 8 * n_x;
  const int ci = is_ff ? row : row - 8 * n_x;
  const int n_ci = is_ff ? a.Cin : a.C;
  const bool row_ok = (lane < 16) && (is_ff || (row - 8 * n_x) < 8 * a.n_z);
  float* part = is_ff ? a.part_ff + (size_t)blockIdx.x * a.C * a.Cin * 9 : a.part_rec + (size_t)blockIdx.x * a.C * a.C * 9;
  for (int tap = tap_par; tap < 9; tap += 2) {
    for (int g = 0; g < (a.C >> 4); ++g) {
      float acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tap * a.C + g * 16), acc);
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < 16; ++c) part[((size_t)(g * 16 + c) * n_ci + ci) * 9 + tap] = acc[c];
      }
    }
  }
  (void)inexact;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

bool wgrad_tc_supported(int Cin, int C, int recurrent) {
  if ((Cin % 8) || (C % 16) || C > 48 || Cin < 8) return false;
  return (Cin >> 3) + (recurrent ? (C >> 3) : 0) <= WG_TC_XCHUNKS;
}

// grid size used by the launcher (partials are sized by the caller from this)
int wgrad_tc_grid(int B, int H, int W) {
  const int n_tiles = B * H * ceil_div(W, TC_TW);
  const int sms = sm_count();
  return n_tiles < sms ? n_tiles : sms;
}

int launch_wgrad_tc(const float* g_cur, const float* x, const float* z, float* part_ff, float* part_rec, int B, int Cin,
                    int C, int H, int W, cudaStream_t st) {
  TcWgradArgs a{};
  a.g_cur = g_cur; a.x = x; a.z = z; a.part_ff = part_ff; a.part_rec = part_rec;
  a.B = B; a.Cin = Cin; a.C = C; a.H = H; a.W = W; a.n_z = z ? (C >> 3) : 0;
  const size_t smem = 1024 + (size_t)WG_TC_STAGES * (WG_TC_XBYTES + wg_tc_gbytes(C));
  static size_t attr = 0;
  if (smem > attr) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int grid = wgrad_tc_grid(B, H, W);
  const double px = (double)B * H * W;
  prof_begin("wgrad_tc", st, 4.0 * px * (C + Cin + (z ? C : 0)) + 4.0 * grid * C * (Cin + (z ? C : 0)) * 9,
             18.0 * px * C * (Cin + (z ? C : 0)));
  wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
  return check_launch("wgrad_tc_kernel");
}


// ================================================================================================
// Data gradient on the tensor cores.
//   g_x[b,ci,y,x]      = sum_{co,ky,kx} g_I[b,co,y+1-ky,x+1-kx] * W_ff [co][ci][ky][kx]
//   g_zprev[b,cz,y,x]  = the same with W_rec                       (recurrent cells)
// = one implicit GEMM per row tile: D[m = pixel][n = ci | Cin + cz] += A_tap[m][k = co] * B_tap[n][k], A = g_I in the
// K-major slot layout (3 rows + halo, bf16 hi + lo planes), B = the transposed, tap-flipped weights (bf16 hi + lo,
// packed by dgrad_pack_kernel into the UMMA image, one TMA bulk copy per CTA).  Three MMAs per (tap, k-step):
// hi*hi + hi*lo + lo*hi (the lo*lo term is below 2^-32 relative).  One persistent CTA per SM with two smem stages
// and two TMEM accumulators: the MMAs of tile i run while tile i-1 is drained to global and tile i+1 is staged.
// ================================================================================================
struct TcDgradArgs {
  const float* g_cur;
  const unsigned char* blob;
  float *g_x, *g_z;
  int B, C, Cin, n_rec, H, W, accumulate_z;
  uint32_t blob_bytes;
};

__host__ __device__ inline size_t dg_tc_blob_bytes(int C, int N) { return (size_t)9 * 2 * N * C * 2; }

// blob: for tap' 0..8: for term (hi, lo): bf16 [C/8][N/8][8 n][8 k];  value(n, k) = W[k][n][8 - tap']
__global__ void __launch_bounds__(256) dgrad_pack_kernel(const float* __restrict__ w_ff, const float* __restrict__ w_rec,
                                                         __nv_bfloat16* __restrict__ blob, int C, int Cin, int n_rec) {
  const int N = Cin + n_rec, per_tile = N * C;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < 9 * per_tile; i += gridDim.x * 256) {
    const int tap = i / per_tile, r = i - tap * per_tile;
    const int n = r / C, k = r - n * C;
    const float v = n < Cin ? w_ff[((size_t)k * Cin + n) * 9 + (8 - tap)] : w_rec[((size_t)k * n_rec + (n - Cin)) * 9 + (8 - tap)];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int idx = ((k >> 3) * (N >> 3) + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7);
    blob[(size_t)(tap * 2 + 0) * per_tile + idx] = hi;
    blob[(size_t)(tap * 2 + 1) * per_tile + idx] = lo;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) dgrad_tc_kernel(TcDgradArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8);   // bar[0..1]: MMAs of the tile in stage s complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  unsigned char* s_w = smem + 1024;
  const int n_ch = a.C >> 3;
  const uint32_t plane_bytes = (uint32_t)n_ch * TC_SLOTS * 16;        // one term of one stage
  unsigned char* s_a = s_w + ((a.blob_bytes + 1023) / 1024) * 1024;    // [stage][term][chunk][slot][8]
  const int N = a.Cin + a.n_rec;
  const uint32_t ncols = 2 * N <= 32 ? 32u : (2 * N <= 64 ? 64u : (2 * N <= 128 ? 128u : 256u));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar_w, a.blob_bytes);
    tma_bulk_g2s(s_w, a.blob, a.blob_bytes, bar_w);
  }

  const int tiles_x = (a.W + TC_TW - 1) / TC_TW;
  const int n_tiles = a.B * a.H * tiles_x;
  const size_t plane = (size_t)a.H * a.W;
  const uint32_t idesc = make_idesc(TC_TW, N, /*bf16*/ 1, 0, 0);
  const bool vec_ok = ((a.W & 3) == 0) && ((((uintptr_t)a.g_cur) & 15) == 0);
  const uint32_t a_lbo = TC_SLOTS * 16, b_lbo = (uint32_t)(N >> 3) * 128, w_tile = (uint32_t)N * a.C * 2;
  unsigned int unused = 0;
  bool weights_ready = false;

  for (int it = 0;; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    const bool have = tile < n_tiles;
    const int s = it & 1;
    if (have) {
      const int b = tile / (a.H * tiles_x);
      const int rem = tile - b * (a.H * tiles_x);
      const int y0 = rem / tiles_x, x0 = (rem - y0 * tiles_x) * TC_TW;
      unsigned char* hi = s_a + (size_t)s * 2 * plane_bytes;
      stage_source<2>(a.g_cur + (size_t)b * a.C * plane, n_ch, hi, a.H, a.W, y0, x0, vec_ok, unused, hi + plane_bytes);
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();   // stage s written; the epilogue of tile it-2 has drained accumulator s
    if (have && tid == 0) {
      if (!weights_ready) { mbar_wait(bar_w, 0); weights_ready = true; }
      tc_fence_after();
      const uint32_t hi_addr = smem_u32(s_a) + (uint32_t)s * 2 * plane_bytes, lo_addr = hi_addr + plane_bytes;
      const uint32_t d = tmem_base + (uint32_t)(s * N);
      uint32_t accumulate = 0;
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t shift = (uint32_t)((tap / 3) * TC_P + (tap % 3)) * 16;
        const uint32_t wb = smem_u32(s_w) + (uint32_t)(tap * 2) * w_tile;
        for (int kk = 0; kk < (a.C >> 4); ++kk) {
          const uint64_t a_hi = make_desc(hi_addr + (uint32_t)(2 * kk) * a_lbo + shift, a_lbo, 128);
          const uint64_t a_lo = make_desc(lo_addr + (uint32_t)(2 * kk) * a_lbo + shift, a_lbo, 128);
          const uint64_t b_hi = make_desc(wb + (uint32_t)(2 * kk) * b_lbo, b_lbo, 128);
          const uint64_t b_lo = make_desc(wb + w_tile + (uint32_t)(2 * kk) * b_lbo, b_lbo, 128);
          umma_f16(d, a_hi, b_hi, idesc, accumulate);
          umma_f16(d, a_hi, b_lo, idesc, 1u);
          umma_f16(d, a_lo, b_hi, idesc, 1u);
          accumulate = 1;
        }
      }
      umma_commit(&bar[s]);
    }
    if (it >= 1) {   // drain tile it-1 while the MMAs of tile it run
      const int pt = blockIdx.x + (it - 1) * gridDim.x, ps = (it - 1) & 1;
      const int b = pt / (a.H * tiles_x);
      const int rem = pt - b * (a.H * tiles_x);
      const int y0 = rem / tiles_x, x0 = (rem - y0 * tiles_x) * TC_TW;
      const int xo = x0 + quarter * 32 + lane;
      const bool px_ok = xo < a.W;
      const size_t pix = (size_t)y0 * a.W + xo;
      mbar_wait(&bar[ps], (uint32_t)(((it - 1) >> 1) & 1));
      tc_fence_after();
      for (int g = half; g < (N >> 4); g += 2) {
        float acc[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ps * N + g * 16), acc);
        if (px_ok) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int n = g * 16 + c;
            if (n < a.Cin) {
              if (a.g_x) a.g_x[((size_t)b * a.Cin + n) * plane + pix] = acc[c];
            } else {
              float* p = a.g_z + ((size_t)b * a.n_rec + (n - a.Cin)) * plane + pix;
              *p = a.accumulate_z ? *p + acc[c] : acc[c];
            }
          }
        }
      }
    }
    if (!have) break;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static size_t dg_tc_smem(int C, int N) {
  return 1024 + align_up(dg_tc_blob_bytes(C, N), 1024) + (size_t)2 * 2 * (C >> 3) * TC_SLOTS * 16;
}

bool dgrad_tc_supported(int Cin, int C, int n_rec, bool need_gx) {
  if ((C % 16) || C > 64 || C < 16) return false;
  const int N = (need_gx || n_rec == 0 ? Cin : Cin) + n_rec;   // the g_x columns are always computed
  if ((Cin % 8) || (N % 16) || N > 128 || N < 16) return false;
  return dg_tc_smem(C, N) <= 227 * 1024;
}

size_t dgrad_tc_workspace_bytes(int Cin, int C, int n_rec) { return align_up(dg_tc_blob_bytes(C, Cin + n_rec), 256); }

int launch_dgrad_tc(const float* g_cur, const float* w_ff, const float* w_rec, float* g_x, float* g_z, int accumulate_z,
                    void* blob_ws, int B, int Cin, int C, int n_rec, int H, int W, cudaStream_t st) {
  const int N = Cin + n_rec;
  prof_begin("dgrad_pack", st, 4.0 * 9 * C * N + 2.0 * 2 * 9 * C * N);
  dgrad_pack_kernel<<<ceil_div(9 * N * C, 256 * 4), 256, 0, st>>>(w_ff, w_rec, (__nv_bfloat16*)blob_ws, C, Cin, n_rec);
  int rc = check_launch("dgrad_pack_kernel");
  if (rc) return rc;
  TcDgradArgs a{};
  a.g_cur = g_cur; a.blob = (const unsigned char*)blob_ws; a.g_x = g_x; a.g_z = g_z;
  a.B = B; a.C = C; a.Cin = Cin; a.n_rec = n_rec; a.H = H; a.W = W; a.accumulate_z = accumulate_z;
  a.blob_bytes = (uint32_t)dg_tc_blob_bytes(C, N);
  const size_t smem = dg_tc_smem(C, N);
  static size_t attr = 0;
  if (smem > attr) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_tiles = B * H * ceil_div(W, TC_TW);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  const double px = (double)B * H * W;
  prof_begin("dgrad_tc", st, 4.0 * px * (C + (g_x ? Cin : 0) + n_rec * (accumulate_z ? 2 : 1)), 18.0 * px * C * N);
  dgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
  return check_launch("dgrad_tc_kernel");
}

}  // namespace snnflow

// ConvLIF / ConvLIFRecurrent forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Implicit GEMM per output-row tile:  D[m = pixel (128)] [n = out channel (C)] += A_tap[m][k = in channel] * W_tap[n][k]
// summed over the 9 taps (and over the two sources x / z_prev of a recurrent cell).
//
//  * A operand: the fp32 NCHW input rows y0-1..y0+1 (with a 1-pixel halo, zero filled at the borders) are
//    converted to fp16 and stored ONCE in shared memory as "slots": slot s = row*130 + col holds the 8 channels
//    of one pixel contiguously (16 B), one plane of slots per 8-channel chunk.  That is the canonical
//    no-swizzle K-major UMMA layout (8 rows x 16 B core matrices, 128 B per 8-slot group), and because the
//    tile is linearised, the A matrix of tap (ky,kx) is the SAME buffer started (ky*130 + kx) slots later:
//    nine shifted descriptors instead of an im2col copy.  Spikes {0,1} and event counts are exact in fp16.
//  * B operand: fp32 weights, scaled by a power of two and split into two fp16 terms (hi + lo = 22 mantissa
//    bits: the conv equals the fp32 conv up to ~1 ulp, and is bit-exact for 2^-12-grid weights), pre-packed by
//    snnflow_convlif_pack into the UMMA smem image and brought in by one TMA bulk copy per CTA.
//  * D accumulates in TMEM (fp32, 128 lanes x C columns); one elected thread issues all tcgen05.mma and
//    commits to an mbarrier; the 4 warps then pull their 32 lanes with tcgen05.ld and run the LIF update
//    (leak, delayed reset, threshold, spike) straight out of the accumulator - v/z state is read and written
//    exactly once, coalesced along x.
//  * Persistent CTAs (several per SM, so one CTA's loads overlap another's epilogue) stride over the tiles.
#include "tcgen05.cuh"

namespace snnflow {

__device__ unsigned int g_tc_inexact = 0;   // inputs that were not exactly representable in fp16

// ---- weight packing ---------------------------------------------------------------------------
// blob = for conv in (ff[, rec]): for tap 0..8: for term (hi, lo): fp16 [K/8][C/8][8 n][8 k]   then 1 float: 1/scale
__host__ __device__ inline size_t tc_conv_bytes(int K, int C) { return (size_t)9 * TC_TERMS * K * C * sizeof(__half); }

__global__ void __launch_bounds__(256) convlif_pack_kernel(const float* __restrict__ w_ff, const float* __restrict__ w_rec,
                                                           unsigned char* __restrict__ blob, int Cin, int C) {
  __shared__ float red[8];
  __shared__ float s_scale;
  const int n_ff = C * Cin * 9, n_rec = w_rec ? C * C * 9 : 0;
  float mx = 0.f;
  for (int i = threadIdx.x; i < n_ff; i += 256) mx = fmaxf(mx, fabsf(w_ff[i]));
  for (int i = threadIdx.x; i < n_rec; i += 256) mx = fmaxf(mx, fabsf(w_rec[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i]);
    int e = 0;
    float scale = 1.0f;
    if (m > 0.f && isfinite(m)) {
      frexpf(m, &e);                       // m = f * 2^e, f in [0.5, 1)
      int s = 14 - e;                      // max * 2^s in [2^13, 2^14): well inside fp16 range
      s = s > 24 ? 24 : (s < -24 ? -24 : s);
      scale = ldexpf(1.0f, s);
    }
    s_scale = scale;
  }
  __syncthreads();
  const float scale = s_scale;
  size_t off = 0;
  for (int conv = 0; conv < (w_rec ? 2 : 1); ++conv) {
    const float* w = conv == 0 ? w_ff : w_rec;
    const int K = conv == 0 ? Cin : C;
    __half* dst = reinterpret_cast<__half*>(blob + off);
    const int per_tile = K * C;
    for (int i = threadIdx.x; i < 9 * per_tile; i += 256) {
      const int tap = i / per_tile, r = i - tap * per_tile;
      const int n = r / K, k = r - n * K;
      const float v = w[((size_t)n * K + k) * 9 + tap] * scale;   // exact: power-of-two scale
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const int idx = ((k >> 3) * (C >> 3) + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7);
      dst[(size_t)(tap * TC_TERMS + 0) * per_tile + idx] = hi;
      dst[(size_t)(tap * TC_TERMS + 1) * per_tile + idx] = lo;
    }
    off += tc_conv_bytes(K, C);
  }
  if (threadIdx.x == 0) *reinterpret_cast<float*>(blob + off) = 1.0f / scale;
}

// ---- forward kernel ---------------------------------------------------------------------------
struct TcFwdArgs {
  const float *x, *z_src;               // z_src: recurrent input (== z_in) or nullptr
  const unsigned char* blob;
  const float *v_in, *z_in, *lam, *theta, *residual;
  float *v_out, *z_out, *out, *cur_out;
  int B, Cin, C, H, W, n_conv, hard_reset;
  uint32_t blob_bytes;                  // weights only (multiple of 16), the scale float follows
};

__device__ __forceinline__ float lif_update_tc(float v, float z, float cur, float lam, float theta, int hard) {
  float a = __fmul_rn(v, lam);
  float c = __fmul_rn(__fsub_rn(1.0f, lam), cur);
  if (hard) return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, z)), c);
  return __fsub_rn(__fadd_rn(a, c), __fmul_rn(z, theta));
}

// One thread issues the MMAs of one conv (9 taps x K/16 k-steps x 2 weight terms) into the accumulator.
__device__ __forceinline__ void issue_conv(uint32_t tmem_d, uint32_t a_base, uint32_t w_base, int K, int C, uint32_t idesc,
                                           uint32_t& accumulate) {
  const uint32_t a_lbo = TC_SLOTS * 16, b_lbo = (uint32_t)(C >> 3) * 128, tile_bytes = (uint32_t)K * C * 2;
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t shift = (uint32_t)((tap / 3) * TC_P + (tap % 3)) * 16;
    for (int kk = 0; kk < (K >> 4); ++kk) {
      const uint64_t adesc = make_desc(a_base + (uint32_t)(2 * kk) * a_lbo + shift, a_lbo, 128);
#pragma unroll
      for (int term = 0; term < TC_TERMS; ++term) {
        const uint64_t bdesc = make_desc(w_base + (uint32_t)(tap * TC_TERMS + term) * tile_bytes + (uint32_t)(2 * kk) * b_lbo, b_lbo, 128);
        umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
        accumulate = 1;
      }
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 2) convlif_fwd_tc_kernel(TcFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);          // weights landed
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + 8);    // MMAs of one conv complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  unsigned char* s_w = smem + 1024;
  unsigned char* s_a = s_w + ((a.blob_bytes + 1023) / 1024) * 1024;   // ONE slot buffer, reused by x then z_prev

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;   // TMEM lane quarter / which 16-channel groups this warp owns
  const uint32_t ncols = a.C <= 32 ? 32u : 64u;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
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
  const float scale_inv = *reinterpret_cast<const float*>(a.blob + a.blob_bytes);

  const int tiles_x = (a.W + TC_TW - 1) / TC_TW;
  const int n_tiles = a.B * a.H * tiles_x;
  const size_t plane = (size_t)a.H * a.W;
  // instruction descriptor (kind::f16): D = f32 [4,6)=1, A = B = f16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (1u << 4) | ((uint32_t)(a.C >> 3) << 17) | ((uint32_t)(TC_TW >> 4) << 24);
  const bool vec_ok = ((a.W & 3) == 0) && ((((uintptr_t)a.x) & 15) == 0) && (a.z_src == nullptr || (((uintptr_t)a.z_src) & 15) == 0);
  const uint32_t ff_bytes = (uint32_t)tc_conv_bytes(a.Cin, a.C);
  uint32_t mma_parity = 0;
  unsigned int inexact = 0;
  bool weights_ready = false;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / (a.H * tiles_x);
    const int rem = tile - b * (a.H * tiles_x);
    const int y0 = rem / tiles_x, x0 = (rem - y0 * tiles_x) * TC_TW;

    uint32_t accumulate = 0;
    for (int conv = 0; conv < a.n_conv; ++conv) {
      if (conv == 1) {   // the slot buffer is still being read by the ff MMAs
        mbar_wait(bar_mma, mma_parity);
        mma_parity ^= 1;
      }
      const float* src = conv == 0 ? a.x + (size_t)b * a.Cin * plane : a.z_src + (size_t)b * a.C * plane;
      stage_source<0>(src, (conv == 0 ? a.Cin : a.C) >> 3, s_a, a.H, a.W, y0, x0, vec_ok, inexact);
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncthreads();
      if (tid == 0) {
        if (!weights_ready) { mbar_wait(bar_w, 0); weights_ready = true; }
        tc_fence_after();
        issue_conv(tmem_base, smem_u32(s_a), smem_u32(s_w) + (conv == 0 ? 0u : ff_bytes), conv == 0 ? a.Cin : a.C, a.C, idesc,
                   accumulate);
        umma_commit(bar_mma);   // implies tcgen05.fence::before_thread_sync
      }
    }

    // ---- epilogue: TMEM -> registers -> LIF -> global; the state loads are issued before waiting ----
    const int xo = x0 + quarter * 32 + lane;
    const bool px_ok = xo < a.W;
    const size_t pix = (size_t)y0 * a.W + xo;
    bool waited = false;
    for (int g = half; g < (a.C >> 4); g += 2) {
      float vin[16], zin[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const size_t idx = ((size_t)b * a.C + g * 16 + c) * plane + pix;
        vin[c] = (px_ok && a.v_in) ? __ldg(a.v_in + idx) : 0.f;
        zin[c] = (px_ok && a.z_in) ? __ldg(a.z_in + idx) : 0.f;
      }
      if (!waited) {
        mbar_wait(bar_mma, mma_parity);
        tc_fence_after();
        waited = true;
      }
      float acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 16), acc);
      if (px_ok) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int co = g * 16 + c;
          const size_t idx = ((size_t)b * a.C + co) * plane + pix;
          const float lam = __ldg(a.lam + co), theta = __ldg(a.theta + co);
          const float cur = acc[c] * scale_inv;
          const float vn = lif_update_tc(vin[c], zin[c], cur, lam, theta, a.hard_reset);
          const float zn = (__fsub_rn(vn, theta) > 0.f) ? 1.f : 0.f;
          a.v_out[idx] = vn;
          a.z_out[idx] = zn;
          if (a.out) a.out[idx] = a.residual ? __fadd_rn(zn, __ldg(a.residual + idx)) : zn;
          if (a.cur_out) a.cur_out[idx] = cur;
        }
      }
    }
    if (!waited) mbar_wait(bar_mma, mma_parity);   // warps without a channel group still track the phase
    mma_parity ^= 1;
    tc_fence_before();
    __syncthreads();   // TMEM drained and slot buffer free before the next tile overwrites them
    tc_fence_after();
  }
  if (inexact) atomicAdd(&g_tc_inexact, inexact);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static bool tc_shape_ok(int Cin, int C) {
  return Cin >= 16 && Cin <= TC_MAX_C && (Cin % 16) == 0 && C >= 16 && C <= TC_MAX_C && (C % 16) == 0;
}
static size_t tc_blob_weight_bytes(int Cin, int C, int recurrent) {
  return tc_conv_bytes(Cin, C) + (recurrent ? tc_conv_bytes(C, C) : 0);
}

}  // namespace snnflow
using namespace snnflow;

static size_t tc_smem_bytes(int Cin, int C, int recurrent) {
  const size_t kc = (size_t)((recurrent && C > Cin ? C : Cin) >> 3);   // one slot buffer, reused by x and z_prev
  return 1024 + align_up(tc_blob_weight_bytes(Cin, C, recurrent), 1024) + kc * TC_SLOTS * 16;
}

extern "C" size_t snnflow_convlif_packed_bytes(int Cin, int C, int recurrent) {
  if (!tc_shape_ok(Cin, C)) return 0;
  if (tc_smem_bytes(Cin, C, recurrent) > 227 * 1024) return 0;   // weights + tile must fit in one SM's shared memory
  return tc_blob_weight_bytes(Cin, C, recurrent) + 16;
}

extern "C" int snnflow_convlif_pack(const float* w_ff, const float* w_rec, void* packed, int Cin, int C,
                                    snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(w_ff && packed, "null pointer");
  SNNFLOW_REQUIRE(snnflow_convlif_packed_bytes(Cin, C, w_rec != nullptr) > 0,
                  "shape not covered by the tensor-core path (Cin, C multiples of 16, <= 64, fitting in shared memory)");
  SNNFLOW_REQUIRE(((uintptr_t)packed & 15) == 0, "packed buffer must be 16-byte aligned");
  prof_begin("convlif_pack", (cudaStream_t)stream, 4.0 * 9 * C * (Cin + (w_rec ? C : 0)) * 2);
  convlif_pack_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(w_ff, w_rec, (unsigned char*)packed, Cin, C);
  return check_launch("convlif_pack_kernel");
}

extern "C" unsigned int snnflow_tc_inexact_count(int reset) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_inexact, sizeof(v));
  if (reset) {
    unsigned int z = 0;
    cudaMemcpyToSymbol(g_tc_inexact, &z, sizeof(z));
  }
  return v;
}

extern "C" int snnflow_convlif_fwd_tc(const float* x, const void* packed, int recurrent, const float* v_in,
                                      const float* z_in, const float* lam, const float* theta, const float* residual,
                                      float* v_out, float* z_out, float* out, float* cur_out, int B, int Cin, int C,
                                      int H, int W, unsigned flags, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(x && packed && lam && theta && v_out && z_out, "null pointer");
  SNNFLOW_REQUIRE((v_in == nullptr) == (z_in == nullptr), "v_in and z_in must both be given or both be NULL");
  SNNFLOW_REQUIRE(snnflow_convlif_packed_bytes(Cin, C, recurrent) > 0, "shape not covered by the tensor-core path");
  SNNFLOW_REQUIRE(B > 0 && H > 0 && W > 0, "bad dims");
  SNNFLOW_REQUIRE(!(residual && !out), "residual given without out");
  SNNFLOW_REQUIRE(((uintptr_t)packed & 15) == 0, "packed weights must be 16-byte aligned");
  TcFwdArgs a{};
  a.x = x; a.blob = (const unsigned char*)packed;
  a.n_conv = (recurrent && z_in) ? 2 : 1;       // z_in == NULL: zero state, the recurrent current vanishes
  a.z_src = a.n_conv == 2 ? z_in : nullptr;
  a.v_in = v_in; a.z_in = z_in; a.lam = lam; a.theta = theta; a.residual = residual;
  a.v_out = v_out; a.z_out = z_out; a.out = out; a.cur_out = cur_out;
  a.B = B; a.Cin = Cin; a.C = C; a.H = H; a.W = W;
  a.hard_reset = (flags & SNNFLOW_HARD_RESET) ? 1 : 0;
  // the whole blob (ff + rec) is staged even on the first step; only the convs in use are issued
  a.blob_bytes = (uint32_t)tc_blob_weight_bytes(Cin, C, recurrent);
  const size_t smem = tc_smem_bytes(Cin, C, recurrent);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(convlif_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int n_tiles = B * H * ceil_div(W, TC_TW);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);   // register-limited to 2 CTAs of 256 threads
  int grid = sm_count() * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  {
    const double px = (double)B * H * W;
    const int planes = Cin + 2 * C + (v_in ? 2 * C : 0) + (out ? C : 0) + (residual ? C : 0) + (cur_out ? C : 0);
    prof_begin("convlif_fwd_tc", (cudaStream_t)stream, 4.0 * px * planes, 18.0 * px * C * (Cin + (a.n_conv == 2 ? C : 0)));
  }
  convlif_fwd_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(a);
  return check_launch("convlif_fwd_tc_kernel");
}

// ConvLIF / ConvLIFRecurrent layer-step: forward (fused conv + LIF) and BPTT backward.
// Reference semantics: models/spiking_submodules.py:121-151, :265-300; models/spiking_util.py.
#include "common.cuh"
#include "conv3x3_simt.cuh"

namespace snnflow {

// tensor-core weight gradient (convlif_bwd_tc.cu)
bool wgrad_tc_supported(int Cin, int C, int recurrent);
int wgrad_tc_grid(int B, int H, int W);
int launch_wgrad_tc(const float* g_cur, const float* x, const float* z, float* part_ff, float* part_rec, int B, int Cin,
                    int C, int H, int W, cudaStream_t st);
// tensor-core data gradient (convlif_bwd_tc.cu)
bool dgrad_tc_supported(int Cin, int C, int n_rec, bool need_gx);
size_t dgrad_tc_workspace_bytes(int Cin, int C, int n_rec);
int launch_dgrad_tc(const float* g_cur, const float* w_ff, const float* w_rec, float* g_x, float* g_z, int accumulate_z,
                    void* blob_ws, int B, int Cin, int C, int n_rec, int H, int W, cudaStream_t st);

// ============================================================================================
// Forward: conv (CUDA cores, exact fp32) + leak + delayed reset + threshold + spike.
// ============================================================================================
struct LifFwdArgs {
  ConvSrc src[2];
  int n_src;
  const float *v_in, *z_in, *lam, *theta, *residual;
  float *v_out, *z_out, *out, *cur_out;
  int B, C, H, W;
  int hard_reset;
};

// The LIF update with the reference's rounding order (each Python operator is one rounded fp32 op):
//   hard: ((v*lam)*(1-z)) + ((1-lam)*I)          spiking_submodules.py:144 / :293
//   soft: ((v*lam) + ((1-lam)*I)) - (z*theta)     spiking_submodules.py:146 / :295
__device__ __forceinline__ float lif_update(float v, float z, float cur, float lam, float theta, int hard) {
  float a = __fmul_rn(v, lam);
  float c = __fmul_rn(__fsub_rn(1.0f, lam), cur);
  if (hard) return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, z)), c);
  return __fsub_rn(__fadd_rn(a, c), __fmul_rn(z, theta));
}

__global__ void __launch_bounds__(CT_THREADS) convlif_fwd_simt_kernel(LifFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int n_cog = (a.C + CT_CO - 1) / CT_CO;
  const int b = blockIdx.z / n_cog, co0 = (blockIdx.z % n_cog) * CT_CO;
  const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;
  float acc[4][CT_CO];
  conv3x3_tile(a.src, a.n_src, b, co0, a.C, a.H, a.W, y0, x0, smem, acc);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = x0 + lane;
  if (x >= a.W) return;
  const size_t plane = (size_t)a.H * a.W;
#pragma unroll
  for (int c = 0; c < CT_CO; ++c) {
    const int co = co0 + c;
    if (co >= a.C) break;
    const float lam = a.lam[co], theta = a.theta[co];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = y0 + warp * 4 + r;
      if (y >= a.H) continue;
      const size_t idx = ((size_t)b * a.C + co) * plane + (size_t)y * a.W + x;
      const float v = a.v_in ? a.v_in[idx] : 0.f;
      const float z = a.z_in ? a.z_in[idx] : 0.f;
      const float cur = acc[r][c];
      const float vn = lif_update(v, z, cur, lam, theta, a.hard_reset);
      const float zn = (__fsub_rn(vn, theta) > 0.f) ? 1.f : 0.f;   // spiking_util.py:21 strict '>'
      a.v_out[idx] = vn;
      a.z_out[idx] = zn;
      if (a.out) a.out[idx] = a.residual ? __fadd_rn(zn, a.residual[idx]) : zn;
      if (a.cur_out) a.cur_out[idx] = cur;
    }
  }
}

// ============================================================================================
// Backward phase A: elementwise surrogate / leak / reset chain + per-channel reductions.
//   gs = g_z * sg(v' - theta);  gv = g_v' + gs;  g_I = gv * (1 - lam)
//   hard: g_v_in = gv*lam*(1-z_in); dlam += gv*(v_in*(1-z_in) - I); dtheta -= gs
//   soft: g_v_in = gv*lam;          dlam += gv*(v_in - I);          dtheta -= gs + gv*z_in
//   reset path (only when not detached): g_z_in = gv*(-v_in*lam) (hard) / gv*(-theta) (soft)
// ============================================================================================
struct LifBwdArgs {
  const float *v_in, *z_in, *v_out, *cur, *lam, *theta, *g_out, *g_v_out, *g_z_out;
  float *g_cur, *g_v_in, *g_z_in;
  float* part;  // [2][C][n_part] partial sums of dlam, dtheta
  int B, C, HW, n_chunk, hard_reset, detach, write_gz, surrogate;
  float width;
};

constexpr int EW_THREADS = 256;
constexpr int EW_PER_THREAD = 4;

__global__ void __launch_bounds__(EW_THREADS) convlif_bwd_pointwise_kernel(LifBwdArgs a) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float lam = a.lam[c], theta = a.theta[c];
  const size_t base = ((size_t)b * a.C + c) * a.HW;
  float s_lam = 0.f, s_theta = 0.f;
  const int p0 = blockIdx.x * (EW_THREADS * EW_PER_THREAD);
#pragma unroll
  for (int k = 0; k < EW_PER_THREAD; ++k) {
    const int p = p0 + k * EW_THREADS + threadIdx.x;
    if (p >= a.HW) continue;
    const size_t i = base + p;
    const float vo = a.v_out[i];
    float gz = a.g_out ? a.g_out[i] : 0.f;
    if (a.g_z_out) gz += a.g_z_out[i];
    const float gs = gz * surrogate(vo - theta, a.width, a.surrogate);
    const float gv = (a.g_v_out ? a.g_v_out[i] : 0.f) + gs;
    const float vi = a.v_in ? a.v_in[i] : 0.f;
    const float zi = a.z_in ? a.z_in[i] : 0.f;
    const float cur = a.cur[i];
    a.g_cur[i] = gv * (1.0f - lam);
    float gzr;
    if (a.hard_reset) {
      a.g_v_in[i] = gv * lam * (1.0f - zi);
      s_lam += gv * (vi * (1.0f - zi) - cur);
      s_theta -= gs;
      gzr = -gv * vi * lam;
    } else {
      a.g_v_in[i] = gv * lam;
      s_lam += gv * (vi - cur);
      s_theta -= gs + gv * zi;
      gzr = -gv * theta;
    }
    if (a.write_gz) a.g_z_in[i] = a.detach ? 0.f : gzr;
  }
  // block reduction in a fixed order (deterministic)
  __shared__ float red[2][EW_THREADS / 32];
  s_lam = warp_sum(s_lam);
  s_theta = warp_sum(s_theta);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s_lam; red[1][warp] = s_theta; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < EW_THREADS / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    const int n_part = a.B * a.n_chunk;
    const int j = b * a.n_chunk + blockIdx.x;
    a.part[(size_t)c * n_part + j] = t0;
    a.part[(size_t)(a.C + c) * n_part + j] = t1;
  }
}

// ============================================================================================
// Backward phase B: data gradient = 3x3 conv of g_I with the transposed, flipped weights.
// ============================================================================================
struct DgradArgs {
  ConvSrc src[2];
  float* out;      // [B, n_out, H, W]
  int accumulate;  // out += conv (reset-path gradient already stored there)
  int B, n_out, H, W;
};

__global__ void __launch_bounds__(CT_THREADS) conv3x3_plain_simt_kernel(DgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int n_cog = (a.n_out + CT_CO - 1) / CT_CO;
  const int b = blockIdx.z / n_cog, co0 = (blockIdx.z % n_cog) * CT_CO;
  const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;
  float acc[4][CT_CO];
  conv3x3_tile(a.src, 1, b, co0, a.n_out, a.H, a.W, y0, x0, smem, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = x0 + lane;
  if (x >= a.W) return;
  const size_t plane = (size_t)a.H * a.W;
#pragma unroll
  for (int c = 0; c < CT_CO; ++c) {
    const int co = co0 + c;
    if (co >= a.n_out) break;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = y0 + warp * 4 + r;
      if (y >= a.H) continue;
      const size_t idx = ((size_t)b * a.n_out + co) * plane + (size_t)y * a.W + x;
      a.out[idx] = a.accumulate ? a.out[idx] + acc[r][c] : acc[r][c];
    }
  }
}

// ============================================================================================
// Backward phase C: weight gradient  dW[co][ci][ky][kx] = sum_{b,y,x} g_I[b,co,y,x] * X[b,ci,y+ky-1,x+kx-1]
// One CTA = 32 co x 32 ci x 9 taps of partial sums kept in registers (thread = 4 co x 2 ci), looping
// over pixel tiles of 4 rows x 32 cols staged in shared memory.  Each CTA writes its partial
// [C][Cin][9] block; wgrad_reduce_kernel sums the partials in a fixed order.
// ============================================================================================
constexpr int WG_THREADS = 128;
constexpr int WG_CO = 32, WG_CI = 32;
constexpr int WG_TW = 32, WG_TH = 4;
constexpr int WG_X_ROW = WG_TW + 2;                 // 34
constexpr int WG_X_CH = 6 * WG_X_ROW + 13;          // 217 == 25 (mod 32): conflict-free across ci
constexpr int WG_G_CH = WG_TH * WG_TW;              // 128
constexpr size_t WG_SMEM_BYTES = (WG_CI * WG_X_CH + WG_CO * WG_G_CH) * sizeof(float);

struct WgradArgs {
  const float* g_cur;  // [B, C, H, W]
  const float* xsrc[2];
  int n_ci[2];
  float* part[2];      // per source: [n_cta][C][n_ci][9]
  int B, C, H, W, n_src;
};

__global__ void __launch_bounds__(WG_THREADS) wgrad_simt_kernel(WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sx = smem;
  float* sg = smem + WG_CI * WG_X_CH;
  const int s = blockIdx.z;
  const float* X = a.xsrc[s];
  const int n_ci = a.n_ci[s];
  const int n_cib = (n_ci + WG_CI - 1) / WG_CI;
  const int cob = blockIdx.y / n_cib, cib = blockIdx.y % n_cib;
  const int co_base = cob * WG_CO, ci_base = cib * WG_CI;
  const int tid = threadIdx.x;
  const int cog = tid >> 4, cig = tid & 15;   // thread owns co = co_base + 4*cog + {0..3}, ci = ci_base + 2*cig + {0,1}

  float acc[4][2][9];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[i][j][t] = 0.f;

  const int tiles_x = (a.W + WG_TW - 1) / WG_TW, tiles_y = (a.H + WG_TH - 1) / WG_TH;
  const int n_tiles = a.B * tiles_y * tiles_x;
  const size_t plane = (size_t)a.H * a.W;

  if (X != nullptr) {
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / (tiles_y * tiles_x);
      const int rem = tile - b * (tiles_y * tiles_x);
      const int y0 = (rem / tiles_x) * WG_TH, x0 = (rem % tiles_x) * WG_TW;
      __syncthreads();
      // stage X tile with halo: [ci][6][34]
      for (int i = tid; i < WG_CI * 6 * WG_X_ROW; i += WG_THREADS) {
        int ci = i / (6 * WG_X_ROW);
        int r2 = i - ci * (6 * WG_X_ROW);
        int dy = r2 / WG_X_ROW, dx = r2 - dy * WG_X_ROW;
        int y = y0 - 1 + dy, x = x0 - 1 + dx, ch = ci_base + ci;
        float v = 0.f;
        if (ch < n_ci && y >= 0 && y < a.H && x >= 0 && x < a.W) v = X[((size_t)b * n_ci + ch) * plane + (size_t)y * a.W + x];
        sx[ci * WG_X_CH + dy * WG_X_ROW + dx] = v;
      }
      // stage g_I tile: [co][4][32]
      for (int i = tid; i < WG_CO * WG_G_CH; i += WG_THREADS) {
        int co = i / WG_G_CH;
        int r2 = i - co * WG_G_CH;
        int dy = r2 / WG_TW, dx = r2 - dy * WG_TW;
        int y = y0 + dy, x = x0 + dx, ch = co_base + co;
        float v = 0.f;
        if (ch < a.C && y < a.H && x < a.W) v = a.g_cur[((size_t)b * a.C + ch) * plane + (size_t)y * a.W + x];
        sg[i] = v;
      }
      __syncthreads();
      const float* px0 = sx + (2 * cig) * WG_X_CH;
      const float* px1 = px0 + WG_X_CH;
      const float* pg = sg + (4 * cog) * WG_G_CH;
#pragma unroll 1
      for (int r = 0; r < WG_TH; ++r) {
        float w0[3][3], w1[3][3];  // sliding 3x3 windows of the two input channels
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          w0[dy][1] = px0[(r + dy) * WG_X_ROW + 0]; w0[dy][2] = px0[(r + dy) * WG_X_ROW + 1];
          w1[dy][1] = px1[(r + dy) * WG_X_ROW + 0]; w1[dy][2] = px1[(r + dy) * WG_X_ROW + 1];
        }
#pragma unroll 4
        for (int c = 0; c < WG_TW; ++c) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            w0[dy][0] = w0[dy][1]; w0[dy][1] = w0[dy][2]; w0[dy][2] = px0[(r + dy) * WG_X_ROW + c + 2];
            w1[dy][0] = w1[dy][1]; w1[dy][1] = w1[dy][2]; w1[dy][2] = px1[(r + dy) * WG_X_ROW + c + 2];
          }
          float g[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) g[i] = pg[i * WG_G_CH + r * WG_TW + c];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              acc[i][0][t] = fmaf(g[i], w0[t / 3][t % 3], acc[i][0][t]);
              acc[i][1][t] = fmaf(g[i], w1[t / 3][t % 3], acc[i][1][t]);
            }
        }
      }
    }
  }
  float* part = a.part[s] + (size_t)blockIdx.x * a.C * n_ci * 9;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co_base + 4 * cog + i;
    if (co >= a.C) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ci = ci_base + 2 * cig + j;
      if (ci >= n_ci) continue;
#pragma unroll
      for (int t = 0; t < 9; ++t) part[((size_t)co * n_ci + ci) * 9 + t] = acc[i][j][t];
    }
  }
}

// Sums partial blocks in a fixed order and accumulates into the destination:
//   job 0/1: dW_ff / dW_rec += sum_p part[p][i]      (thread per element)
//   job 2:   dlam / dtheta  += sum_j part[c][j]      (warp per channel, fixed shuffle tree)
struct ReduceArgs {
  const float* wpart[2];
  float* wdst[2];
  int wcount[2];
  int n_wpart;
  const float* cpart;  // [2][C][n_cpart]
  float* cdst[2];      // dlam, dtheta
  int C, n_cpart;
};

__global__ void __launch_bounds__(256) bwd_reduce_kernel(ReduceArgs a) {
  // block = 32 (elements) x 8 (partial-stripes); every thread sums a fixed stripe of partials, then the 8
  // stripes are added in a fixed order: deterministic, coalesced, and short dependent chains.
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, stripe = threadIdx.x >> 5;
  const int job = blockIdx.y;
  if (job < 2) {
    if (a.wdst[job] == nullptr) return;
    const int n = a.wcount[job];
    for (int i0 = blockIdx.x * 32; i0 < n; i0 += gridDim.x * 32) {
      const int i = i0 + lane;
      float s = 0.f;
      if (i < n) {
#pragma unroll 4
        for (int p = stripe; p < a.n_wpart; p += 8) s += a.wpart[job][(size_t)p * n + i];
      }
      red[stripe][lane] = s;
      __syncthreads();
      if (stripe == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        a.wdst[job][i] += t;
      }
      __syncthreads();
    }
  } else {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp_global; r < 2 * a.C; r += n_warps) {
      float s = 0.f;
      for (int j = lane; j < a.n_cpart; j += 32) s += a.cpart[(size_t)r * a.n_cpart + j];
      s = warp_sum(s);
      if (lane == 0) {
        float* dst = a.cdst[r / a.C];
        if (dst) dst[r % a.C] += s;
      }
    }
  }
}

static int wgrad_grid_x(int B, int H, int W) {
  int n_tiles = B * ceil_div(H, WG_TH) * ceil_div(W, WG_TW);
  int cap = 2 * sm_count();
  return n_tiles < cap ? n_tiles : cap;
}

struct BwdLayout {
  size_t off_gcur, off_cpart, off_wpart0, off_wpart1, off_dgblob, total;
  int n_chunk, gx;   // gx: grid of the CUDA-core wgrad; the partial buffers hold max(gx, tensor-core grid) blocks
};

static BwdLayout bwd_layout(int B, int Cin, int C, int H, int W, int recurrent) {
  BwdLayout L;
  size_t n = (size_t)B * C * H * W;
  L.n_chunk = ceil_div(H * W, EW_THREADS * EW_PER_THREAD);
  L.gx = wgrad_grid_x(B, H, W);
  size_t o = 0;
  L.off_gcur = o; o += align_up(n * sizeof(float), 256);
  L.off_cpart = o; o += align_up((size_t)2 * C * B * L.n_chunk * sizeof(float), 256);
  const int tcg = wgrad_tc_grid(B, H, W);
  const int n_part = L.gx > tcg ? L.gx : tcg;
  L.off_wpart0 = o; o += align_up((size_t)n_part * C * Cin * 9 * sizeof(float), 256);
  L.off_wpart1 = o; if (recurrent) o += align_up((size_t)n_part * C * C * 9 * sizeof(float), 256);
  L.off_dgblob = o; o += dgrad_tc_workspace_bytes(Cin, C, recurrent ? C : 0);
  L.total = o;
  return L;
}

}  // namespace snnflow

using namespace snnflow;

extern "C" int snnflow_convlif_fwd(const float* x, const float* w_ff, const float* w_rec, const float* v_in,
                                   const float* z_in, const float* lam, const float* theta,
                                   const float* residual, float* v_out, float* z_out, float* out, float* cur_out,
                                   int B, int Cin, int C, int H, int W, unsigned flags, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(x && w_ff && lam && theta && v_out && z_out, "null pointer");
  SNNFLOW_REQUIRE((v_in == nullptr) == (z_in == nullptr), "v_in and z_in must both be given or both be NULL");
  SNNFLOW_REQUIRE(B > 0 && Cin > 0 && C > 0 && H > 0 && W > 0, "bad dims");
  SNNFLOW_REQUIRE(!(residual && !out), "residual given without out");
  LifFwdArgs a{};
  a.src[0] = ConvSrc{x, w_ff, Cin, 9, Cin * 9, 0};
  a.n_src = 1;
  if (w_rec && z_in) {  // z_in == NULL means zeros: the recurrent current vanishes
    a.src[1] = ConvSrc{z_in, w_rec, C, 9, C * 9, 0};
    a.n_src = 2;
  }
  a.v_in = v_in; a.z_in = z_in; a.lam = lam; a.theta = theta; a.residual = residual;
  a.v_out = v_out; a.z_out = z_out; a.out = out; a.cur_out = cur_out;
  a.B = B; a.C = C; a.H = H; a.W = W;
  a.hard_reset = (flags & SNNFLOW_HARD_RESET) ? 1 : 0;
  static bool attr_done = false;
  if (!attr_done) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(convlif_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)CT_SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(ceil_div(W, CT_W), ceil_div(H, CT_H), B * ceil_div(C, CT_CO));
  {
    const double px = (double)B * H * W;
    const int planes = Cin + 2 * C + (v_in ? 2 * C : 0) + (out ? C : 0) + (residual ? C : 0) + (cur_out ? C : 0);
    prof_begin("convlif_fwd_simt", (cudaStream_t)stream, 4.0 * px * planes, 18.0 * px * C * (Cin + (a.n_src == 2 ? C : 0)));
  }
  convlif_fwd_simt_kernel<<<grid, CT_THREADS, CT_SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return check_launch("convlif_fwd_simt_kernel");
}

extern "C" size_t snnflow_convlif_bwd_workspace_bytes(int B, int Cin, int C, int H, int W, int recurrent) {
  if (B <= 0 || Cin <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return bwd_layout(B, Cin, C, H, W, recurrent).total;
}

static int launch_dgrad(const float* g_cur, const float* w, int n_in_orig, float* out, int accumulate, int B, int C,
                        int H, int W, cudaStream_t st) {
  // out[b, ci, y, x] = sum_{co, tap} g_cur[b, co, ...] * w[co][ci][8 - tap]   (w is [C][n_in_orig][9])
  DgradArgs d{};
  d.src[0] = ConvSrc{g_cur, w, C, n_in_orig * 9, 9, 1};
  d.out = out; d.accumulate = accumulate; d.B = B; d.n_out = n_in_orig; d.H = H; d.W = W;
  static bool attr_done = false;
  if (!attr_done) {
    SNNFLOW_CUDA(cudaFuncSetAttribute(conv3x3_plain_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)CT_SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(ceil_div(W, CT_W), ceil_div(H, CT_H), B * ceil_div(n_in_orig, CT_CO));
  {
    const double px = (double)B * H * W;
    prof_begin("dgrad_simt", st, 4.0 * px * (C + n_in_orig * (accumulate ? 2 : 1)), 18.0 * px * C * n_in_orig);
  }
  conv3x3_plain_simt_kernel<<<grid, CT_THREADS, CT_SMEM_BYTES, st>>>(d);
  return check_launch("conv3x3_plain_simt_kernel");
}

extern "C" int snnflow_convlif_bwd(const float* x, const float* w_ff, const float* w_rec, const float* v_in,
                                   const float* z_in, const float* v_out, const float* cur, const float* lam,
                                   const float* theta, const float* g_out, const float* g_v_out,
                                   const float* g_z_out, float* g_x, float* g_v_in, float* g_z_in, float* dw_ff,
                                   float* dw_rec, float* dlam, float* dtheta, void* workspace,
                                   size_t workspace_bytes, int B, int Cin, int C, int H, int W, unsigned flags,
                                   int surrogate, float act_width, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(x && w_ff && v_out && cur && lam && theta && g_v_in && workspace, "null pointer");
  SNNFLOW_REQUIRE((v_in == nullptr) == (z_in == nullptr), "v_in and z_in must both be given or both be NULL");
  SNNFLOW_REQUIRE(B > 0 && Cin > 0 && C > 0 && H > 0 && W > 0, "bad dims");
  SNNFLOW_REQUIRE(surrogate >= 0 && surrogate <= 3, "unknown surrogate");
  const int recurrent = w_rec != nullptr;
  const int detach = (flags & SNNFLOW_DETACH_RESET) ? 1 : 0;
  SNNFLOW_REQUIRE(g_z_in || (!recurrent && detach), "g_z_in required for recurrent cells / non-detached reset");
  SNNFLOW_REQUIRE(!recurrent || dw_rec, "dw_rec required for recurrent cells");
  SNNFLOW_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  BwdLayout L = bwd_layout(B, Cin, C, H, W, recurrent);
  if (workspace_bytes < L.total) {
    set_error("snnflow_convlif_bwd: workspace %zu < %zu", workspace_bytes, L.total);
    return SNNFLOW_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* g_cur = (float*)(ws + L.off_gcur);
  float* cpart = (float*)(ws + L.off_cpart);
  float* wpart0 = (float*)(ws + L.off_wpart0);
  float* wpart1 = (float*)(ws + L.off_wpart1);

  // phase A
  LifBwdArgs p{};
  p.v_in = v_in; p.z_in = z_in; p.v_out = v_out; p.cur = cur; p.lam = lam; p.theta = theta;
  p.g_out = g_out; p.g_v_out = g_v_out; p.g_z_out = g_z_out;
  p.g_cur = g_cur; p.g_v_in = g_v_in; p.g_z_in = g_z_in; p.part = cpart;
  p.B = B; p.C = C; p.HW = H * W; p.n_chunk = L.n_chunk;
  p.hard_reset = (flags & SNNFLOW_HARD_RESET) ? 1 : 0;
  p.detach = detach;
  // g_z_in gets the reset-path term here unless the recurrent dgrad overwrites it anyway
  p.write_gz = (g_z_in != nullptr) && !(recurrent && detach);
  p.surrogate = surrogate; p.width = act_width;
  {
    const int planes = 4 + (g_out ? 1 : 0) + (g_v_out ? 1 : 0) + (g_z_out ? 1 : 0) + (v_in ? 2 : 0) + (p.write_gz ? 1 : 0);
    prof_begin("lif_bwd_pointwise", st, 4.0 * B * C * H * W * planes);
  }
  convlif_bwd_pointwise_kernel<<<dim3(L.n_chunk, C, B), EW_THREADS, 0, st>>>(p);
  int rc = check_launch("convlif_bwd_pointwise_kernel");
  if (rc) return rc;

  // phase B: data gradients (g_x through W_ff, g_z_in through W_rec)
  if (g_x || recurrent) {
    if (!(flags & SNNFLOW_NO_TENSOR_CORES) && dgrad_tc_supported(Cin, C, recurrent ? C : 0, g_x != nullptr)) {
      rc = launch_dgrad_tc(g_cur, w_ff, w_rec, g_x, recurrent ? g_z_in : nullptr, detach ? 0 : 1, ws + L.off_dgblob, B,
                           Cin, C, recurrent ? C : 0, H, W, st);
      if (rc) return rc;
    } else {
      if (g_x) {
        rc = launch_dgrad(g_cur, w_ff, Cin, g_x, 0, B, C, H, W, st);
        if (rc) return rc;
      }
      if (recurrent) {
        rc = launch_dgrad(g_cur, w_rec, C, g_z_in, detach ? 0 : 1, B, C, H, W, st);
        if (rc) return rc;
      }
    }
  }

  // phase C: weight gradients (partials) + fixed-order reduction
  int n_wpart = L.gx;
  const bool use_tc = (flags & SNNFLOW_INPUT_EXACT16) && !(flags & SNNFLOW_NO_TENSOR_CORES) &&
                      wgrad_tc_supported(Cin, C, recurrent);
  if (use_tc) {
    // x (tagged exact by the caller) and z_prev (spikes) are exact in bf16; g_I is split hi + lo
    rc = launch_wgrad_tc(g_cur, x, recurrent ? z_in : nullptr, wpart0, wpart1, B, Cin, C, H, W, st);
    if (rc) return rc;
    n_wpart = wgrad_tc_grid(B, H, W);
  } else {
    WgradArgs w{};
    w.g_cur = g_cur; w.xsrc[0] = x; w.n_ci[0] = Cin; w.part[0] = wpart0;
    w.xsrc[1] = recurrent ? z_in : nullptr; w.n_ci[1] = C; w.part[1] = wpart1;
    w.B = B; w.C = C; w.H = H; w.W = W; w.n_src = (recurrent && z_in) ? 2 : 1;
    static bool attr_done = false;
    if (!attr_done) {
      SNNFLOW_CUDA(cudaFuncSetAttribute(wgrad_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)WG_SMEM_BYTES));
      attr_done = true;
    }
    for (int s = 0; s < w.n_src; ++s) {
      // one launch per source: the (co-block, ci-block) grid differs
      WgradArgs ws1 = w;
      if (s == 1) { ws1.xsrc[0] = w.xsrc[1]; ws1.n_ci[0] = w.n_ci[1]; ws1.part[0] = w.part[1]; }
      dim3 grid(L.gx, ceil_div(C, WG_CO) * ceil_div(ws1.n_ci[0], WG_CI), 1);
      {
        const double px = (double)B * H * W;
        prof_begin("wgrad_simt", st, 4.0 * px * (C + ws1.n_ci[0]) + 4.0 * L.gx * C * ws1.n_ci[0] * 9,
                   18.0 * px * C * ws1.n_ci[0]);
      }
      wgrad_simt_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, st>>>(ws1);
      rc = check_launch("wgrad_simt_kernel");
      if (rc) return rc;
    }
  }
  ReduceArgs r{};
  r.wpart[0] = wpart0; r.wdst[0] = dw_ff; r.wcount[0] = C * Cin * 9;
  // with a zero initial state (z_in == NULL) the recurrent weight gradient of this step vanishes
  r.wpart[1] = wpart1; r.wdst[1] = (recurrent && z_in) ? dw_rec : nullptr; r.wcount[1] = C * C * 9;
  r.n_wpart = n_wpart;
  r.cpart = cpart; r.cdst[0] = dlam; r.cdst[1] = dtheta; r.C = C; r.n_cpart = B * L.n_chunk;
  prof_begin("bwd_reduce", st, 4.0 * n_wpart * C * (Cin + (recurrent ? C : 0)) * 9 + 8.0 * C * B * L.n_chunk);
  bwd_reduce_kernel<<<dim3(ceil_div(C * (Cin > C ? Cin : C) * 9, 32), 3), 256, 0, st>>>(r);
  return check_launch("bwd_reduce_kernel");
}

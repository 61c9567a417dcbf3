// Layer-major window engine: CUDA-core kernels (packing, time-fused pointwise BPTT, flow head on planes, reductions).
#include <cuda_bf16.h>

#include "window.cuh"

#include <stdlib.h>

namespace snnflow {

// ---- fp32 NCHW -> bf16 planes ------------------------------------------------------------------------
// thread = four consecutive pixels of one (image, chunk): reads up to 8 channel planes (float4, coalesced along x), writes
// four 16-B slots (64 contiguous bytes)
// (PX = 1: any width / alignment)
template <int PX>
__global__ void __launch_bounds__(256) pack_planes_kernel(const float* __restrict__ in, unsigned char* __restrict__ planes,
                                                          int n_ch, int n_chunks, int H, int W,
                                                          unsigned int* __restrict__ inexact) {
  const int HW = H * W, Wp = W + 2;
  const int p = (blockIdx.x * 256 + threadIdx.x) * PX;
  const int chunk = blockIdx.y, img = blockIdx.z;
  if (p >= HW) return;
  const int y = p / W, x = p - y * W;   // W is a multiple of 4: the four pixels share a row
  uint32_t u[PX][4] = {};
  unsigned int bad = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ch = chunk * 8 + c;
    float v[PX] = {};
    if (in != nullptr && ch < n_ch) {
      const float* src = in + ((size_t)img * n_ch + ch) * HW + p;
      if constexpr (PX == 4) {
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(src));
        v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
      } else {
        v[0] = __ldg(src);
      }
    }
#pragma unroll
    for (int k = 0; k < PX; ++k) {
      const __nv_bfloat16 b = __float2bfloat16_rn(v[k]);
      bad += (__bfloat162float(b) != v[k]);
      u[k][c >> 1] |= (uint32_t)__bfloat16_as_ushort(b) << ((c & 1) * 16);
    }
  }
  const size_t plane_bytes = (size_t)(H + 2) * Wp * 16;
  unsigned char* dst = planes + ((size_t)img * n_chunks + chunk) * plane_bytes + ((size_t)(y + 1) * Wp + x + 1) * 16;
#pragma unroll
  for (int k = 0; k < PX; ++k) reinterpret_cast<uint4*>(dst)[k] = make_uint4(u[k][0], u[k][1], u[k][2], u[k][3]);
  // values that one bf16 term does not represent exactly: counted into the arena's sticky status word
  if (inexact != nullptr && bad) atomicAdd(inexact, bad);
}

int launch_pack_input(const float* in, unsigned char* planes, int n_img, int nb, int n_chunks, int H, int W,
                      unsigned int* inexact, cudaStream_t st) {
  prof_begin("win_pack_input", st, (double)n_img * H * W * (4.0 * nb + 16.0 * n_chunks));
  if ((W & 3) == 0 && ((uintptr_t)in & 15) == 0)
    pack_planes_kernel<4><<<dim3(ceil_div(H * W, 1024), n_chunks, n_img), 256, 0, st>>>(in, planes, nb, n_chunks, H, W, inexact);
  else
    pack_planes_kernel<1><<<dim3(ceil_div(H * W, 256), n_chunks, n_img), 256, 0, st>>>(in, planes, nb, n_chunks, H, W, inexact);
  return check_launch("pack_planes_kernel");
}

int launch_pack_spikes(const float* z, unsigned char* planes, int n_img, int C, int H, int W, cudaStream_t st) {
  prof_begin("win_pack_state", st, (double)n_img * H * W * (4.0 * C + 2.0 * C));
  if ((W & 3) == 0 && ((uintptr_t)z & 15) == 0)
    pack_planes_kernel<4><<<dim3(ceil_div(H * W, 1024), C / 8, n_img), 256, 0, st>>>(z, planes, C, C / 8, H, W, nullptr);
  else
    pack_planes_kernel<1><<<dim3(ceil_div(H * W, 256), C / 8, n_img), 256, 0, st>>>(z, planes, C, C / 8, H, W, nullptr);
  return check_launch("pack_planes_kernel");
}

// ---- streaming state: NCHW fp32 (v, z) <-> c8 membranes + bf16 spike planes -----------------------------------------------
// thread = one pixel of one (image, 8-channel chunk); NCHW side coalesced along x, engine side one 32-byte + one 16-byte vector
__global__ void __launch_bounds__(256) state_import_kernel(const float* __restrict__ v_nchw, const float* __restrict__ z_nchw,
                                                           float* __restrict__ v_c8, unsigned char* __restrict__ planes,
                                                           unsigned long long img_stride, int C, int H, int W) {
  const int HW = H * W, Wp = W + 2, nch = C >> 3;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int chunk = blockIdx.y, img = blockIdx.z;
  if (p >= HW) return;
  const int y = p / W, x = p - y * W;
  float v[8];
  uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const size_t o = ((size_t)img * C + chunk * 8 + c) * HW + p;
    v[c] = v_nchw ? __ldg(v_nchw + o) : 0.f;
    const float z = z_nchw ? __ldg(z_nchw + o) : 0.f;
    u[c >> 1] |= (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(z)) << ((c & 1) * 16);
  }
  stg256(v_c8 + (((size_t)img * nch + chunk) * HW + p) * 8, v);
  const size_t plane_bytes = (size_t)(H + 2) * Wp * 16;
  *reinterpret_cast<uint4*>(planes + (size_t)img * img_stride + (size_t)chunk * plane_bytes + ((size_t)(y + 1) * Wp + x + 1) * 16) =
      make_uint4(u[0], u[1], u[2], u[3]);
}

__global__ void __launch_bounds__(256) state_export_kernel(const float* __restrict__ v_c8, const unsigned char* __restrict__ planes,
                                                           unsigned long long img_stride, float* __restrict__ v_nchw,
                                                           float* __restrict__ z_nchw, int C, int H, int W) {
  const int HW = H * W, Wp = W + 2, nch = C >> 3;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int chunk = blockIdx.y, img = blockIdx.z;
  if (p >= HW) return;
  const int y = p / W, x = p - y * W;
  float v[8];
  ldg256_coherent(v_c8 + (((size_t)img * nch + chunk) * HW + p) * 8, v);
  const size_t plane_bytes = (size_t)(H + 2) * Wp * 16;
  const uint4 zz = *reinterpret_cast<const uint4*>(planes + (size_t)img * img_stride + (size_t)chunk * plane_bytes +
                                                   ((size_t)(y + 1) * Wp + x + 1) * 16);
  const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const size_t o = ((size_t)img * C + chunk * 8 + c) * HW + p;
    v_nchw[o] = v[c];
    z_nchw[o] = __bfloat162float(__ushort_as_bfloat16((unsigned short)((zw[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu)));
  }
}

int launch_state_import(const float* v_nchw, const float* z_nchw, float* v_c8, unsigned char* planes, unsigned long long img_stride,
                        int B, int C, int H, int W, cudaStream_t st) {
  prof_begin("win_state_import", st, (double)B * C * H * W * 14.0);
  state_import_kernel<<<dim3(ceil_div(H * W, 256), C / 8, B), 256, 0, st>>>(v_nchw, z_nchw, v_c8, planes, img_stride, C, H, W);
  return check_launch("state_import_kernel");
}

int launch_state_export(const float* v_c8, const unsigned char* planes, unsigned long long img_stride, float* v_nchw, float* z_nchw,
                        int B, int C, int H, int W, cudaStream_t st) {
  prof_begin("win_state_export", st, (double)B * C * H * W * 14.0);
  state_export_kernel<<<dim3(ceil_div(H * W, 256), C / 8, B), 256, 0, st>>>(v_c8, planes, img_stride, v_nchw, z_nchw, C, H, W);
  return check_launch("state_export_kernel");
}

// ---- weights -> UMMA shared-memory images, per-channel parameters ---------------------------------------
// The bf16 terms of a weight sit side by side along the MMA N dimension (column n' = term * N + n), so one MMA per tap
// and k-step produces all partial products and the epilogue adds the column groups.
// forward blob : for conv in (ff[, rec]): [tap][K/8][3C/8][8 n'][8 k] bf16, value(n = co, k = ci) = w[co][ci][tap]
// gradient blob: [tap'][C/8][2N/8][8 n'][8 k] bf16, value(n = ci, k = co) = w[co][ci][8 - tap']
__device__ __forceinline__ void split3(float v, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(a);
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}

__global__ void __launch_bounds__(256) window_pack_weights_kernel(const PackArgs p) {
  const PackLayer& L = p.L[blockIdx.x];
  const int tid = threadIdx.x;
  const int C = L.C;
  // forward
  size_t off = 0;
  for (int conv = 0; conv < (L.w_rec ? 2 : 1); ++conv) {
    const float* w = conv == 0 ? L.w_ff : L.w_rec;
    const int K = conv == 0 ? L.Kin : C, Kreal = conv == 0 ? L.Cin : C;
    const int per_tile = K * C;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(L.fwd_blob + off);
    for (int i = blockIdx.y * 256 + tid; i < 9 * per_tile; i += gridDim.y * 256) {
      const int tap = i / per_tile, r = i - tap * per_tile;
      const int n = r / K, k = r - n * K;
      const float v = k < Kreal ? w[((size_t)n * Kreal + k) * 9 + tap] : 0.f;
      __nv_bfloat16 t0, t1, t2;
      split3(v, t0, t1, t2);
      const __nv_bfloat16 tt[3] = {t0, t1, t2};
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const int nc = term * C + n;
        dst[(size_t)tap * 3 * per_tile + ((k >> 3) * (3 * C >> 3) + (nc >> 3)) * 64 + (nc & 7) * 8 + (k & 7)] = tt[term];
      }
    }
    off += (size_t)9 * 3 * per_tile * 2;
  }
  // gradient blobs (K = co)
  for (int which = 0; which < 2; ++which) {
    unsigned char* blob = which == 0 ? L.dg_blob : L.rb_blob;
    const float* w = which == 0 ? L.w_ff : L.w_rec;
    if (blob == nullptr || w == nullptr) continue;
    const int N = which == 0 ? L.Kin : C, Nreal = which == 0 ? L.Cin : C;
    const int per_tile = N * C;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(blob);
    for (int i = blockIdx.y * 256 + tid; i < 9 * per_tile; i += gridDim.y * 256) {
      const int tap = i / per_tile, r = i - tap * per_tile;
      const int n = r / C, k = r - n * C;
      const float v = n < Nreal ? w[((size_t)k * Nreal + n) * 9 + (8 - tap)] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      const __nv_bfloat16 tt[2] = {hi, lo};
#pragma unroll
      for (int term = 0; term < 2; ++term) {
        const int nc = term * N + n;
        dst[(size_t)tap * 2 * per_tile + ((k >> 3) * (2 * N >> 3) + (nc >> 3)) * 64 + (nc & 7) * 8 + (k & 7)] = tt[term];
      }
    }
  }
  for (int c = tid; c < C && blockIdx.y == 0; c += 256) {
    const float lam = L.leak_lam[c];
    const float oml = __fsub_rn(1.0f, lam);
    // .w = 1 / (1 - lam) for the backward's current-free form of d loss / d lam (0 when lam == 1: sigmoid' is 0 too)
    reinterpret_cast<float4*>(L.par)[c] = make_float4(lam, oml, L.theta[c], oml > 0.f ? 1.0f / oml : 0.f);
  }
}

int launch_pack_weights(const PackArgs& p, cudaStream_t st) {
  prof_begin("win_pack_weights", st, 0.0);
  window_pack_weights_kernel<<<dim3(WIN_LAYERS, 16), 256, 0, st>>>(p);
  return check_launch("window_pack_weights_kernel");
}

// ---- time-fused pointwise BPTT of a feed-forward ConvLIF layer --------------------------------------------
// thread = one pixel x 8 channels of one sample, walking t = T-1 .. 0 with the membrane gradient in registers:
//   gs = g_out[t] * sg(v[t] - theta);  gv = carry + gs;  g_I[t] = gv * (1 - lam)  -> bf16 hi/lo planes
//   hard: carry = gv*lam*(1-z_in); dlam += gv*(v_in*(1-z_in) - I[t]); dtheta -= gs
//   soft: carry = gv*lam;          dlam += gv*(v_in - I[t]);          dtheta -= gs + gv*z_in
// with v_in = v[t-1] (the window's initial state at t = 0) and z_in = spike(v_in) (z_init at t = 0).
// The input current I[t] is not stored: from v[t] = lam*a + (1-lam)*I[t] (a = v_in*(1-z_in), hard reset) follows
// a - I[t] = (a - v[t]) / (1-lam), and for the soft reset v_in - I[t] = (v_in - v[t] - z_in*theta) / (1-lam); the
// per-channel factor 1/(1-lam) is applied once to the block sums.
__device__ __forceinline__ void ld8_c8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <int SG, bool HARD, bool TOP>
__global__ void __launch_bounds__(256, 2) pw_seq_kernel(const PwSeqArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int HW = a.H * a.W, Wp = a.W + 2;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const bool ok = p < HW;
  const int y = ok ? p / a.W : 0, x = ok ? p - y * a.W : 0;
  const size_t plane_bytes = (size_t)(a.H + 2) * Wp * 16;
  const int nch = a.C >> 3;
  float lam[8], oml[8], th[8], inv_oml[8], carry[8], s_lam[8], s_th[8], v_cur[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 pr = __ldg(reinterpret_cast<const float4*>(a.par) + chunk * 8 + c);
    lam[c] = pr.x; oml[c] = pr.y; th[c] = pr.z; inv_oml[c] = pr.w;
    carry[c] = 0.f; s_lam[c] = 0.f; s_th[c] = 0.f; v_cur[c] = 0.f;
  }
  // c8 layout: [image][chunk][H*W][8]
  const size_t img_stride = (size_t)nch * HW * 8, px_off = ((size_t)chunk * HW + (ok ? p : 0)) * 8;
  if (ok) ld8_c8(a.v + (size_t)((a.T - 1) * a.B + b) * img_stride + px_off, v_cur);
  float w0[8], w1[8], dw0[8], dw1[8], db0 = 0.f, db1 = 0.f;   // flow head (TOP): weights of this chunk, gradient partials
  if (TOP) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      w0[c] = __ldg(a.pred_w + chunk * 8 + c);
      w1[c] = __ldg(a.pred_w + a.C + chunk * 8 + c);
      dw0[c] = dw1[c] = 0.f;
    }
  }
  for (int t = a.T - 1; t >= 0; --t) {
    const size_t img = (size_t)(t * a.B + b);
    if (ok) {
      float v_in[8], z_in[8], go[8];
      if (TOP) {
        const size_t fo = img * 2 * HW + p;
        const float f0 = __ldg(a.flow + fo), f1 = __ldg(a.flow + fo + HW);
        const float g0 = __ldg(a.g_flow + fo) * (1.0f - f0 * f0), g1 = __ldg(a.g_flow + fo + HW) * (1.0f - f1 * f1);
        db0 += g0; db1 += g1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          go[c] = g0 * w0[c] + g1 * w1[c];
          const float z = (__fsub_rn(v_cur[c], th[c]) > 0.f) ? 1.f : 0.f;   // this bin's spikes: the head's input
          dw0[c] = fmaf(g0, z, dw0[c]);
          dw1[c] = fmaf(g1, z, dw1[c]);
        }
      } else {
        ld8_c8(a.g_out + img * img_stride + px_off, go);
      }
      if (t > 0) {
        ld8_c8(a.v + (img - a.B) * img_stride + px_off, v_in);
#pragma unroll
        for (int c = 0; c < 8; ++c) z_in[c] = (__fsub_rn(v_in[c], th[c]) > 0.f) ? 1.f : 0.f;
      } else {   // the window's initial state: NCHW tensors of the caller
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const size_t i0 = ((size_t)b * a.C + chunk * 8 + c) * HW + p;
          v_in[c] = a.v_init ? __ldg(a.v_init + i0) : 0.f;
          z_in[c] = a.z_init ? __ldg(a.z_init + i0) : 0.f;
        }
      }
      uint32_t hi[4], lo[4];
      float gI[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float gs = go[c] * surrogate_fast<SG>(v_cur[c] - th[c], a.width);
        const float gv = carry[c] + gs;
        gI[c] = gv * oml[c];
        if (HARD) {
          const float omz = 1.0f - z_in[c];
          carry[c] = gv * lam[c] * omz;
          s_lam[c] += gv * (v_in[c] * omz - v_cur[c]);
          s_th[c] -= gs;
        } else {
          carry[c] = gv * lam[c];
          s_lam[c] += gv * (v_in[c] - v_cur[c] - z_in[c] * th[c]);
          s_th[c] -= gs + gv * z_in[c];
        }
        v_cur[c] = v_in[c];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) split_bf16_pair(gI[2 * c], gI[2 * c + 1], hi[c], lo[c]);
      unsigned char* gp = a.gp + img * a.gp_img_stride + (size_t)chunk * plane_bytes + ((size_t)(y + 1) * Wp + x + 1) * 16;
      *reinterpret_cast<uint4*>(gp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(gp + a.gp_term_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  // fixed-order block reduction of the per-channel sums
  __shared__ float red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float l = warp_sum(ok ? s_lam[c] * inv_oml[c] : 0.f), t = warp_sum(ok ? s_th[c] : 0.f);
    if (lane == 0) { red[warp][c] = l; red[warp][8 + c] = t; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int which = threadIdx.x >> 3, c = threadIdx.x & 7;
    const int j = b * gridDim.x + blockIdx.x;
    a.part[((size_t)which * a.C + chunk * 8 + c) * a.n_part + j] = t;
  }
  if (TOP) {   // flow-head gradients: same fixed-order block reduction, rows [dw0 | dw1 | db] of pred_part
    __syncthreads();
    __shared__ float pred[8][18];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float l = warp_sum(ok ? dw0[c] : 0.f), t = warp_sum(ok ? dw1[c] : 0.f);
      if (lane == 0) { pred[warp][c] = l; pred[warp][8 + c] = t; }
    }
    const float e0 = warp_sum(ok ? db0 : 0.f), e1 = warp_sum(ok ? db1 : 0.f);
    if (lane == 0) { pred[warp][16] = e0; pred[warp][17] = e1; }
    __syncthreads();
    if (threadIdx.x < 18) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += pred[w][threadIdx.x];
      const int j = b * gridDim.x + blockIdx.x;
      if (threadIdx.x < 16) {
        const int which = threadIdx.x >> 3, c = threadIdx.x & 7;
        a.pred_part[((size_t)which * a.C + chunk * 8 + c) * a.n_part + j] = t;
      } else if (chunk == 0) {   // every chunk sees the same g_pre: count the bias gradient once
        a.pred_part[((size_t)2 * a.C + (threadIdx.x - 16)) * a.n_part + j] = t;
      }
    }
  }
}

// The top-layer (flow head fused) kernel as launched by default (SNNFLOW_PW2=0 selects pw_seq_kernel<.., TOP> instead):
// ncu shows pw_seq_kernel<.., TOP> long-scoreboard bound at 128 registers per thread with one 32-byte membrane load in
// flight per thread (2.8 TB/s, 143 us); this variant measured 132 us on B200 with identical test results.  Here the per-channel constants (lam, 1-lam, theta,
// 1/(1-lam), head weights) live in shared memory instead of 48 registers, which pays for a one-bin-deeper prefetch: while
// bin t is processed, the membranes of bin t-2 and the flow / flow-gradient values of bin t-1 are already in flight.
// Same arithmetic, same order of operations as pw_seq_kernel.
template <int SG, bool HARD>
__global__ void __launch_bounds__(256, 2) pw_seq_top2_kernel(const PwSeqArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 s_par[8];
  __shared__ float s_w[2][8];
  const int HW = a.H * a.W, Wp = a.W + 2;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x < 8) {
    s_par[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(a.par) + chunk * 8 + threadIdx.x);
    s_w[0][threadIdx.x] = __ldg(a.pred_w + chunk * 8 + threadIdx.x);
    s_w[1][threadIdx.x] = __ldg(a.pred_w + a.C + chunk * 8 + threadIdx.x);
  }
  __syncthreads();
  const bool ok = p < HW;
  const int y = ok ? p / a.W : 0, x = ok ? p - y * a.W : 0;
  const size_t plane_bytes = (size_t)(a.H + 2) * Wp * 16;
  const int nch = a.C >> 3;
  const size_t img_stride = (size_t)nch * HW * 8, px_off = ((size_t)chunk * HW + (ok ? p : 0)) * 8;
  float carry[8], s_lam[8], s_th[8], dw0[8], dw1[8], db0 = 0.f, db1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) carry[c] = s_lam[c] = s_th[c] = dw0[c] = dw1[c] = 0.f;
  float va[8], vb[8], vc[8], fa[4], fb[4];   // v[t], v[t-1], v[t-2] ; (f0, f1, g0, g1) of bins t, t-1
#pragma unroll
  for (int c = 0; c < 8; ++c) va[c] = vb[c] = vc[c] = 0.f;
  fa[0] = fa[1] = fa[2] = fa[3] = fb[0] = fb[1] = fb[2] = fb[3] = 0.f;
  auto load_v = [&](int t, float (&dst)[8]) {   // t = -1: the window's initial membrane (NCHW tensor of the caller, or zero)
    if (t >= 0) {
      ld8_c8(a.v + (size_t)(t * a.B + b) * img_stride + px_off, dst);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c] = a.v_init ? __ldg(a.v_init + ((size_t)b * a.C + chunk * 8 + c) * HW + p) : 0.f;
    }
  };
  auto load_f = [&](int t, float (&dst)[4]) {
    const size_t fo = (size_t)(t * a.B + b) * 2 * HW + p;
    dst[0] = __ldg(a.flow + fo); dst[1] = __ldg(a.flow + fo + HW);
    dst[2] = __ldg(a.g_flow + fo); dst[3] = __ldg(a.g_flow + fo + HW);
  };
  if (ok) {
    load_v(a.T - 1, va);
    load_v(a.T - 2, vb);
    load_f(a.T - 1, fa);
  }
  for (int t = a.T - 1; t >= 0; --t) {
    const size_t img = (size_t)(t * a.B + b);
    if (ok) {
      if (t >= 1) {          // prefetch for the next trip
        load_v(t - 2, vc);
        load_f(t - 1, fb);
      }
      float z0[8];
      if (t == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) z0[c] = a.z_init ? __ldg(a.z_init + ((size_t)b * a.C + chunk * 8 + c) * HW + p) : 0.f;
      }
      const float g0 = fa[2] * (1.0f - fa[0] * fa[0]), g1 = fa[3] * (1.0f - fa[1] * fa[1]);
      db0 += g0; db1 += g1;
      uint32_t hi[4], lo[4];
      float gI[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 pr = s_par[c];   // (lam, 1 - lam, theta, 1 / (1 - lam))
        const float go = g0 * s_w[0][c] + g1 * s_w[1][c];
        const float z = (__fsub_rn(va[c], pr.z) > 0.f) ? 1.f : 0.f;   // this bin's spikes: the head's input
        dw0[c] = fmaf(g0, z, dw0[c]);
        dw1[c] = fmaf(g1, z, dw1[c]);
        const float z_in = t > 0 ? ((__fsub_rn(vb[c], pr.z) > 0.f) ? 1.f : 0.f) : z0[c];
        const float gs = go * surrogate_fast<SG>(va[c] - pr.z, a.width);
        const float gv = carry[c] + gs;
        gI[c] = gv * pr.y;
        if (HARD) {
          const float omz = 1.0f - z_in;
          carry[c] = gv * pr.x * omz;
          s_lam[c] += gv * (vb[c] * omz - va[c]);
          s_th[c] -= gs;
        } else {
          carry[c] = gv * pr.x;
          s_lam[c] += gv * (vb[c] - va[c] - z_in * pr.z);
          s_th[c] -= gs + gv * z_in;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) split_bf16_pair(gI[2 * c], gI[2 * c + 1], hi[c], lo[c]);
      unsigned char* gp = a.gp + img * a.gp_img_stride + (size_t)chunk * plane_bytes + ((size_t)(y + 1) * Wp + x + 1) * 16;
      *reinterpret_cast<uint4*>(gp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(gp + a.gp_term_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
      for (int c = 0; c < 8; ++c) { va[c] = vb[c]; vb[c] = vc[c]; }
#pragma unroll
      for (int c = 0; c < 4; ++c) fa[c] = fb[c];
    }
  }
  // fixed-order block reductions, exactly as in pw_seq_kernel
  __shared__ float red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float l = warp_sum(ok ? s_lam[c] * s_par[c].w : 0.f), t = warp_sum(ok ? s_th[c] : 0.f);
    if (lane == 0) { red[warp][c] = l; red[warp][8 + c] = t; }
  }
  __syncthreads();
  const int j = b * gridDim.x + blockIdx.x;
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int which = threadIdx.x >> 3, c = threadIdx.x & 7;
    a.part[((size_t)which * a.C + chunk * 8 + c) * a.n_part + j] = t;
  }
  __syncthreads();
  __shared__ float pred[8][18];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float l = warp_sum(ok ? dw0[c] : 0.f), t = warp_sum(ok ? dw1[c] : 0.f);
    if (lane == 0) { pred[warp][c] = l; pred[warp][8 + c] = t; }
  }
  const float e0 = warp_sum(ok ? db0 : 0.f), e1 = warp_sum(ok ? db1 : 0.f);
  if (lane == 0) { pred[warp][16] = e0; pred[warp][17] = e1; }
  __syncthreads();
  if (threadIdx.x < 18) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += pred[w][threadIdx.x];
    if (threadIdx.x < 16) {
      const int which = threadIdx.x >> 3, c = threadIdx.x & 7;
      a.pred_part[((size_t)which * a.C + chunk * 8 + c) * a.n_part + j] = t;
    } else if (chunk == 0) {   // every chunk sees the same g_pre: count the bias gradient once
      a.pred_part[((size_t)2 * a.C + (threadIdx.x - 16)) * a.n_part + j] = t;
    }
  }
}

int launch_pw_seq(const PwSeqArgs& a, cudaStream_t st) {
  const int gx = ceil_div(a.H * a.W, 256);
  if (a.n_part != a.B * gx) {
    set_error("launch_pw_seq: n_part mismatch");
    return SNNFLOW_EINVAL;
  }
  prof_begin("win_pw_seq", st, (double)a.T * a.B * a.H * a.W * (8.0 * a.C + (a.g_out ? 4.0 * a.C : 16.0)));   // v in, g_I hi + lo planes out ; g_out or flow + g_flow in
  const dim3 grid(gx, a.C / 8, a.B);
  const bool top = a.g_out == nullptr;
  static const int staged_v2 = [] { const char* v = getenv("SNNFLOW_PW2"); return (v && *v) ? atoi(v) : 1; }();
#define PW_CASE(SGV, HARDV) \
  if (a.surrogate == SGV && (a.hard_reset != 0) == HARDV) { \
    if (top && staged_v2) launch_pdl(pw_seq_top2_kernel<SGV, HARDV>, grid, dim3(256), 0, st, a); \
    else if (top) launch_pdl(pw_seq_kernel<SGV, HARDV, true>, grid, dim3(256), 0, st, a); \
    else launch_pdl(pw_seq_kernel<SGV, HARDV, false>, grid, dim3(256), 0, st, a); \
  }
  PW_CASE(0, true) PW_CASE(0, false) PW_CASE(1, true) PW_CASE(1, false) PW_CASE(2, true) PW_CASE(2, false)
#undef PW_CASE
  return check_launch("pw_seq_kernel");
}

// ---- flow head on spike planes: flow = tanh(conv1x1(z) + b)  (models/submodules.py:96-113) ----------------
constexpr int PP_MAX_C = 64;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

__global__ void __launch_bounds__(256) pred_fwd_planes_kernel(const unsigned char* __restrict__ zp, unsigned long long img_stride,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ flow, int C, int H, int W) {
  __shared__ float sw[2 * PP_MAX_C + 2];
  for (int i = threadIdx.x; i < 2 * C; i += 256) sw[i] = w[i];
  if (threadIdx.x < 2) sw[2 * C + threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int HW = H * W, Wp = W + 2;
  const int p = blockIdx.x * 256 + threadIdx.x, img = blockIdx.y;
  if (p >= HW) return;
  const int y = p / W, x = p - y * W;
  const size_t plane_bytes = (size_t)(H + 2) * Wp * 16;
  const unsigned char* src = zp + (size_t)img * img_stride + ((size_t)(y + 1) * Wp + x + 1) * 16;
  float a0 = 0.f, a1 = 0.f;
  for (int ch = 0; ch < (C >> 3); ++ch) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + ch * plane_bytes)), f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {   // same accumulation order as pred_fwd_kernel (pred.cu)
      a0 = fmaf(f[c], sw[ch * 8 + c], a0);
      a1 = fmaf(f[c], sw[C + ch * 8 + c], a1);
    }
  }
  flow[((size_t)img * 2 + 0) * HW + p] = tanhf(a0 + sw[2 * C]);
  flow[((size_t)img * 2 + 1) * HW + p] = tanhf(a1 + sw[2 * C + 1]);
}

int launch_pred_fwd_planes(const unsigned char* zp, unsigned long long img_stride, const float* w, const float* b,
                           float* flow, int n_img, int C, int H, int W, cudaStream_t st) {
  if (C > PP_MAX_C) {
    set_error("pred planes: C <= 64");
    return SNNFLOW_EINVAL;
  }
  prof_begin("win_pred_fwd", st, (double)n_img * H * W * (2.0 * C + 8.0), 4.0 * n_img * H * W * C);
  pred_fwd_planes_kernel<<<dim3(ceil_div(H * W, 256), n_img), 256, 0, st>>>(zp, img_stride, w, b, flow, C, H, W);
  return check_launch("pred_fwd_planes_kernel");
}

// g_pre = g_flow * (1 - flow^2); g_x[c] = g_pre0*w[0][c] + g_pre1*w[1][c]; per-CTA partials of dw [2][C], db [2].
// Every thread walks PB_PX pixels (256 apart: coalesced) and keeps its 2C + 2 partial sums in registers, so the block
// reduction (66 warp reductions) is paid once per 2048 pixels instead of once per 256.
constexpr int PB_PX = 2;
constexpr int PB_MAX_C = 32;

__global__ void __launch_bounds__(256) pred_bwd_planes_kernel(const unsigned char* __restrict__ zp, unsigned long long img_stride,
                                                              const float* __restrict__ w, const float* __restrict__ flow,
                                                              const float* __restrict__ g_flow, float* __restrict__ g_x,
                                                              float* __restrict__ part, int C, int H, int W) {
  __shared__ float sw[2 * PB_MAX_C];
  __shared__ float red[8][2 * PB_MAX_C + 2];
  for (int i = threadIdx.x; i < 2 * C; i += 256) sw[i] = w[i];
  __syncthreads();
  const int HW = H * W, Wp = W + 2, img = blockIdx.y, nch = C >> 3;
  const size_t plane_bytes = (size_t)(H + 2) * Wp * 16;
  float a0[PB_MAX_C], a1[PB_MAX_C], b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int c = 0; c < PB_MAX_C; ++c) a0[c] = a1[c] = 0.f;
  for (int j = 0; j < PB_PX; ++j) {
    const int p = (blockIdx.x * PB_PX + j) * 256 + threadIdx.x;
    if (p >= HW) break;
    const int y = p / W, x = p - y * W;
    const float f0 = flow[((size_t)img * 2 + 0) * HW + p], f1 = flow[((size_t)img * 2 + 1) * HW + p];
    const float g0 = g_flow[((size_t)img * 2 + 0) * HW + p] * (1.0f - f0 * f0);
    const float g1 = g_flow[((size_t)img * 2 + 1) * HW + p] * (1.0f - f1 * f1);
    b0 += g0; b1 += g1;
    const unsigned char* src = zp + (size_t)img * img_stride + ((size_t)(y + 1) * Wp + x + 1) * 16;
#pragma unroll
    for (int ch = 0; ch < PB_MAX_C / 8; ++ch) {
      if (ch >= nch) break;
      float f[8], gx[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + ch * plane_bytes)), f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        gx[c] = g0 * sw[ch * 8 + c] + g1 * sw[C + ch * 8 + c];
        a0[ch * 8 + c] = fmaf(g0, f[c], a0[ch * 8 + c]);
        a1[ch * 8 + c] = fmaf(g1, f[c], a1[ch * 8 + c]);
      }
      float4* gxp = reinterpret_cast<float4*>(g_x + (((size_t)img * nch + ch) * HW + p) * 8);   // c8 layout
      gxp[0] = make_float4(gx[0], gx[1], gx[2], gx[3]);
      gxp[1] = make_float4(gx[4], gx[5], gx[6], gx[7]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < PB_MAX_C; ++c) {
    if (c >= C) break;
    const float s0 = warp_sum(a0[c]), s1 = warp_sum(a1[c]);
    if (lane == 0) { red[warp][c] = s0; red[warp][C + c] = s1; }
  }
  const float s0 = warp_sum(b0), s1 = warp_sum(b1);
  if (lane == 0) { red[warp][2 * C] = s0; red[warp][2 * C + 1] = s1; }
  __syncthreads();
  float* mypart = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (2 * C + 2);
  for (int i = threadIdx.x; i < 2 * C + 2; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][i];
    mypart[i] = t;
  }
}

int pred_planes_parts(int n_img, int H, int W) { return n_img * ceil_div(H * W, 256 * PB_PX); }

int launch_pred_bwd_planes(const unsigned char* zp, unsigned long long img_stride, const float* w, const float* flow,
                           const float* g_flow, float* g_x, float* part, int n_img, int C, int H, int W,
                           cudaStream_t st) {
  if (C > PB_MAX_C) {
    set_error("pred planes: C <= 32");
    return SNNFLOW_EINVAL;
  }
  prof_begin("win_pred_bwd", st, (double)n_img * H * W * (2.0 * C + 4.0 * C + 16.0), 8.0 * n_img * H * W * C);
  pred_bwd_planes_kernel<<<dim3(ceil_div(H * W, 256 * PB_PX), n_img), 256, 0, st>>>(zp, img_stride, w, flow, g_flow, g_x, part, C, H, W);
  return check_launch("pred_bwd_planes_kernel");
}

// dw[i] += sum_p part[p][i]  (i < 2C), db[i - 2C] += ...   : one warp per output, fixed shuffle tree
__global__ void __launch_bounds__(256) pred_reduce_planes_kernel(const float* __restrict__ part, float* dw, float* db, int C,
                                                                 int n_part) {
  const int n = 2 * C + 2;
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n) return;
  float s = 0.f;
  for (int p = lane; p < n_part; p += 32) s += part[(size_t)p * n + wid];
  s = warp_sum(s);
  if (lane == 0) {
    if (wid < 2 * C) { if (dw) dw[wid] += s; }
    else if (db) db[wid - 2 * C] += s;
  }
}

// same for the row-major partials [2C + 2][n_part] written by the fused top-layer pointwise kernel
__global__ void __launch_bounds__(256) pred_reduce_rows_kernel(const float* __restrict__ part, float* dw, float* db, int C, int n_part) {
  const int n = 2 * C + 2;
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n) return;
  float s = 0.f;
  for (int p = lane; p < n_part; p += 32) s += part[(size_t)wid * n_part + p];
  s = warp_sum(s);
  if (lane == 0) {
    if (wid < 2 * C) { if (dw) dw[wid] += s; }
    else if (db) db[wid - 2 * C] += s;
  }
}

int launch_pred_reduce_rows(const float* part, float* dw, float* db, int C, int n_part, cudaStream_t st) {
  prof_begin("win_pred_reduce", st, 4.0 * n_part * (2 * C + 2));
  pred_reduce_rows_kernel<<<ceil_div((2 * C + 2) * 32, 256), 256, 0, st>>>(part, dw, db, C, n_part);
  return check_launch("pred_reduce_rows_kernel");
}

int launch_pred_reduce_planes(const float* part, float* dw, float* db, int C, int n_part, cudaStream_t st) {
  prof_begin("win_pred_reduce", st, 4.0 * n_part * (2 * C + 2));
  pred_reduce_planes_kernel<<<ceil_div((2 * C + 2) * 32, 256), 256, 0, st>>>(part, dw, db, C, n_part);
  return check_launch("pred_reduce_planes_kernel");
}

// ---- reductions of the per-CTA partials (fixed order: run-to-run deterministic) --------------------------
//   blockIdx.y 0/1: dW_ff / dW_rec[co][ci][tap] += sum_p part[p][tap][ci][co]
//   blockIdx.y 2  : dlam / dtheta[c] += sum_j cpart
//   blockIdx.z    : layer - the reductions of all layers of a window are ONE launch at the end of the backward pass
//                   (every layer keeps its own partial blocks in the workspace): 7 x 12 us of launch latency -> one
constexpr int WIN_REDUCE_MAX_LAYERS = 8;
struct WinReduceBatch {
  WinReduceArgs layer[WIN_REDUCE_MAX_LAYERS];
};

__global__ void __launch_bounds__(256) win_reduce_kernel(const __grid_constant__ WinReduceBatch batch) {
  pdl_launch_dependents();
  pdl_wait();
  const WinReduceArgs& a = batch.layer[blockIdx.z];
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, stripe = threadIdx.x >> 5;
  const int job = blockIdx.y;
  if (job < 2) {
    if (a.wdst[job] == nullptr) return;
    const int C = a.C, cr = a.cin_real[job], ca = a.cin_alloc[job];
    const int n = C * cr * 9;                 // outputs
    const size_t pstride = (size_t)9 * ca * C;
    for (int i0 = blockIdx.x * 32; i0 < n; i0 += gridDim.x * 32) {
      // consecutive lanes take consecutive co (contiguous in the partials)
      const int i = i0 + lane;
      const int co = i % C, rest = i / C, ci = rest % cr, tap = rest / cr;
      float s = 0.f;
      if (i < n) {
        // all partials of a thread in flight at once (20 x 8 stripes covers one partial block per SM): the kernel is
        // a handful of dependent DRAM round trips, not bandwidth (45 MB); the sum keeps its order
        const float* src = a.wpart[job] + ((size_t)tap * ca + ci) * C + co;
#pragma unroll 1
        for (int p0 = stripe; p0 < a.n_wpart; p0 += 8 * 20) {
          float v[20];
#pragma unroll
          for (int k = 0; k < 20; ++k) {
            const int p = p0 + 8 * k;
            v[k] = p < a.n_wpart ? __ldg(src + (size_t)p * pstride) : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 20; ++k) s += v[k];
        }
      }
      red[stripe][lane] = s;
      __syncthreads();
      if (stripe == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        a.wdst[job][((size_t)co * cr + ci) * 9 + tap] += t;
      }
      __syncthreads();
    }
  } else {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp_global; r < 2 * a.C; r += n_warps) {
      float s = 0.f;
      const float* src = a.cpart_layout == 0 ? a.cpart + (size_t)r * a.n_cpart : a.cpart + r;
      const size_t js = a.cpart_layout == 0 ? 1 : (size_t)2 * a.C;
#pragma unroll 1
      for (int j0 = lane; j0 < a.n_cpart; j0 += 32 * 16) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int j = j0 + 32 * k;
          v[k] = j < a.n_cpart ? __ldg(src + (size_t)j * js) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s += v[k];
      }
      s = warp_sum(s);
      if (lane == 0) {
        const int c = r % a.C;
        float* dst = r < a.C ? a.dlam : a.dtheta;
        if (dst) dst[c] += s;
        if (r < a.C) {
          if (a.d_leak) {                                                          // sigmoid'
            const float lam = a.lam[c];
            a.d_leak[c] += __fmul_rn(__fmul_rn(s, lam), __fsub_rn(1.0f, lam));
          }
        } else if (a.d_thresh) {                                                   // clamp_min'
          a.d_thresh[c] += (a.thresh_raw[c] >= 0.01f) ? s : 0.f;
        }
      }
    }
  }
}

int launch_win_reduce(const WinReduceArgs* layers, int n_layers, cudaStream_t st) {
  if (n_layers < 1 || n_layers > WIN_REDUCE_MAX_LAYERS) {
    set_error("launch_win_reduce: %d layers", n_layers);
    return SNNFLOW_EINVAL;
  }
  WinReduceBatch batch{};
  int gx = 1;
  double bytes = 0.0;
  for (int l = 0; l < n_layers; ++l) {
    const WinReduceArgs& a = layers[l];
    batch.layer[l] = a;
    const int cmax = a.cin_real[0] > a.cin_real[1] ? a.cin_real[0] : a.cin_real[1];
    const int g = ceil_div(a.C * cmax * 9, 32);
    if (g > gx) gx = g;
    bytes += 4.0 * a.n_wpart * 9.0 * a.C * (a.cin_alloc[0] + (a.wdst[1] ? a.cin_alloc[1] : 0)) + 8.0 * a.C * a.n_cpart;
  }
  if (gx > 4 * sm_count()) gx = 4 * sm_count();
  prof_begin("win_reduce", st, bytes);
  launch_pdl(win_reduce_kernel, dim3(gx, 3, n_layers), dim3(256), 0, st, batch);
  return check_launch("win_reduce_kernel");
}

}  // namespace snnflow

// Exact-fp32 3x3 convolution (stride 1, pad 1) on the CUDA cores with a fused epilogue.
//
// This is the general path of the ConvLIF kernels: it takes any fp32 input (event counts,
// avg-pooled counts, gradients) and any channel counts.  The tensor-core path (conv3x3_tc.cuh)
// handles the C->C spike layers; this one handles the head layer (Cin = 2, K = 18: too thin for an
// MMA tile), the backward data-gradient and every shape the tensor-core path does not cover.
//
// Tiling: one CTA = 16 rows x 32 cols of output pixels of one sample x CO_T = 16 output channels;
// 128 threads: lane = column, warp = 4-row band; each thread keeps 4 rows x 16 channels of fp32
// accumulators in registers.  The K loop runs over "virtual" input channels in chunks of 8:
// source 0 is `x` with weights w0, optional source 1 is the recurrent input with weights w1
// (ConvLIFRecurrent sums both currents).  Chunks are staged in shared memory by cp.async, double
// buffered; image borders and channel tails are zero-filled by the copy itself.
#pragma once
#include "common.cuh"

namespace snnflow {

constexpr int CT_W = 32;       // tile width  (= warp size)
constexpr int CT_H = 16;       // tile height (4 warps x 4 rows)
constexpr int CT_CO = 16;      // output channels per CTA
constexpr int CT_CI = 8;       // input channels per K chunk
constexpr int CT_THREADS = 128;
constexpr int CT_IN_W = CT_W + 2;
constexpr int CT_IN_H = CT_H + 2;
constexpr int CT_IN_PITCH = 36;
constexpr int CT_IN_ELEMS = CT_CI * CT_IN_H * CT_IN_PITCH;
constexpr int CT_W_ELEMS = CT_CI * 9 * CT_CO;
constexpr int CT_STAGE_ELEMS = CT_IN_ELEMS + CT_W_ELEMS;
constexpr size_t CT_SMEM_BYTES = 2 * CT_STAGE_ELEMS * sizeof(float);

struct ConvSrc {
  const float* data;   // [B, n_ch, H, W] or nullptr (= all zeros, skipped)
  const float* w;      // weights, addressed as w[in*s_in + out*s_out + tap']
  int n_ch;            // channels of this source
  int s_in, s_out;     // element strides of the weight tensor for the (virtual) in / out channel
  int flip;            // 1: tap' = 8 - tap (transposed convolution for the data gradient)
};

// Accumulates acc[r][c] = sum over sources, channels, taps for the thread's 4 pixels x 16 channels.
__device__ __forceinline__ void conv3x3_tile(const ConvSrc* srcs, int n_src, int b, int co0, int n_out, int H,
                                             int W, int y0, int x0, float* smem, float (&acc)[4][CT_CO]) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CT_CO; ++c) acc[r][c] = 0.f;

  // flatten (source, chunk) into one sequence
  int n_chunks[2] = {0, 0};
  int total = 0;
  for (int s = 0; s < n_src; ++s) {
    n_chunks[s] = srcs[s].data ? (srcs[s].n_ch + CT_CI - 1) / CT_CI : 0;
    total += n_chunks[s];
  }
  if (total == 0) return;

  auto issue = [&](int k, int stage) {
    int s = (k < n_chunks[0]) ? 0 : 1;
    int ci0 = (s == 0 ? k : k - n_chunks[0]) * CT_CI;
    const ConvSrc& S = srcs[s];
    float* sin = smem + stage * CT_STAGE_ELEMS;
    float* sw = sin + CT_IN_ELEMS;
    const size_t plane = (size_t)H * W;
    const float* base = S.data + (size_t)b * S.n_ch * plane;
    for (int i = tid; i < CT_CI * CT_IN_H * CT_IN_W; i += CT_THREADS) {
      int ci = i / (CT_IN_H * CT_IN_W);
      int rem = i - ci * (CT_IN_H * CT_IN_W);
      int dy = rem / CT_IN_W, dx = rem - dy * CT_IN_W;
      int y = y0 - 1 + dy, x = x0 - 1 + dx, ch = ci0 + ci;
      bool ok = (ch < S.n_ch) && (y >= 0) && (y < H) && (x >= 0) && (x < W);
      const float* g = ok ? base + (size_t)ch * plane + (size_t)y * W + x : S.data;
      cp_async4(sin + ci * (CT_IN_H * CT_IN_PITCH) + dy * CT_IN_PITCH + dx, g, ok);
    }
    for (int i = tid; i < CT_W_ELEMS; i += CT_THREADS) {
      int ci = i / (9 * CT_CO);
      int rem = i - ci * (9 * CT_CO);
      int tap = rem / CT_CO, co = rem - tap * CT_CO;
      int ch = ci0 + ci, oc = co0 + co;
      bool ok = (ch < S.n_ch) && (oc < n_out);
      int t2 = S.flip ? 8 - tap : tap;
      const float* g = ok ? S.w + (size_t)ch * S.s_in + (size_t)oc * S.s_out + t2 : S.w;
      cp_async4(sw + i, g, ok);
    }
    cp_async_commit();
  };

  issue(0, 0);
  for (int k = 0; k < total; ++k) {
    if (k + 1 < total) {
      issue(k + 1, (k + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* sin = smem + (k & 1) * CT_STAGE_ELEMS;
    const float* sw = sin + CT_IN_ELEMS;
#pragma unroll 1
    for (int ci = 0; ci < CT_CI; ++ci) {
      float xr[6][3];
      const float* p = sin + ci * (CT_IN_H * CT_IN_PITCH) + (warp * 4) * CT_IN_PITCH + lane;
#pragma unroll
      for (int dy = 0; dy < 6; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) xr[dy][dx] = p[dy * CT_IN_PITCH + dx];
      const float4* wp = reinterpret_cast<const float4*>(sw + ci * 9 * CT_CO);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float wv[CT_CO];
#pragma unroll
          for (int q = 0; q < CT_CO / 4; ++q) {
            float4 t = wp[(ky * 3 + kx) * (CT_CO / 4) + q];
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < CT_CO; ++c) acc[r][c] = fmaf(xr[r + ky][kx], wv[c], acc[r][c]);
        }
    }
    __syncthreads();
  }
}

}  // namespace snnflow

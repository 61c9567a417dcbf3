// snntorch-style Leaky neuron step: the LIF of SNNtorch_ConvLIF / SNNtorch_ConvLIFRecurrent
// (models/SNNtorch_spiking_submodules.py:124-322, :324-567), which wrap snn.Leaky(beta, threshold, reset_mechanism in
// {"zero", "subtract"}, reset_delay=False) around BatchNorm2d(conv(x) [+ conv_rec(z_prev)]).
//
// PARITY UNPINNED: snntorch 0.9.4 (requirements.txt:8) is not installed and its source is not on disk, so the arithmetic
// below restates the published Leaky.forward from memory (the tests compare with a CPU restatement of the same text):
//   beta_c  = clamp(beta, 0, 1);  theta = threshold (clamped to >= 0.01 in place by the cell, :284)
//   reset   = H(mem_in - theta)                                   (detached: the membrane entering the step)
//   zero    : m = beta_c * ((1 - reset) * mem_in) + I             subtract: m = beta_c * mem_in + I - reset * theta
//   spk     = H(m - theta)                                        (surrogate ATan(alpha = 2): d spk / d m = 1 / (1 + (pi (m - theta))^2))
//   no reset delay: do_reset = spk - reset;  zero: mem_out = m - do_reset * m;  subtract: mem_out = m - do_reset * theta
// The cell detaches mem_out (:309-311), so the only differentiable output of a step is spk; I is the batch-normalised input
// current (BatchNorm2d stays a PyTorch op: it couples the whole batch).
// One streaming pass: 4 B/elem in (I) [+ 4 mem_in], 8 out (+ 4 saved m for the backward): HBM bound.
#include "common.cuh"

namespace snnflow {

constexpr int LK_THREADS = 256;
constexpr float LK_PI = 3.14159265358979323846f;

__global__ void __launch_bounds__(LK_THREADS) leaky_fwd_kernel(const float4* __restrict__ cur, const float4* __restrict__ mem_in,
                                                               const float* __restrict__ beta, const float* __restrict__ theta,
                                                               float4* __restrict__ mem_out, float4* __restrict__ spk,
                                                               float4* __restrict__ m_pre, int C, int HW4, int64_t n4, int subtract) {
  const int64_t i = (int64_t)blockIdx.x * LK_THREADS + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)((i / HW4) % C);
  const float b = fminf(fmaxf(__ldg(beta + c), 0.f), 1.f), th = __ldg(theta + c);
  const float4 I = __ldg(cur + i);
  const float4 mi = mem_in ? __ldg(mem_in + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float Iv[4] = {I.x, I.y, I.z, I.w}, mv[4] = {mi.x, mi.y, mi.z, mi.w};
  float mo[4], sp[4], mp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float reset = (__fsub_rn(mv[k], th) > 0.f) ? 1.f : 0.f;
    float m;
    if (subtract) m = __fsub_rn(__fadd_rn(__fmul_rn(b, mv[k]), Iv[k]), __fmul_rn(reset, th));
    else m = __fadd_rn(__fmul_rn(b, __fmul_rn(__fsub_rn(1.f, reset), mv[k])), Iv[k]);
    const float s = (__fsub_rn(m, th) > 0.f) ? 1.f : 0.f;
    const float do_reset = __fsub_rn(s, reset);
    mp[k] = m;
    sp[k] = s;
    mo[k] = subtract ? __fsub_rn(m, __fmul_rn(do_reset, th)) : __fsub_rn(m, __fmul_rn(do_reset, m));
  }
  mem_out[i] = make_float4(mo[0], mo[1], mo[2], mo[3]);
  spk[i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
  if (m_pre) m_pre[i] = make_float4(mp[0], mp[1], mp[2], mp[3]);
}

// g_cur = g_spk * sg(m - theta); per-block partial sums of d beta_c (before the clamp mask) and d theta
//   d m / d beta_c = (1 - reset) mem_in (zero) | mem_in (subtract);  d spk / d theta = -sg (+ subtract: d m / d theta = -reset)
__global__ void __launch_bounds__(LK_THREADS) leaky_bwd_kernel(const float* __restrict__ g_spk, const float* __restrict__ m_pre,
                                                               const float* __restrict__ mem_in, const float* __restrict__ theta,
                                                               float* __restrict__ g_cur, float* __restrict__ part, int C, int HW,
                                                               int n_blk_hw, int subtract) {
  // grid: (n_blk_hw, C, B)
  const int c = blockIdx.y, b = blockIdx.z;
  const float th = __ldg(theta + c);
  const size_t base = ((size_t)b * C + c) * HW;
  float s_b = 0.f, s_t = 0.f;
  for (int p = blockIdx.x * LK_THREADS + threadIdx.x; p < HW; p += n_blk_hw * LK_THREADS) {
    const float m = m_pre[base + p], g = g_spk[base + p];
    const float u = LK_PI * (m - th);
    const float gm = g / (1.f + u * u);
    g_cur[base + p] = gm;
    const float mi = mem_in ? mem_in[base + p] : 0.f;
    const float reset = (__fsub_rn(mi, th) > 0.f) ? 1.f : 0.f;
    s_b += gm * (subtract ? mi : (1.f - reset) * mi);
    s_t -= gm * (subtract ? 1.f + reset : 1.f);
  }
  __shared__ float red[2][LK_THREADS / 32];
  s_b = warp_sum(s_b);
  s_t = warp_sum(s_t);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s_b; red[1][warp] = s_t; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, t = 0.f;
#pragma unroll
    for (int w = 0; w < LK_THREADS / 32; ++w) { a += red[0][w]; t += red[1][w]; }
    float* o = part + (((size_t)c * gridDim.z + b) * n_blk_hw + blockIdx.x) * 2;
    o[0] = a; o[1] = t;
  }
}

// fixed-order reduction of the partials: d beta [C] (masked by the clamp: 0 <= beta <= 1), d theta [C]
__global__ void leaky_reduce_kernel(const float* __restrict__ part, const float* __restrict__ beta, float* __restrict__ d_beta,
                                    float* __restrict__ d_theta, int C, int n_per_c) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, t = 0.f;
  for (int k = 0; k < n_per_c; ++k) { a += part[((size_t)c * n_per_c + k) * 2]; t += part[((size_t)c * n_per_c + k) * 2 + 1]; }
  const float b = beta[c];
  d_beta[c] = (b >= 0.f && b <= 1.f) ? a : 0.f;
  d_theta[c] = t;
}

static int leaky_blocks_hw(int HW) {
  int n = ceil_div(HW, LK_THREADS * 4);
  return n < 1 ? 1 : (n > 64 ? 64 : n);
}

}  // namespace snnflow
using namespace snnflow;

extern "C" int snnflow_leaky_fwd(const float* cur, const float* mem_in, const float* beta, const float* theta, float* mem_out,
                                 float* spk, float* m_pre, int B, int C, int H, int W, int subtract, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(cur && beta && theta && mem_out && spk, "null pointer");
  SNNFLOW_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && ((H * W) & 3) == 0, "H*W must be a multiple of 4");
  SNNFLOW_REQUIRE(((((uintptr_t)cur | (uintptr_t)mem_in | (uintptr_t)mem_out | (uintptr_t)spk | (uintptr_t)m_pre)) & 15) == 0,
                  "tensors must be 16-byte aligned");
  const int64_t n4 = (int64_t)B * C * H * W / 4;
  prof_begin("leaky_fwd", (cudaStream_t)stream, 4.0 * n4 * (4.0 + (mem_in ? 4.0 : 0.0) + 8.0 + (m_pre ? 4.0 : 0.0)));
  leaky_fwd_kernel<<<(unsigned)ceil_div64(n4, LK_THREADS), LK_THREADS, 0, (cudaStream_t)stream>>>(
      (const float4*)cur, (const float4*)mem_in, beta, theta, (float4*)mem_out, (float4*)spk, (float4*)m_pre, C, H * W / 4, n4,
      subtract);
  return check_launch("leaky_fwd_kernel");
}

extern "C" size_t snnflow_leaky_bwd_workspace_bytes(int B, int C, int H, int W) {
  return (size_t)B * C * leaky_blocks_hw(H * W) * 2 * sizeof(float);
}

extern "C" int snnflow_leaky_bwd(const float* g_spk, const float* m_pre, const float* mem_in, const float* beta,
                                 const float* theta, float* g_cur, float* d_beta, float* d_theta, void* workspace,
                                 size_t workspace_bytes, int B, int C, int H, int W, int subtract, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(g_spk && m_pre && beta && theta && g_cur && d_beta && d_theta && workspace, "null pointer");
  SNNFLOW_REQUIRE(workspace_bytes >= snnflow_leaky_bwd_workspace_bytes(B, C, H, W), "workspace too small");
  const int nb = leaky_blocks_hw(H * W);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin("leaky_bwd", st, 4.0 * B * C * H * W * (3.0 + (mem_in ? 1.0 : 0.0)));
  leaky_bwd_kernel<<<dim3(nb, C, B), LK_THREADS, 0, st>>>(g_spk, m_pre, mem_in, theta, g_cur, (float*)workspace, C, H * W, nb, subtract);
  int rc = check_launch("leaky_bwd_kernel");
  if (rc) return rc;
  prof_begin("leaky_reduce", st, 8.0 * B * C * nb);
  leaky_reduce_kernel<<<ceil_div(C, 64), 64, 0, st>>>((const float*)workspace, beta, d_beta, d_theta, C, B * nb);
  return check_launch("leaky_reduce_kernel");
}

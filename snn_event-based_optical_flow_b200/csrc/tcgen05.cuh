// tcgen05 / TMEM / mbarrier / TMA-bulk PTX wrappers and UMMA descriptor helpers shared by the tensor-core
// kernels (sm_100a).  Bit layouts follow cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace snnflow {

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef SNNFLOW_MBAR_HINT_NS
#define SNNFLOW_MBAR_HINT_NS 20000   // suspend-time hint of try_wait; 0 = no hint operand (hardware default)
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if SNNFLOW_MBAR_HINT_NS > 0
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SNNFLOW_MBAR_HINT_NS)   // a waiting warp sleeps instead of polling
      : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok;
}
// try_wait suspends in hardware up to a system time limit per attempt; the attempt counter is a watchdog that turns
// a protocol bug (a barrier that is never completed) into a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA with the two shared-memory descriptors passed as 32-bit halves: in a loop over taps / k-steps only the
// 14-bit start-address field in the low word changes, so the issuing thread spends one integer add per operand.
//   low word : (start >> 4) | (LBO >> 4) << 16        high word: (SBO >> 4) | version 1 << 14
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ void umma_f16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

// asynchronous variants: issue several loads, then one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&u)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// warp-uniform election of one lane (the form ptxas recognises: code under it is known to run in ONE thread, so
// tcgen05.mma / cp.async.bulk issue without a per-thread "waterfall" loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred != 0;
}
// 32 lanes x 8 consecutive fp32 columns -> 8 registers per thread
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor):
//   [0,14) start>>4   [16,30) LBO>>4 = byte distance between the two 8-element K chunks of one MMA
//   [32,46) SBO>>4 = byte distance between 8-row groups   [46,48) version = 1   [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}


// MN-major, no-swizzle descriptor: the MN index runs over 8 contiguous elements (16 B) and then in groups
// `sbo` bytes apart; the K index runs over 8 rows 16 B apart and then in groups `lbo` bytes apart
// (canonical layout ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), cute/atom/mma_traits_sm100.hpp).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) { return make_desc(saddr, lbo, sbo); }

// instruction descriptor for kind::f16: fp32 accumulate; fmt 0 = fp16, 1 = bf16; major 0 = K, 1 = MN
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int fmt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int TC_THREADS = 256;         // 8 warps: TMEM lane quarter = warp & 3, channel half = warp >> 2
constexpr int TC_TW = 128;            // output pixels per tile = UMMA M
constexpr int TC_P = TC_TW + 2;       // padded row pitch in slots
constexpr int TC_SLOTS = 392;         // 3 rows * 130 = 390 slots, padded to a multiple of 8
constexpr int TC_TERMS = 2;           // fp16 hi + lo
constexpr int TC_MAX_C = 64;


// Stage one source (n_chunks 8-channel chunks of an fp32 NCHW tensor) of a row tile into the slot buffer:
// rows y0-1..y0+1, columns x0-1..x0+128 -> 16-bit, slot s = rr*130 + cc holds 8 channels (16 B).
//   MODE 0: fp16 (one term)   MODE 1: bf16 (one term)   MODE 2: bf16 hi + lo (lo planes go to s_lo)
// `inexact` counts values that a single term does not represent exactly (MODE 0/1).
// Vector path: aligned float4 loads of 4 consecutive pixels for 8 channels (8 x 16 B in flight per task).
template <int MODE>
__device__ __forceinline__ void stage_source(const float* __restrict__ src, int n_chunks, unsigned char* s_a, int H, int W,
                                             int y0, int x0, bool vec_ok, unsigned int& inexact,
                                             unsigned char* s_lo = nullptr) {
  const int tid = threadIdx.x;
  const size_t plane = (size_t)H * W;
  auto put = [&](const float (&f)[8], int j, int slot) {
    uint32_t u[4], l[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float2 back;
      if (MODE == 0) {
        __half2 h = __floats2half2_rn(f[2 * c], f[2 * c + 1]);
        back = __half22float2(h);
        u[c] = *reinterpret_cast<uint32_t*>(&h);
      } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * c], f[2 * c + 1]);
        back = __bfloat1622float2(h);
        u[c] = *reinterpret_cast<uint32_t*>(&h);
      }
      if (MODE == 2) {
        __nv_bfloat162 r = __floats2bfloat162_rn(f[2 * c] - back.x, f[2 * c + 1] - back.y);
        l[c] = *reinterpret_cast<uint32_t*>(&r);
      } else {
        inexact += (back.x != f[2 * c]) + (back.y != f[2 * c + 1]);
      }
    }
    reinterpret_cast<uint4*>(s_a + (size_t)j * (TC_SLOTS * 16))[slot] = make_uint4(u[0], u[1], u[2], u[3]);
    if (MODE == 2) reinterpret_cast<uint4*>(s_lo + (size_t)j * (TC_SLOTS * 16))[slot] = make_uint4(l[0], l[1], l[2], l[3]);
  };
  if (vec_ok) {
    // interior: columns cc = 1..128 (xx = x0 .. x0+127) as 32 groups of 4 pixels
    const int n_tasks = n_chunks * 3 * 32;
    for (int task = tid; task < n_tasks; task += TC_THREADS) {
      const int q = task & 31, rr = (task >> 5) % 3, j = task / 96;
      const int y = y0 - 1 + rr, xx = x0 + 4 * q;
      const bool ok = (y >= 0) && (y < H) && (xx < W);
      const float* p = src + (size_t)j * 8 * plane + (size_t)(ok ? y : 0) * W + (ok ? xx : 0);
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = ok ? __ldg(reinterpret_cast<const float4*>(p + (size_t)c * plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int slot = rr * TC_P + 1 + 4 * q;
      const float f0[8] = {v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x};
      const float f1[8] = {v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y};
      const float f2[8] = {v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z};
      const float f3[8] = {v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, v[6].w, v[7].w};
      put(f0, j, slot); put(f1, j, slot + 1); put(f2, j, slot + 2); put(f3, j, slot + 3);
    }
    // the two halo columns cc = 0 and cc = 129
    const int n_edge = n_chunks * 3 * 2;
    for (int task = tid; task < n_edge; task += TC_THREADS) {
      const int side = task & 1, rr = (task >> 1) % 3, j = task / 6;
      const int cc = side ? TC_P - 1 : 0;
      const int y = y0 - 1 + rr, xx = x0 - 1 + cc;
      const bool ok = (y >= 0) && (y < H) && (xx >= 0) && (xx < W);
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = ok ? __ldg(src + ((size_t)j * 8 + c) * plane + (size_t)y * W + xx) : 0.f;
      put(f, j, rr * TC_P + cc);
    }
  } else {
    const int n_tasks = n_chunks * 3 * TC_P;
    for (int task = tid; task < n_tasks; task += TC_THREADS) {
      const int cc = task % TC_P, rr = (task / TC_P) % 3, j = task / (3 * TC_P);
      const int y = y0 - 1 + rr, xx = x0 - 1 + cc;
      const bool ok = (y >= 0) && (y < H) && (xx >= 0) && (xx < W);
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = ok ? __ldg(src + ((size_t)j * 8 + c) * plane + (size_t)y * W + xx) : 0.f;
      put(f, j, rr * TC_P + cc);
    }
  }
}

}  // namespace snnflow

// tma.cuh -- TMA bulk-copy / mbarrier primitives (cp.async.bulk + mbarrier::complete_tx, SASS UBLKCP / SYNCS) and the
// fused complex multiply-add helpers shared by the streaming dense-block kernels (coarse_kernel.cu, schur_kernel.cu).
#pragma once
#include "common.cuh"
#include <cstdint>

namespace dda {

#ifndef DDA_HOST_EMU

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// acc += m * v  /  acc += conj(m) * v  on split real/imaginary accumulators: 4 FFMA each (written with explicit fma, the
// compiler may not re-associate `a += b*c - d*e` into two fused operations)
#ifdef DDA_NO_FUSE
__device__ __forceinline__ void cmac(float &ar, float &ai, float mx, float my, float vx, float vy) { ar += mx * vx - my * vy; ai += mx * vy + my * vx; }
__device__ __forceinline__ void cmacc(float &ar, float &ai, float mx, float my, float vx, float vy) { ar += mx * vx + my * vy; ai += mx * vy - my * vx; }
#else
__device__ __forceinline__ void cmac(float &ar, float &ai, float mx, float my, float vx, float vy) {
  ar = __fmaf_rn(-my, vy, __fmaf_rn(mx, vx, ar)); ai = __fmaf_rn(my, vx, __fmaf_rn(mx, vy, ai));
}
__device__ __forceinline__ void cmacc(float &ar, float &ai, float mx, float my, float vx, float vy) {
  ar = __fmaf_rn(my, vy, __fmaf_rn(mx, vx, ar)); ai = __fmaf_rn(-my, vx, __fmaf_rn(mx, vy, ai));
}
#endif

// ---- dense n x n block (column-major, in shared memory) times a vector, register tiling of the streaming kernels:
// thread (grp, p) owns the component pair 2p, 2p+1 and the chunk [grp*ch, grp*ch+ch) of the summation index.
// Two independent accumulator sets per output (even / odd summation index) halve the dependent FFMA chains: a thread
// has only ch/2 iterations per block, so the latency of the chain is what it waits for, not the issue rate.
// forward product: rows 2p, 2p+1 times the column chunk; results are ADDED to f0, f1
__device__ __forceinline__ void blk_forward(const cf *M, const cf *v, int n, int grp, int ch, int p, float &f0r, float &f0i, float &f1r, float &f1i) {
  const float4 *v4 = reinterpret_cast<const float4 *>(v + grp * ch);
  const cf *Mb = M + (size_t)(grp * ch) * n + 2 * p;
  float g0r = 0.f, g0i = 0.f, g1r = 0.f, g1i = 0.f;
#pragma unroll 2
  for (int cc = 0; cc < ch; cc += 2) {
    const float4 vv = v4[cc >> 1];
    const float4 m0 = *reinterpret_cast<const float4 *>(Mb + (size_t)cc * n);
    const float4 m1 = *reinterpret_cast<const float4 *>(Mb + (size_t)(cc + 1) * n);
    cmac(f0r, f0i, m0.x, m0.y, vv.x, vv.y);
    cmac(f1r, f1i, m0.z, m0.w, vv.x, vv.y);
    cmac(g0r, g0i, m1.x, m1.y, vv.z, vv.w);
    cmac(g1r, g1i, m1.z, m1.w, vv.z, vv.w);
  }
  f0r += g0r; f0i += g0i; f1r += g1r; f1i += g1i;
}
// daggered product: columns 2p, 2p+1 of the block, conjugated, times the row chunk of w; the walk over the row pairs is
// rotated by p (bank-conflict free stride-n accesses).  Results are ASSIGNED to a0, a1.
__device__ __forceinline__ void blk_dagger(const cf *M, const cf *w, int n, int grp, int ch, int p, float &a0r, float &a0i, float &a1r, float &a1i) {
  const int P = n >> 1;
  const cf *M0 = M + (size_t)(2 * p) * n, *M1 = M0 + n;
  int ip = grp * (ch >> 1) + p; if (ip >= P) ip -= P;
  float x0r = 0.f, x0i = 0.f, x1r = 0.f, x1i = 0.f, y0r = 0.f, y0i = 0.f, y1r = 0.f, y1i = 0.f;
#pragma unroll 2
  for (int i = 0; i < (ch >> 1); i++) {
    const float4 wv = *reinterpret_cast<const float4 *>(w + 2 * ip);
    const float4 m0 = *reinterpret_cast<const float4 *>(M0 + 2 * ip);
    const float4 m1 = *reinterpret_cast<const float4 *>(M1 + 2 * ip);
    cmacc(x0r, x0i, m0.x, m0.y, wv.x, wv.y);
    cmacc(y0r, y0i, m0.z, m0.w, wv.z, wv.w);
    cmacc(x1r, x1i, m1.x, m1.y, wv.x, wv.y);
    cmacc(y1r, y1i, m1.z, m1.w, wv.z, wv.w);
    ip++; if (ip == P) ip = 0;
  }
  a0r = x0r + y0r; a0i = x0i + y0i; a1r = x1r + y1r; a1i = x1i + y1i;
}

#endif

}  // namespace dda

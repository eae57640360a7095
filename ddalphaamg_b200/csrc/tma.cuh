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
__device__ __forceinline__ void cmac(float &ar, float &ai, float mx, float my, float vx, float vy) {
  ar = __fmaf_rn(-my, vy, __fmaf_rn(mx, vx, ar)); ai = __fmaf_rn(my, vx, __fmaf_rn(mx, vy, ai));
}
__device__ __forceinline__ void cmacc(float &ar, float &ai, float mx, float my, float vx, float vy) {
  ar = __fmaf_rn(my, vy, __fmaf_rn(mx, vx, ar)); ai = __fmaf_rn(-my, vx, __fmaf_rn(mx, vy, ai));
}

#endif

}  // namespace dda

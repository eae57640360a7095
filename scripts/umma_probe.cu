// umma_probe.cu -- development probe for the tcgen05.mma (kind::tf32) operand conventions used by mrhs_kernel.cu.
// One CTA, one MMA D[128 x 32] = A[128 x 8] * B[32 x 8]^T per variant; A and B hold small integers (exact in TF32), the result
// is compared with the host product.  Variants: A K-major / MN-major, and for each operand the two possible assignments
// of (leading, stride) byte offsets in the shared-memory descriptor.  Prints the maximum error of every variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/umma_probe scripts/umma_probe.cu && scripts/umma_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t instr_desc(int a_mn, int b_mn, int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// variant bits: 0: A MN-major; 1: swap A (lbo, sbo); 2: swap B (lbo, sbo)
__global__ void __launch_bounds__(128) probe(const float *Ag, const float *Bg, float *Dg, int variant) {
  __shared__ __align__(1024) float As[128 * 8];
  __shared__ __align__(1024) float Bs[32 * 8];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_mn = variant & 1;
  // A element (m, k): K-major canonical: (m%8)*4 + (m/8)*64 + (k%4) + (k/4)*32   [two K cores of a row group contiguous]
  //                   MN-major canonical: (m%4) + (m/4)*32 + k*4                  [one core = 8 k x 4 m]
  for (int q = tid; q < 128 * 8; q += 128) {
    const int m = q / 8, k = q % 8;
    const int o = a_mn ? ((m & 3) + (m >> 2) * 32 + k * 4) : ((m & 7) * 4 + (m >> 3) * 64 + (k & 3) + (k >> 2) * 32);
    As[o] = Ag[q];
  }
  // B element (n, k): K-major canonical like A K-major
  for (int q = tid; q < 32 * 8; q += 128) {
    const int n = q / 8, k = q % 8;
    Bs[(n & 7) * 4 + (n >> 3) * 64 + (k & 3) + (k >> 2) * 32] = Bg[q];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    // K-major: cores adjacent in K are 128 B apart (leading), row groups 256 B apart (stride)
    // MN-major: 4-element chunks along M are 128 B apart (stride), the next 8-k group would be 4096 B away (leading, unused at K = 8)
    uint32_t al = a_mn ? 4096 : 128, as_ = a_mn ? 128 : 256, bl = 128, bs_ = 256;
    if (variant & 2) { uint32_t t = al; al = as_; as_ = t; }
    if (variant & 4) { uint32_t t = bl; bl = bs_; bs_ = t; }
    const uint64_t da = smem_desc(smem_u32(As), al, as_), db = smem_desc(smem_u32(Bs), bl, bs_);
    const uint32_t id = instr_desc(a_mn, 0, 128, 32);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p; \n\t}\n"
        ::"r"(tmem), "l"(da), "l"(db), "r"(id), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  for (long it = 0; it < (1L << 24); it++) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    if (ok) break;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tmem + (((uint32_t)(warp * 32)) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 32; k++) Dg[tid * 32 + k] = __uint_as_float(r[k]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
  float hA[128 * 8], hB[32 * 8], hD[128 * 32], ref[128 * 32];
  for (int m = 0; m < 128; m++) for (int k = 0; k < 8; k++) hA[m * 8 + k] = (float)((m * 7 + k * 3) % 61 - 30);
  for (int n = 0; n < 32; n++) for (int k = 0; k < 8; k++) hB[n * 8 + k] = (float)((n * 5 + k * 11) % 23 - 11);
  for (int m = 0; m < 128; m++) for (int n = 0; n < 32; n++) { float s = 0; for (int k = 0; k < 8; k++) s += hA[m * 8 + k] * hB[n * 8 + k]; ref[m * 32 + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  for (int v = 0; v < 8; v++) {
    cudaMemset(dD, 0, sizeof(hD));
    probe<<<1, 128>>>(dA, dB, dD, v);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", v, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < 128 * 32; i++) err = fmax(err, fabs((double)hD[i] - ref[i]));
    printf("variant %d (A %s-major, A lbo/sbo %s, B lbo/sbo %s): max error %g   D[0][0..3] = %g %g %g %g (ref %g %g %g %g) D[5][1] = %g (ref %g)\n", v,
           (v & 1) ? "MN" : "K", (v & 2) ? "swapped" : "as assumed", (v & 4) ? "swapped" : "as assumed", err, hD[0], hD[1], hD[2], hD[3],
           ref[0], ref[1], ref[2], ref[3], hD[5 * 32 + 1], ref[5 * 32 + 1]);
  }
  return 0;
}

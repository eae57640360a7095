// coarse_kernel.cu -- hand-tuned coarse-operator apply for sm_100a:  eta_c = D_c phi_c  on all sites of a level.
//
// Reference counterparts: apply_coarse_operator_PRECISION (coarse_operator_generic.c:383-395) = coarse_self_couplings
// (:288-315) + coarse_hopping_term (coarse_oddeven_generic.c:447-581) with the dense kernels coarse_hopp /
// coarse_daggered_hopp (coarse_operator_generic.h:119-172).
//
// The operator is pure streaming of dense n x n complex blocks (n = 2 Nv = 40..64): per site the self coupling S(x)
// and the four forward hops F_mu(x); the backward hop of site x+mu is gamma5 F_mu(x)^H gamma5.  Like the reference,
// every F_mu(x) is read from HBM ONCE and used twice (scatter form): forward product for eta(x), daggered product for
// eta(x+mu).  The daggered results go to a small scratch field Z[x][mu][n] (4n complex per site, 1/n of the matrix
// traffic) that a second, trivial kernel adds at the destination sites.
//
// Persistent CTAs (128 threads); each walks over its sites and streams the 5 blocks of a site through a ring of shared
// memory stages filled by TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), one elected thread issuing, so the
// copy engine runs ahead of the arithmetic by STAGES-1 blocks.  Threads [0,n) own the rows of the forward product,
// threads [n,2n) the columns of the daggered product; the column walk is rotated by the column index so that the
// column-major block is read bank-conflict free in both directions.
// Algorithmic traffic per site: (5 n^2 + 6 n + 4 n) * 8 B  (SURVEY.md section 8d counts (4n^2 + n(n+1)/2 + 2n) * 8 B
// because the reference stores S packed Hermitian; bench.py reports against the SURVEY figure).
#include "coarse_op.h"
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

template <int STAGES>
__global__ void __launch_bounds__(128)
k_coarse_full(CoarseOp op, cf *__restrict__ out, const cf *__restrict__ in, cf *__restrict__ Z, int nsites) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, nh = n / 2;
  cf *Ms = reinterpret_cast<cf *>(smem_raw);                    // [STAGES][n*n]
  cf *vec = Ms + (size_t)STAGES * nn;                           // [6][n]: v(x), v(x+mu) x4, gamma5 v(x)
  uint64_t *full = reinterpret_cast<uint64_t *>(vec + 6 * n);   // [STAGES]
  const int tid = threadIdx.x;
  const uint32_t bytes = (uint32_t)(nn * sizeof(cf));
  const int my_sites = (nsites - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = 5 * my_sites;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // job j = 5 * k + m of this CTA: site x = blockIdx.x + k * gridDim.x, block m (0: S, 1..4: F_{m-1})
  auto issue = [&](int j) {
    const int k = j / 5, m = j - 5 * k;
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
    const cf *src = (m == 0) ? op.S + x * nn : op.F + (x * 4 + (m - 1)) * nn;
    const int st = j % STAGES;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the stage precede the async write
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(Ms + (size_t)st * nn, src, bytes, &full[st]);
  };
  if (tid == 0) for (int j = 0; j < STAGES && j < total; j++) issue(j);

  for (int k = 0; k < my_sites; k++) {
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
    if (tid < n) {
      const cf v = in[x * n + tid];
      vec[tid] = v;
      vec[5 * n + tid] = (tid < nh) ? v : -v;
#pragma unroll
      for (int mu = 0; mu < 4; mu++) vec[(1 + mu) * n + tid] = in[(long)op.nb[(long)mu * op.V + x] * n + tid];
    }
    __syncthreads();
    cf acc(0.f, 0.f);
    for (int m = 0; m < 5; m++) {
      const int j = 5 * k + m, st = j % STAGES;
      mbar_wait(&full[st], (uint32_t)((j / STAGES) & 1));
      const cf *M = Ms + (size_t)st * nn;
      if (tid < n) {                      // forward: row tid of M times v
        const cf *v = vec + m * n;
        cf a0(0.f, 0.f), a1(0.f, 0.f);
        int c = 0;
        for (; c + 1 < n; c += 2) { fma_(a0, M[c * n + tid], v[c]); fma_(a1, M[(c + 1) * n + tid], v[c + 1]); }
        if (c < n) fma_(a0, M[c * n + tid], v[c]);
        acc += a0 + a1;
      } else if (tid < 2 * n && m > 0) {  // daggered: column (tid - n) of M, conjugated, times gamma5 v(x)
        const int col = tid - n;
        const cf *Mc = M + col * n, *w = vec + 5 * n;
        cf a0(0.f, 0.f), a1(0.f, 0.f);
        int rr = col;                     // rotated start: bank-conflict free column walk
        int r = 0;
        for (; r + 1 < n; r += 2) {
          fmac_(a0, Mc[rr], w[rr]); rr++; if (rr == n) rr = 0;
          fmac_(a1, Mc[rr], w[rr]); rr++; if (rr == n) rr = 0;
        }
        if (r < n) fmac_(a0, Mc[rr], w[rr]);
        cf z = a0 + a1;
        if (col >= nh) z = -z;
        Z[(x * 4 + (m - 1)) * n + col] = z;
      }
      __syncthreads();
      if (tid == 0 && j + STAGES < total) issue(j + STAGES);
    }
    if (tid < n) out[x * n + tid] = acc;
  }
}

// eta(x) += sum_mu Z[x-mu][mu]; backward hops whose source site is a ghost (other rank) are computed directly from the
// ghost copies of F and phi
__global__ void k_coarse_combine(CoarseOp op, cf *__restrict__ out, const cf *__restrict__ in, const cf *__restrict__ Z, long total) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = op.n, nh = n / 2;
  const long x = i / n; const int r = (int)(i - x * n);
  const long nn = (long)n * n;
  cf acc = out[i];
#pragma unroll
  for (int mu = 0; mu < 4; mu++) {
    const long nbr = op.nb[(long)(4 + mu) * op.V + x];
    if (nbr < op.V) acc += Z[(nbr * 4 + mu) * n + r];
    else {
      const cf *M = op.F + (nbr * 4 + mu) * nn + (long)r * n, *v = in + nbr * n;
      cf a1(0.f, 0.f), a2(0.f, 0.f);
      for (int c = 0; c < nh; c++) fmac_(a1, M[c], v[c]);
      for (int c = nh; c < n; c++) fmac_(a2, M[c], v[c]);
      acc += (r < nh) ? (a1 - a2) : (a2 - a1);
    }
  }
  out[i] = acc;
}

bool coarse_apply_fast(const CoarseOp &op, cf *out, const cf *in, cf *Z) {
  const int n = op.n;
  if (n > 64 || (n & 1) || !Z || op.V <= 0) return false;
  const size_t nn = (size_t)n * n;
  int dev = 0; cudaGetDevice(&dev);
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int stages = n <= 48 ? 4 : 3;
  const size_t smem = stages * nn * sizeof(cf) + 6 * n * sizeof(cf) + 8 * sizeof(uint64_t);
  static size_t attr4 = 0, attr3 = 0;
  if (stages == 4 && smem > attr4) { CUDA_CHECK(cudaFuncSetAttribute(k_coarse_full<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr4 = smem; }
  if (stages == 3 && smem > attr3) { CUDA_CHECK(cudaFuncSetAttribute(k_coarse_full<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr3 = smem; }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) return false;
  if (per_sm > 8) per_sm = 8;
  long grid = std::min<long>(op.V, (long)sms * per_sm);
  if (stages == 4) k_coarse_full<4><<<(unsigned)grid, 128, smem, g_stream>>>(op, out, in, Z, (int)op.V);
  else k_coarse_full<3><<<(unsigned)grid, 128, smem, g_stream>>>(op, out, in, Z, (int)op.V);
  g_launch_count++;
  const long total = op.V * n;
  k_coarse_combine<<<(unsigned)((total + 127) / 128), 128, 0, g_stream>>>(op, out, in, Z, total);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
}

#endif

}  // namespace dda

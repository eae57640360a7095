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

// Register tiling: thread t of the 128 owns the component PAIR p = t % (n/2) of group g = t / (n/2)  (G groups, G | n/2
// chosen by the host, threads beyond G*n/2 idle).  Forward product: rows 2p, 2p+1 times the group's chunk of n/G
// columns; daggered product: columns 2p, 2p+1 times the group's chunk of rows.  Both walk the column-major block with
// 16-byte shared-memory loads (two complex per load); the daggered walk is rotated by p row pairs, which makes the
// stride-n column accesses bank-conflict free.  ~4.75 instructions per complex multiply-add (4 are the FFMAs).
// Partial sums stay in registers over the 5 blocks of a site and are combined once per site.
template <int STAGES>
__global__ void __launch_bounds__(128)
k_coarse_full(CoarseOp op, cf *__restrict__ out, const cf *__restrict__ in, cf *__restrict__ Z, int nsites, int G) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, nh = n / 2, P = n / 2;
  const int ch = n / G;                                         // even chunk of columns (forward) / rows (daggered)
  cf *Ms = reinterpret_cast<cf *>(smem_raw);                    // [STAGES][n*n]
  cf *vec = Ms + (size_t)STAGES * nn;                           // [6][n]: v(x), v(x+mu) x4, gamma5 v(x)
  cf *part = vec + 6 * n;                                       // [5][G][n] partial sums (forward, 4 x daggered)
  uint64_t *full = reinterpret_cast<uint64_t *>(part + 5 * G * n);   // [STAGES]
  const int tid = threadIdx.x;
  const int grp = tid / P, p = tid - grp * P;
  const bool active = grp < G;
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

  // the input vectors of a site (5 n complex: phi(x), phi(x+mu)) are fetched one site ahead into registers, so that the
  // dependent index -> vector loads are in flight while the previous site's blocks are processed
  cf pre[3];
  auto prefetch = [&](int k) {
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int q = tid + 128 * i;
      if (q < 5 * n) {
        const int vsel = q / n, c = q - vsel * n;
        const long src = (vsel == 0) ? x : (long)op.nb[(long)(vsel - 1) * op.V + x];
        pre[i] = in[src * n + c];
      }
    }
  };
  if (my_sites > 0) prefetch(0);

  for (int k = 0; k < my_sites; k++) {
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int q = tid + 128 * i;
      if (q < 5 * n) {
        vec[q] = pre[i];
        if (q < n) vec[5 * n + q] = (q < nh) ? pre[i] : -pre[i];
      }
    }
    __syncthreads();
    if (k + 1 < my_sites) prefetch(k + 1);
    float f0r = 0.f, f0i = 0.f, f1r = 0.f, f1i = 0.f;             // forward rows 2p, 2p+1
    float zr[4][2], zi[4][2];                                     // daggered columns 2p, 2p+1 per direction
#pragma unroll
    for (int mu = 0; mu < 4; mu++) { zr[mu][0] = zr[mu][1] = zi[mu][0] = zi[mu][1] = 0.f; }
#pragma unroll
    for (int m = 0; m < 5; m++) {
      const int j = 5 * k + m, st = j % STAGES;
      mbar_wait(&full[st], (uint32_t)((j / STAGES) & 1));
      const cf *M = Ms + (size_t)st * nn;
      if (active) {
        {                                 // forward
          const float4 *v4 = reinterpret_cast<const float4 *>(vec + m * n + grp * ch);
          const cf *Mb = M + (size_t)(grp * ch) * n + 2 * p;
#pragma unroll 2
          for (int cc = 0; cc < ch; cc += 2) {
            const float4 v = v4[cc >> 1];
            const float4 m0 = *reinterpret_cast<const float4 *>(Mb + (size_t)cc * n);
            const float4 m1 = *reinterpret_cast<const float4 *>(Mb + (size_t)(cc + 1) * n);
            f0r += m0.x * v.x - m0.y * v.y; f0i += m0.x * v.y + m0.y * v.x;
            f1r += m0.z * v.x - m0.w * v.y; f1i += m0.z * v.y + m0.w * v.x;
            f0r += m1.x * v.z - m1.y * v.w; f0i += m1.x * v.w + m1.y * v.z;
            f1r += m1.z * v.z - m1.w * v.w; f1i += m1.z * v.w + m1.w * v.z;
          }
        }
        if (m > 0) {                      // daggered: conj(M[r][c]) w[r]
          const cf *w = vec + 5 * n;
          const cf *M0 = M + (size_t)(2 * p) * n, *M1 = M0 + n;
          float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
          int ip = grp * (ch >> 1) + p; if (ip >= P) ip -= P;    // rotated row pair
#pragma unroll 2
          for (int i = 0; i < (ch >> 1); i++) {
            const float4 wv = *reinterpret_cast<const float4 *>(w + 2 * ip);
            const float4 m0 = *reinterpret_cast<const float4 *>(M0 + 2 * ip);
            const float4 m1 = *reinterpret_cast<const float4 *>(M1 + 2 * ip);
            a0r += m0.x * wv.x + m0.y * wv.y; a0i += m0.x * wv.y - m0.y * wv.x;
            a0r += m0.z * wv.z + m0.w * wv.w; a0i += m0.z * wv.w - m0.w * wv.z;
            a1r += m1.x * wv.x + m1.y * wv.y; a1i += m1.x * wv.y - m1.y * wv.x;
            a1r += m1.z * wv.z + m1.w * wv.w; a1i += m1.z * wv.w - m1.w * wv.z;
            ip++; if (ip == P) ip = 0;
          }
          zr[m - 1][0] = a0r; zi[m - 1][0] = a0i; zr[m - 1][1] = a1r; zi[m - 1][1] = a1i;
        }
      }
      __syncthreads();
      if (tid == 0 && j + STAGES < total) issue(j + STAGES);
    }
    // combine the G partial sums
    if (active) {
      float4 *pf = reinterpret_cast<float4 *>(part + grp * n + 2 * p);
      *pf = make_float4(f0r, f0i, f1r, f1i);
#pragma unroll
      for (int mu = 0; mu < 4; mu++)
        *reinterpret_cast<float4 *>(part + ((1 + mu) * G + grp) * n + 2 * p) = make_float4(zr[mu][0], zi[mu][0], zr[mu][1], zi[mu][1]);
    }
    __syncthreads();
    for (int q = tid; q < 5 * n; q += 128) {
      const int sel = q / n, c = q - sel * n;
      cf a = part[(sel * G) * n + c];
      for (int g2 = 1; g2 < G; g2++) a += part[(sel * G + g2) * n + c];
      if (sel == 0) out[x * n + c] = a;
      else Z[(x * 4 + (sel - 1)) * n + c] = (c < nh) ? a : -a;
    }
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

int g_coarse_stages = 0;   // 0: default; 3 / 4: force the depth of the TMA ring (tuning knob, env DDA_COARSE_STAGES)

bool coarse_apply_fast(const CoarseOp &op, cf *out, const cf *in, cf *Z) {
  static bool env_read = false;
  if (!env_read) { const char *e = getenv("DDA_COARSE_STAGES"); if (e) g_coarse_stages = atoi(e); env_read = true; }
  const int n = op.n;
  if (n > 64 || n < 8 || (n & 3) || !Z || op.V <= 0) return false;
  const size_t nn = (size_t)n * n;
  int dev = 0; cudaGetDevice(&dev);
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int stages = (g_coarse_stages >= 2 && g_coarse_stages <= 4) ? g_coarse_stages : 2;   // measured on B200: 2 stages x 7 CTAs/SM beats deeper rings
  int G = 128 / (n / 2);
  while (G > 1 && n % (2 * G) != 0) G--;          // even chunk n / G
  if (n % (2 * G) != 0) return false;
  const size_t smem = stages * nn * sizeof(cf) + (6 + 5 * G) * n * sizeof(cf) + 8 * sizeof(uint64_t);
  static size_t attr4 = 0, attr3 = 0, attr2 = 0;
  if (stages == 2 && smem > attr2) { CUDA_CHECK(cudaFuncSetAttribute(k_coarse_full<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr2 = smem; }
  if (stages == 4 && smem > attr4) { CUDA_CHECK(cudaFuncSetAttribute(k_coarse_full<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr4 = smem; }
  if (stages == 3 && smem > attr3) { CUDA_CHECK(cudaFuncSetAttribute(k_coarse_full<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr3 = smem; }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) return false;
  if (per_sm > 12) per_sm = 12;
  long grid = std::min<long>(op.V, (long)sms * per_sm);
  if (stages == 2) k_coarse_full<2><<<(unsigned)grid, 128, smem, g_stream>>>(op, out, in, Z, (int)op.V, G);
  else if (stages == 4) k_coarse_full<4><<<(unsigned)grid, 128, smem, g_stream>>>(op, out, in, Z, (int)op.V, G);
  else k_coarse_full<3><<<(unsigned)grid, 128, smem, g_stream>>>(op, out, in, Z, (int)op.V, G);
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

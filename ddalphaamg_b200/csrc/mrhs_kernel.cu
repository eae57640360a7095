// mrhs_kernel.cu -- coarse operator applied to 12 right-hand sides at once on the 5th-generation tensor cores (sm_100a):
//   Y_j = D_c V_j ,  j = 0..11      (BASELINE.json configs[4] "12-RHS batched coarse solves")
//
// With one right-hand side the coarse operator is a stream of dense n x n complex blocks used for ONE matrix-vector
// product each (coarse_kernel.cu, HBM bound at 0.7 of the copy rate).  With 12 right-hand sides every block is used for a
// n x n by n x 12 product: the bytes per right-hand side drop by 12 and the work is a dense contraction, which is what
// tcgen05.mma is for.  No reference counterpart (the reference has no multi-RHS path); per column the result is the
// single-RHS operator apply_coarse_operator_PRECISION (coarse_operator_generic.c:383-395), which is what the parity test
// compares with.
//
// Real arithmetic on the tensor core, complex blocks as they lie in memory.  A block is column-major complex, i.e. a
// REAL matrix W[rho][c] with rho = 2 r + (re|im), 2n x n, "rho-major".  With B = [Re V | Im V] (n x 24)
//     P = W B :  P[2r][j] = sum Mr Vr, P[2r+1][j] = sum Mi Vr, P[2r][12+j] = sum Mr Vi, P[2r+1][12+j] = sum Mi Vi
//     Re Y[r][j] = P[2r][j] - P[2r+1][12+j] ,  Im Y[r][j] = P[2r+1][j] + P[2r][12+j]                     (forward product)
// and with B' = [w | -i w] laid out along rho (w = G5 V(x))
//     Q = W^T B' :  Q[c][j] = sum_r Mr wr + Mi wi = Re (M^H w)_c ,  Q[c][12+j] = sum_r Mr wi - Mi wr = Im (M^H w)_c  (daggered)
// Both products use K-major operands in the canonical no-swizzle layout of tcgen05.mma (128-byte core matrices of 8 rows x
// 4 K-elements).  For the daggered product (M = column c, K = rho) that is the block as it lies in memory, 16-byte chunks
// of 4 consecutive rho grouped by 8 columns; for the forward product (M = rho, K = c) the block is transposed while it is
// re-tiled.
// fp32 accuracy on TF32 hardware: every operand is split x = hi + lo (hi = 11 significant bits) and each product is three
// MMAs hi*hi + lo*hi + hi*lo accumulated in the same fp32 TMEM accumulator (error ~2^-21 relative).
//
// Per CTA (256 threads, persistent over sites): TMA bulk copies stream the raw blocks into a ring (cp.async.bulk + mbarrier,
// as in coarse_kernel.cu); all threads re-tile the block (with the hi/lo split) and the right-hand sides into operand
// buffers; ONE thread issues the tcgen05.mma sequence and commits it to an mbarrier; accumulators live in TMEM (forward:
// 128 lanes x 32 columns, daggered: 4 x 32 columns) and are read back with tcgen05.ld for the epilogue.  The daggered
// results go to the scratch field Z and are added at the target sites by k_coarse_combine (scatter form: every hop matrix
// is read from HBM once per 12 right-hand sides).
#include "coarse_op.h"
#include "tma.cuh"

namespace dda {

#ifndef DDA_HOST_EMU

namespace mrhs {

const int NR = 12;                       // right-hand sides
const int NB = 32;                       // N of the MMA (24 used: [Re | Im], padded to a multiple of 16)
const int TMEM_COLS = 512;               // 5 accumulator sets x 96 columns (see ACC below), power of two
const int ACC = 96;                      // columns per accumulator set: [hi*hi | hi*lo] (64, one MMA with N = 64) + [lo*hi] (32)

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  // mbarrier wait with a bound: a protocol error traps (CUDA error on the host) instead of hanging the GPU
  for (long it = 0; it < (1L << 26); it++) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets
// in 16-byte units, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, majors, N >> 3, M >> 4
__device__ __forceinline__ uint32_t instr_desc(int a_mn_major, int b_mn_major, int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p; \n\t"
      "}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (lane = row, register k = column k)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
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
#pragma unroll
  for (int k = 0; k < 32; k++) v[k] = __uint_as_float(r[k]);
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

// Operand buffers (floats), all in the canonical no-swizzle K-major layout of tcgen05.mma (verified on the B200 with
// scripts/umma_probe.cu: 128-byte core matrices = 8 rows x 4 K-elements, leading byte offset = distance of the cores
// adjacent in K, stride byte offset = distance of the 8-row groups):
//   Af: forward view,  M = rho (16 row groups), K = column c (n/4 cores per row group)      -- the block TRANSPOSED
//   Ad: daggered view, M = column c (8 row groups of 32 cores), K = rho                     -- the block as it lies in memory
//   Bf: [Re V | Im V], N = 32 right-hand-side slots (4 row groups), K = c;   Bd: [w | -i w], K = rho
const int AD_FLOATS = 8 * 32 * 32;

const int NT = 256;                      // threads per CTA (8 warps: re-tiling and operand fills are what the CTA spends its time on)

template <int STAGES>
__global__ void __launch_bounds__(NT)
k_coarse_mrhs(CoarseOp op, cf *__restrict__ out, const cf *__restrict__ in, cf *__restrict__ Z, long vstride, long zstride, int nsites) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, nh = n / 2, n2 = 2 * n;
  const int kcf = n / 4, kcd = n2 / 4;                              // K cores of the forward / daggered operands
  const int AF_FLOATS = 16 * kcf * 32;
  float *Afh = reinterpret_cast<float *>(smem_raw);
  float *Afl = Afh + AF_FLOATS;
  float *Adh = Afl + AF_FLOATS;
  float *Adl = Adh + AD_FLOATS;
  float *Bfh = Adl + AD_FLOATS;                                     // [4][kcf][32]
  float *Bfl = Bfh + 4 * kcf * 32;
  float *Bdh = Bfl + 4 * kcf * 32;                                  // [4][kcd][32]
  float *Bdl = Bdh + 4 * kcd * 32;
  cf *raw = reinterpret_cast<cf *>(Bdl + 4 * kcd * 32);             // [STAGES][nn]
  uint64_t *full = reinterpret_cast<uint64_t *>(raw + (size_t)STAGES * nn);   // [STAGES]
  uint64_t *mma_done = full + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bytes = (uint32_t)(nn * sizeof(cf));
  const int my_sites = (nsites - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = 5 * my_sites;

  for (int q = tid; q < 2 * AF_FLOATS + 2 * AD_FLOATS + 8 * (kcf + kcd) * 32; q += NT) Afh[q] = 0.f;   // padding stays zero
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    mbar_init(mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  auto issue = [&](int j) {
    const int k = j / 5, m = j - 5 * k;
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
    const cf *src = (m == 0) ? op.S + x * nn : op.F + (x * 4 + (m - 1)) * nn;
    const int st = j % STAGES;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(raw + (size_t)st * nn, src, bytes, &full[st]);
  };
  if (tid == 0) for (int j = 0; j < STAGES && j < total; j++) issue(j);

  // A and B K-major.  TF32 x 3 with two MMAs per K step into INDEPENDENT accumulators: A_hi x [B_hi | B_lo] (N = 64: the hi and
  // lo operand buffers of B are adjacent row groups) and A_lo x B_hi (N = 32), summed in the epilogue.  A chain of dependent
  // MMAs on one accumulator costs its full latency per link (the first version ran 63 of them back to back per block).
  const uint32_t idesc64 = instr_desc(0, 0, 128, 2 * NB), idesc32 = instr_desc(0, 0, 128, NB);
  const uint32_t aAfh = smem_u32(Afh), aAfl = smem_u32(Afl), aAdh = smem_u32(Adh), aAdl = smem_u32(Adl);
  const uint32_t aBfh = smem_u32(Bfh), aBfl = smem_u32(Bfl), aBdh = smem_u32(Bdh), aBdl = smem_u32(Bdl);
  uint32_t mma_phase = 0;

  // the right-hand sides a block multiplies (12 x n complex) are fetched one block ahead into registers
  const int NPRE = 3;                                               // 12 * 64 / NT
  cf pre[NPRE];
  auto prefetch = [&](int j) {                                      // vectors of block j: site x (S) or x + mu (F_mu)
    const int k = j / 5, m = j - 5 * k;
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
    const long src = (m == 0) ? x : (long)op.nb[(long)(m - 1) * op.V + x];
#pragma unroll
    for (int i = 0; i < NPRE; i++) {
      const int q = tid + NT * i;
      if (q < NR * n) { const int jr = q / n, c = q - jr * n; pre[i] = in[(long)jr * vstride + src * n + c]; }
    }
  };
  if (total > 0) prefetch(0);

  for (int k = 0; k < my_sites; k++) {
    const long x = (long)blockIdx.x + (long)k * gridDim.x;
    for (int m = 0; m < 5; m++) {
      const int j = 5 * k + m, st = j % STAGES;
      // forward B = [Re V | Im V] from the prefetched registers; for m = 0 these are the site's own vectors, from which
      // the daggered operand B' = [w | -i w], w = G5 V(x), is filled as well (rows j and 12 + j, K = rho = 2 r + re|im)
#pragma unroll
      for (int i = 0; i < NPRE; i++) {
        const int q = tid + NT * i;
        if (q < NR * n) {
          const int jr = q / n, c = q - jr * n;
          const cf v = pre[i];
          const int row1 = NR + jr;
          const int o0 = ((jr >> 3) * kcf + (c >> 2)) * 32 + (jr & 7) * 4 + (c & 3);
          const int o1 = ((row1 >> 3) * kcf + (c >> 2)) * 32 + (row1 & 7) * 4 + (c & 3);
          float hi, lo;
          split_tf32(v.re, hi, lo); Bfh[o0] = hi; Bfl[o0] = lo;
          split_tf32(v.im, hi, lo); Bfh[o1] = hi; Bfl[o1] = lo;
          if (m == 0) {
            const cf w = (c >= nh) ? -v : v;
            const float val[2][2] = {{w.re, w.im}, {w.im, -w.re}};   // [row block][re | im position]
#pragma unroll
            for (int blk = 0; blk < 2; blk++) {
              const int row = blk * NR + jr;
#pragma unroll
              for (int ri = 0; ri < 2; ri++) {
                const int kk = 2 * c + ri;
                const int o = ((row >> 3) * kcd + (kk >> 2)) * 32 + (row & 7) * 4 + (kk & 3);
                split_tf32(val[blk][ri], hi, lo);
                Bdh[o] = hi; Bdl[o] = lo;
              }
            }
          }
        }
      }
      if (j + 1 < total) prefetch(j + 1);
      mbar_wait_bounded(&full[st], (uint32_t)((j / STAGES) & 1));
      // re-tile the raw block (column-major complex = real W[rho][c], rho fastest) in two passes, each followed by its MMAs, so
      // that the tensor core works on the daggered product while the threads build the forward operand:
      //   pass 1, daggered view: 16-byte chunk (column c, rho = 4g..4g+3) -> core matrix (c / 8, g), row c % 8   (one 16-byte store)
      //   pass 2, forward view:  element (rho, c) -> core (rho / 8, c / 4), row rho % 8, position c % 4          (four scalar stores)
      // lanes: c % 4 = lane % 4, g % 2 = (lane / 4) % 2, and in store step t lane group r = lane / 8 writes rho = 4 g + (t + r) % 4:
      // the 32 lanes of a scalar store hit 32 distinct banks (the first version of this kernel lost 71 % of its shared-memory
      // wavefronts to conflicts here, profiles/r2_ncu_full_k_coarse_mrhs_a.txt)
      const float4 *R4 = reinterpret_cast<const float4 *>(raw + (size_t)st * nn);
      const int ncq = n >> 2, ngq = kcd >> 1, combos = ncq * ngq;
      if (m > 0) {
        for (int u = warp * 4 + (lane >> 3); u < combos; u += 4 * (NT / 32)) {
          const int cq = u / ngq, gq = u - cq * ngq;
          const int c = 4 * cq + (lane & 3), g = 2 * gq + ((lane >> 2) & 1);
          const float4 v = R4[c * kcd + g];
          float4 h, l;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          const int od = ((c >> 3) * 32 + g) * 32 + (c & 7) * 4;
          *reinterpret_cast<float4 *>(Adh + od) = h;
          *reinterpret_cast<float4 *>(Adl + od) = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand buffers written by the generic proxy
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0 && m > 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // daggered: D_mu[c][slot] = sum_rho W[rho][c] B'[slot][rho]; K = 2n, 8 per MMA = two cores (256 B) of A and of B
        const uint32_t td = tmem + (uint32_t)(ACC * m);
        for (int ks = 0; ks < n2 / 8; ks++) {
          const uint32_t o = (uint32_t)ks * 256;
          const uint64_t ah = smem_desc(aAdh + o, 128, 4096), al = smem_desc(aAdl + o, 128, 4096);
          const uint64_t bhl = smem_desc(aBdh + o, 128, (uint32_t)kcd * 128);          // rows 0..31 = hi, 32..63 = lo
          mma_tf32(td, ah, bhl, idesc64, ks > 0 ? 1u : 0u);
          mma_tf32(td + 64u, al, bhl, idesc32, ks > 0 ? 1u : 0u);
        }
      }
      {
        const int r = lane >> 3;
        for (int u = warp * 4 + r; u < combos; u += 4 * (NT / 32)) {
          const int cq = u / ngq, gq = u - cq * ngq;
          const int c = 4 * cq + (lane & 3), g = 2 * gq + ((lane >> 2) & 1);
          const float4 v = R4[c * kcd + g];
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const int tt = (t + r) & 3;
            const float x = tt == 0 ? v.x : (tt == 1 ? v.y : (tt == 2 ? v.z : v.w));
            float hi, lo; split_tf32(x, hi, lo);
            const int rho = 4 * g + tt;
            const int of = ((rho >> 3) * kcf + (c >> 2)) * 32 + (rho & 7) * 4 + (c & 3);
            Afh[of] = hi; Afl[of] = lo;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // forward: D_fwd[rho][slot] += sum_c W[rho][c] B[slot][c]; K = n
        for (int ks = 0; ks < n / 8; ks++) {
          const uint32_t o = (uint32_t)ks * 256;
          const uint64_t ah = smem_desc(aAfh + o, 128, (uint32_t)kcf * 128), al = smem_desc(aAfl + o, 128, (uint32_t)kcf * 128);
          const uint64_t bhl = smem_desc(aBfh + o, 128, (uint32_t)kcf * 128);
          mma_tf32(tmem, ah, bhl, idesc64, (m > 0 || ks > 0) ? 1u : 0u);
          mma_tf32(tmem + 64u, al, bhl, idesc32, (m > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(mma_done);                                        // arrives when every MMA issued so far has completed
      }
      mbar_wait_bounded(mma_done, mma_phase & 1);                    // operand buffers and the raw stage are free again
      mma_phase++;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0 && j + STAGES < total) issue(j + STAGES);
    }
    // epilogue: accumulators -> registers -> global.  A warp reaches the TMEM lanes 32 (warp % 4) .. +31 = rows of the
    // accumulators; warps 0-3 take the forward accumulator and the daggered ones of mu = 0, 1, warps 4-7 those of mu = 2, 3.
    {
      const int row = (warp & 3) * 32 + lane;
      const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
      float v[32], v2[32];
      auto load_sum = [&](uint32_t base) {        // [hi*hi] + [hi*lo] + [lo*hi]
        tmem_ld32(base, v);
        tmem_ld32(base + 32u, v2);
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] += v2[k];
        tmem_ld32(base + 64u, v2);
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] += v2[k];
      };
      if (warp < 4) {
        load_sum(tmem + lane_base);
        // forward rows rho = 2 r + (re|im): Re Y = P[2r][j] - P[2r+1][12+j], Im Y = P[2r+1][j] + P[2r][12+j]
        const int rho = row, r = rho >> 1, im = rho & 1;
#pragma unroll
        for (int j = 0; j < NR; j++) {
          const float other = __shfl_xor_sync(0xffffffffu, v[NR + j], 1);     // partner row's [12 + j] entry
          const float val = im ? v[j] + other : v[j] - other;
          if (rho < n2) reinterpret_cast<float *>(out + (long)j * vstride + x * n + r)[im] = val;
        }
      }
#pragma unroll 1
      for (int mu = (warp < 4 ? 0 : 2); mu < (warp < 4 ? 2 : 4); mu++) {
        load_sum(tmem + lane_base + (uint32_t)(ACC * (1 + mu)));
        const int c = row;
        if (c < n) {
          const float sg = (c < nh) ? 1.f : -1.f;
#pragma unroll
          for (int j = 0; j < NR; j++) Z[(long)j * zstride + (x * 4 + mu) * n + c] = cf(sg * v[j], sg * v[NR + j]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                 // accumulators are free for the next site
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS));
}

}  // namespace mrhs

void coarse_combine(const CoarseOp &op, cf *out, const cf *in, const cf *Z);   // coarse_kernel.cu: eta(x) += sum_mu Z[x-mu][mu]

// out_j = D_c in_j for j < 12: vectors j at in + j * vstride / out + j * vstride (vstride >= (V + ghost sites) * n), Z: scratch
// of 12 x zstride complex (zstride >= 4 n V).  Halo slabs of the inputs must be current.  Returns false when the shape is
// not supported.
bool coarse_apply_mrhs(const CoarseOp &op, cf *out, const cf *in, cf *Z, long vstride, long zstride) {
  const int n = op.n;
  if (n > 64 || n < 8 || (n & 7) || op.V <= 0) return false;
  const size_t nn = (size_t)n * n;
  const int STAGES = 2;
  const size_t smem = (2 * (size_t)(16 * (n / 4) * 32) + 2 * (size_t)mrhs::AD_FLOATS + 8 * (size_t)(n / 4 + n / 2) * 32) * sizeof(float) +
                      STAGES * nn * sizeof(cf) + (STAGES + 1) * sizeof(uint64_t) + 16;
  static size_t attr = 0;
  if (smem > attr) { CUDA_CHECK(cudaFuncSetAttribute(mrhs::k_coarse_mrhs<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int per_sm = (int)std::min<size_t>(2, (227 * 1024) / (smem + 1024));    // TMEM: 2 x 256 columns per SM
  if (per_sm < 1) return false;
  const long grid = std::min<long>(op.V, (long)sms * per_sm);
  mrhs::k_coarse_mrhs<2><<<(unsigned)grid, mrhs::NT, smem, g_stream>>>(op, out, in, Z, vstride, zstride, (int)op.V);
  g_launch_count++;
  for (int j = 0; j < mrhs::NR; j++) coarse_combine(op, out + (long)j * vstride, in + (long)j * vstride, Z + (long)j * zstride);
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
}

#endif

}  // namespace dda

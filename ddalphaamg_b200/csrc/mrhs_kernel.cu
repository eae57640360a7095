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
// Operator images.  The operand layouts are fixed by the hardware, the operator is constant over a solve: coarse_mrhs_tile
// writes, once per operator, every block in the two operand layouts (forward view for S and F_mu, daggered view for F_mu;
// 9 images of n x n complex per site, 72 n^2 bytes) so that ONE TMA bulk copy per block lands an MMA-ready operand in shared
// memory.  (The first versions re-tiled the column-major blocks inside the kernel: 45 instructions per matrix element,
// 1.5 ms per application at 32^3 x 64 against 0.10 ms for one single-RHS application, profiles/r2_ncu_full_k_coarse_mrhs_c.txt.)
//
// Per CTA (448 threads, one CTA per SM, persistent over sites), warp-specialised, mbarriers only inside the site loop:
//   warp 12 (one thread)  producer: cp.async.bulk of block j's image(s) into landing buffer j % 2          -> full[stage]
//   warps 0-7             workers (two groups of four warps, group g owns the blocks j = g mod 2): B operands of block j
//                                   (right-hand sides, hi/lo split), then the TF32 split of the landed image into operand set
//                                   j % 2 (hi = truncated value, lo = remainder)               -> ready[stage], rawfree[stage]
//   warp 13 (one thread)  issuer:   tcgen05.mma sequence of block j, tcgen05.commit                      -> empty[stage], acc_full
//   warps 8-11            epilogue: tcgen05.ld of the site's five accumulator sets, stores                -> acc_empty
// (for n > 40 the landing buffers do not fit next to two operand sets: the copies land in the operand set and are split in
// place; for n >= 56 there is one operand set.)  Accumulators live in TMEM (forward: 128 lanes x 96 columns, daggered: 4 x 96
// columns).  The daggered results go to the scratch field Z and are added at the target sites by k_coarse_combine_batch
// (scatter form: every hop matrix is read from HBM once per 12 right-hand sides).
//
// Measured (B200, 32^3 x 64 level 1: 8 192 sites, n = 40): 0.47 ms per 12-RHS application against 12 x 0.10 ms single-RHS
// applications (2.6 x); profiles/r2_ncu_full_k_coarse_mrhs_d.txt: tensor pipe 24 % active, DRAM 32 %, shared-memory pipe ~80 %
// busy (per block the tensor core reads 165 KB of operands, M = 128 rows per MMA although only 2n = 80 / n = 40 are occupied,
// and the split adds 100 KB, against 25.6 KB of HBM traffic).  An experiment with the operand roles swapped (right-hand sides
// as the M = 64 operand: 110 KB of operand reads) ran at the same speed (profiles/RESULTS_r2.md), so the bound is the
// per-block hand-off chain and the 30 MMAs issued by one thread, not shared-memory bandwidth.
#include "coarse_op.h"
#include "tma.cuh"

namespace dda {

#ifndef DDA_HOST_EMU

namespace mrhs {

const int NR = 12;                       // right-hand sides
const int NB = 32;                       // N of the MMA (24 used: [Re | Im], padded to a multiple of 16)
const int TMEM_COLS = 512;               // 5 accumulator sets x 96 columns (see ACC below), power of two
const int ACC = 96;                      // columns per accumulator set: [hi*hi | hi*lo] (64, one MMA with N = 64) + [lo*hi] (32)

template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  // mbarrier wait with a bound: a protocol error traps (CUDA error on the host) instead of hanging the GPU.  BACKOFF: roles
  // that wait for a whole block / site sleep between polls and leave the issue slots to the worker warps.
  for (long it = 0; it < (1L << 26); it++) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (BACKOFF) __nanosleep(64);
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

// Operand layouts (floats), all the canonical no-swizzle K-major layout of tcgen05.mma (verified on the B200 with
// scripts/umma_probe.cu: 128-byte core matrices = 8 rows x 4 K-elements, leading byte offset = distance of the cores
// adjacent in K, stride byte offset = distance of the 8-row groups).  A block is the real matrix W[rho][c] (rho = 2 r + re|im):
//   forward image:  M = rho (n/4 row groups), K = c   (n/4 cores per row group)   -- the block TRANSPOSED, 2 n^2 floats
//   daggered image: M = c   (n/8 row groups), K = rho (n/2 cores per row group)   -- 2 n^2 floats
//   Bf: [Re V | Im V], N = 32 right-hand-side slots (4 row groups), K = c;   Bd: [w | -i w], K = rho; hi rows 0..31, lo rows 32..63
// The MMAs run with M = 128: operand rows beyond the image read whatever follows it in shared memory and produce accumulator
// rows nobody reads (an accumulator row depends on its own operand row only).
__host__ __device__ inline long fwd_off(int rho, int c, int kcf) { return ((long)(rho >> 3) * kcf + (c >> 2)) * 32 + (rho & 7) * 4 + (c & 3); }
__host__ __device__ inline long dag_off(int c, int rho, int kcd) { return ((long)(c >> 3) * kcd + (rho >> 2)) * 32 + (c & 7) * 4 + (rho & 3); }

// images of site x: [fwd S][fwd F_0, dag F_0] ... [fwd F_3, dag F_3], 18 n^2 floats
__global__ void __launch_bounds__(256) k_mrhs_tile(CoarseOp op, float *__restrict__ T) {
  const int n = op.n, n2 = 2 * n, kcf = n / 4, kcd = n / 2, m = blockIdx.y;
  const long nn = (long)n * n, x = blockIdx.x;
  const float *W = reinterpret_cast<const float *>(m == 0 ? op.S + x * nn : op.F + (x * 4 + (m - 1)) * nn);
  float *dst = T + x * 18 * nn + (m == 0 ? 0 : 2 * nn + (long)(m - 1) * 4 * nn);
  for (int e = threadIdx.x; e < 2 * nn; e += blockDim.x) {
    const int c = e / n2, rho = e - c * n2;
    const float v = W[e];
    dst[fwd_off(rho, c, kcf)] = v;
    if (m > 0) dst[2 * nn + dag_off(c, rho, kcd)] = v;
  }
}

// shared-memory layout (bytes): STAGES operand sets [hi fwd 2nn][hi dag 2nn][lo fwd 2nn][lo dag 2nn][Bf hi 32n][Bf lo 32n] (floats),
// 2 x [Bd hi 64n][Bd lo 64n], RAW landing buffers of 4nn floats (RAW = 0: the bulk copies land in the hi buffers of the operand
// set and are split in place), slack for the M = 128 reads past the last daggered image, mbarriers
struct Layout { size_t stage_f, bd_f, raw_f, bar_off, total; };
__host__ __device__ inline Layout layout(int n, int stages, int raw) {
  Layout L;
  const size_t nn = (size_t)n * n;
  L.stage_f = 8 * nn + 64 * (size_t)n; L.bd_f = 128 * (size_t)n; L.raw_f = 4 * nn;
  const size_t bytes = (stages * L.stage_f + 2 * L.bd_f + raw * L.raw_f) * sizeof(float);
  const size_t overrun = (size_t)(16 - n / 8) * 64 * n, following = (64 * (size_t)n + 2 * L.bd_f + raw * L.raw_f) * sizeof(float);
  const size_t slack = overrun > following ? ((overrun - following + 127) / 128) * 128 : 0;
  L.bar_off = bytes + slack;
  L.total = L.bar_off + 10 * sizeof(uint64_t) + 16;
  return L;
}

const int NW = 256;                      // worker threads (warps 0-7)
const int NT = NW + 128 + 64;            // + epilogue warps 8-11 (TMEM lane quarter = warp % 4) + producer warp 12 + issuer warp 13

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

template <int STAGES, int RAW>
__global__ void __launch_bounds__(NT, 1)
k_coarse_mrhs(CoarseOp op, const float *__restrict__ T, cf *__restrict__ out, const cf *__restrict__ in, cf *__restrict__ Z,
              long vstride, long zstride, int nsites) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, nh = n / 2, n2 = 2 * n;
  const int kcf = n / 4, kcd = n2 / 4;                              // K cores of the forward / daggered operands
  // stage: [hi fwd 2nn][hi dag 2nn][lo fwd 2nn][lo dag 2nn][Bf hi 32n][Bf lo 32n];  then 2 x [Bd hi 64n][Bd lo 64n]
  const Layout lay = layout(n, STAGES, RAW);
  const int SF = (int)lay.stage_f, BDF = (int)lay.bd_f, RF = (int)lay.raw_f;
  float *stage0 = reinterpret_cast<float *>(smem_raw);
  float *Bd0 = stage0 + (size_t)STAGES * SF;
  float *raw0 = Bd0 + 2 * (size_t)BDF;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + lay.bar_off);
  uint64_t *ready = full + 2, *empty = ready + 2, *rawfree = empty + 2, *acc_full = rawfree + 2, *acc_empty = acc_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_sites = (nsites - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = 5 * my_sites;

  for (int q = tid; q < STAGES * SF + 2 * BDF; q += NT) stage0[q] = 0.f;        // unused right-hand-side slots stay zero
  static_assert(RAW == 0 || RAW == STAGES, "landing buffers are owned by the worker group of the same index");
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1); mbar_init(&ready[s], NW / STAGES / 32); mbar_init(&empty[s], 1); mbar_init(&rawfree[s], NW / STAGES / 32);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 12) {
    // ---------------- producer: one bulk copy per block ----------------
    if (lane == 0) {
      for (int j = 0; j < total; j++) {
        const int k = j / 5, m = j - 5 * k, s = j % STAGES, u = j / STAGES;
        const long x = (long)blockIdx.x + (long)k * gridDim.x;
        // the landing buffer is free: RAW: the workers have split its previous block into the operand set; in place: the MMAs
        // that read the operand set have completed
        if (u > 0) mbar_wait_bounded<true>(RAW ? &rawfree[s] : &empty[s], (uint32_t)((u - 1) & 1));
        const uint32_t bytes = (uint32_t)((m == 0 ? 2 : 4) * nn * sizeof(float));
        const float *src = T + x * 18 * (long)nn + (m == 0 ? 0 : 2 * (long)nn + (long)(m - 1) * 4 * nn);
        mbar_expect_tx(&full[s], bytes);
        tma_bulk_g2s(RAW ? raw0 + (size_t)s * RF : stage0 + (size_t)s * SF, src, bytes, &full[s]);
      }
    }
  } else if (warp == 13) {
    // ---------------- issuer: A and B K-major.  TF32 x 3 with two MMAs per K step into INDEPENDENT accumulators:
    // A_hi x [B_hi | B_lo] (N = 64: the hi and lo operand buffers of B are adjacent row groups) and A_lo x B_hi (N = 32),
    // summed in the epilogue ----------------
    if (lane == 0) {
      const uint32_t idesc64 = instr_desc(0, 0, 128, 2 * NB), idesc32 = instr_desc(0, 0, 128, NB);
      const uint32_t sbf = (uint32_t)kcf * 128, sbd = (uint32_t)kcd * 128;
      for (int j = 0; j < total; j++) {
        const int k = j / 5, m = j - 5 * k, s = j % STAGES, u = j / STAGES;
        mbar_wait_bounded(&ready[s], (uint32_t)(u & 1));
        if (m == 0 && k > 0) mbar_wait_bounded(acc_empty, (uint32_t)((k - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // descriptors: the start-address field counts 16-byte units, a K step of 8 (two cores, 256 bytes) adds 16 to it
        const uint32_t a0 = smem_u32(stage0 + (size_t)s * SF);
        if (m > 0) {
          // daggered: D_mu[c][slot] = sum_rho W[rho][c] B'[slot][rho]; K = 2n
          const uint32_t td = tmem + (uint32_t)(ACC * m);
          uint64_t ah = smem_desc(a0 + 8u * nn, 128, sbd), al = smem_desc(a0 + 24u * nn, 128, sbd);
          uint64_t bhl = smem_desc(smem_u32(Bd0 + (size_t)(k & 1) * BDF), 128, sbd);
          for (int ks = 0; ks < n2 / 8; ks++, ah += 16, al += 16, bhl += 16) {
            mma_tf32(td, ah, bhl, idesc64, ks > 0 ? 1u : 0u);
            mma_tf32(td + 64u, al, bhl, idesc32, ks > 0 ? 1u : 0u);
          }
        }
        // forward: D_fwd[rho][slot] += sum_c W[rho][c] B[slot][c]; K = n
        {
          uint64_t ah = smem_desc(a0, 128, sbf), al = smem_desc(a0 + 16u * nn, 128, sbf), bhl = smem_desc(a0 + 32u * nn, 128, sbf);
          for (int ks = 0; ks < n / 8; ks++, ah += 16, al += 16, bhl += 16) {
            mma_tf32(tmem, ah, bhl, idesc64, (m > 0 || ks > 0) ? 1u : 0u);
            mma_tf32(tmem + 64u, al, bhl, idesc32, (m > 0 || ks > 0) ? 1u : 0u);
          }
        }
        mma_commit(&empty[s]);                                       // arrives when every MMA issued so far has completed
        if (m == 4) mma_commit(acc_full);
      }
    }
  } else if (warp >= 8) {
    // ---------------- epilogue: accumulators -> registers -> global.  A warp reaches the TMEM lanes 32 (warp % 4) .. +31 =
    // rows of the accumulators ----------------
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    float v[32], v2[32];
    auto load_sum = [&](uint32_t base) {        // [hi*hi] + [hi*lo] + [lo*hi]
      tmem_ld32(base, v);
      tmem_ld32(base + 32u, v2);
#pragma unroll
      for (int q = 0; q < 32; q++) v[q] += v2[q];
      tmem_ld32(base + 64u, v2);
#pragma unroll
      for (int q = 0; q < 32; q++) v[q] += v2[q];
    };
    for (int k = 0; k < my_sites; k++) {
      const long x = (long)blockIdx.x + (long)k * gridDim.x;
      mbar_wait_bounded<true>(acc_full, (uint32_t)(k & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      load_sum(tmem + lane_base);
      {
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
      for (int mu = 0; mu < 4; mu++) {
        load_sum(tmem + lane_base + (uint32_t)(ACC * (1 + mu)));
        const int c = row;
        if (c < n) {
          const float sg = (c < nh) ? 1.f : -1.f;
#pragma unroll
          for (int j = 0; j < NR; j++) Z[(long)j * zstride + (x * 4 + mu) * n + c] = cf(sg * v[j], sg * v[NR + j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);                          // accumulators are free for the next site
    }
  } else {
    // ---------------- workers: STAGES groups of NW / STAGES threads, group g owns stage g = the blocks j = g (mod STAGES), so
    // that the latency chains of consecutive blocks (wait, global loads, shared-memory round trips) overlap ----------------
    const int NWG = NW / STAGES, NPRE = (16 * 64 + NWG - 1) / NWG;
    const int g = tid / NWG, gt = tid - g * NWG;
    // a thread's right-hand-side elements: slot q = gt + NWG i -> (jr, c) with c % 4 = q % 4 and jr % 8 = (q / 4) % 8, i.e. the 32
    // lanes of a warp store to 32 distinct banks of the core-matrix layout (8 rows x 4 K-elements).  Offsets are per thread
    // constants: element offset in a vector, operand offsets of Re (row jr) / Im (row 12 + jr) in Bf and of the two rows in Bd.
    long voff[NPRE]; int o0[NPRE], o1[NPRE], d0[NPRE], d1[NPRE]; bool hi_half[NPRE];
#pragma unroll
    for (int i = 0; i < NPRE; i++) {
      const int q = gt + NWG * i, rest = q >> 5;
      const int jr = (rest & 1) * 8 + ((q >> 2) & 7), c = (rest >> 1) * 4 + (q & 3);
      const bool ok = jr < NR && c < n;
      const int row1 = NR + jr, kk = 2 * c;
      voff[i] = ok ? (long)jr * vstride + c : -1;
      o0[i] = ((jr >> 3) * kcf + (c >> 2)) * 32 + (jr & 7) * 4 + (c & 3);
      o1[i] = ((row1 >> 3) * kcf + (c >> 2)) * 32 + (row1 & 7) * 4 + (c & 3);
      d0[i] = ((jr >> 3) * kcd + (kk >> 2)) * 32 + (jr & 7) * 4 + (kk & 3);
      d1[i] = ((row1 >> 3) * kcd + (kk >> 2)) * 32 + (row1 & 7) * 4 + (kk & 3);
      hi_half[i] = c >= nh;
    }
    // the right-hand sides a block multiplies (12 x n complex) are fetched one block of this group ahead into registers, the
    // neighbour index two ahead
    cf pre[NPRE];
    auto source = [&](int j) -> int {                                // vectors of block j: site x (S) or x + mu (F_mu)
      const int k = j / 5, m = j - 5 * k;
      const long x = (long)blockIdx.x + (long)k * gridDim.x;
      return (m == 0) ? (int)x : __ldg(op.nb + (long)(m - 1) * op.V + x);
    };
    auto prefetch = [&](int src) {
#pragma unroll
      for (int i = 0; i < NPRE; i++)
        if (voff[i] >= 0) pre[i] = in[voff[i] + (long)src * n];
    };
    int src_next = 0;
    if (g < total) prefetch(source(g));
    if (g + STAGES < total) src_next = source(g + STAGES);
    for (int j = g; j < total; j += STAGES) {
      const int k = j / 5, m = j - 5 * k, s = g, u = j / STAGES;
      float *stg = stage0 + (size_t)s * SF;
      float *Bfh = stg + 8 * nn, *Bfl = Bfh + 32 * n;
      float *Bdh = Bd0 + (size_t)(k & 1) * BDF, *Bdl = Bdh + 64 * n;
      if (u > 0) mbar_wait_bounded(&empty[s], (uint32_t)((u - 1) & 1));   // the MMAs that read this stage's B operands (and, in
                                                                         // order, everything before them) have completed
      // forward B = [Re V | Im V] from the prefetched registers; for m = 0 these are the site's own vectors, from which
      // the daggered operand B' = [w | -i w], w = G5 V(x), is filled as well (rows j and 12 + j, K = rho = 2 r + re|im)
#pragma unroll
      for (int i = 0; i < NPRE; i++) {
        if (voff[i] >= 0) {
          const cf v = pre[i];
          float hi, lo;
          split_tf32(v.re, hi, lo); Bfh[o0[i]] = hi; Bfl[o0[i]] = lo;
          split_tf32(v.im, hi, lo); Bfh[o1[i]] = hi; Bfl[o1[i]] = lo;
          if (m == 0) {
            const cf w = hi_half[i] ? -v : v;
            split_tf32(w.re, hi, lo);  Bdh[d0[i]] = hi;     Bdl[d0[i]] = lo;        // row jr:      [ w.re,  w.im ]
            split_tf32(w.im, hi, lo);  Bdh[d0[i] + 1] = hi; Bdl[d0[i] + 1] = lo;
            Bdh[d1[i]] = hi;           Bdl[d1[i]] = lo;                             // row 12 + jr: [ w.im, -w.re ]
            split_tf32(-w.re, hi, lo); Bdh[d1[i] + 1] = hi; Bdl[d1[i] + 1] = lo;
          }
        }
      }
      if (j + STAGES < total) {
        prefetch(src_next);
        if (j + 2 * STAGES < total) src_next = source(j + 2 * STAGES);
      }
      mbar_wait_bounded(&full[s], (uint32_t)(u & 1));
      // TF32 split of the image(s) into the operand set (RAW = 0: in place): hi = value truncated to 11 significant bits,
      // lo = value - hi
      float4 *H4 = reinterpret_cast<float4 *>(stg), *L4 = H4 + nn;
      const float4 *R4 = RAW ? reinterpret_cast<const float4 *>(raw0 + (size_t)s * RF) : H4;
      const int cnt = (m == 0 ? nn : 2 * nn) / 2;                        // float4 elements
      int i = gt;
      for (; i + 3 * NWG < cnt; i += 4 * NWG) {                          // four independent load -> split -> store chains
        float4 v[4], h[4], l[4];
#pragma unroll
        for (int t = 0; t < 4; t++) v[t] = R4[i + t * NWG];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          split_tf32(v[t].x, h[t].x, l[t].x); split_tf32(v[t].y, h[t].y, l[t].y);
          split_tf32(v[t].z, h[t].z, l[t].z); split_tf32(v[t].w, h[t].w, l[t].w);
        }
#pragma unroll
        for (int t = 0; t < 4; t++) { H4[i + t * NWG] = h[t]; L4[i + t * NWG] = l[t]; }
      }
      for (; i < cnt; i += NWG) {
        const float4 v = R4[i];
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        H4[i] = h; L4[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand buffers written by the generic proxy
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ready[s]); if (RAW) mbar_arrive(&rawfree[s]); }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS));
}

}  // namespace mrhs

// coarse_kernel.cu: eta_j(x) += sum_mu Z_j[x-mu][mu] for nrhs vectors in one launch
void coarse_combine_batch(const CoarseOp &op, cf *out, const cf *in, const cf *Z, int nrhs, long vstride, long zstride);

bool coarse_mrhs_supported(const CoarseOp &op) { return !(op.n > 64 || op.n < 8 || (op.n & 7) || op.V <= 0); }

// operator images of the 12-RHS kernel (72 n^2 bytes per site); rebuild after every change of op.S / op.F.  dev_free() it.
float *coarse_mrhs_tile(const CoarseOp &op) {
  if (!coarse_mrhs_supported(op)) return nullptr;
  float *T = dev_alloc<float>((size_t)op.V * 18 * op.n * op.n);
  mrhs::k_mrhs_tile<<<dim3((unsigned)op.V, 5), 256, 0, g_stream>>>(op, T);
  g_launch_count++;
  CUDA_CHECK(cudaGetLastError());
  return T;
}

// out_j = D_c in_j for j < 12: vectors j at in + j * vstride / out + j * vstride (vstride >= (V + ghost sites) * n), Z: scratch
// of 12 x zstride complex (zstride >= 4 n V), T: coarse_mrhs_tile(op).  Halo slabs of the inputs must be current.  Returns
// false when the shape is not supported.
bool coarse_apply_mrhs(const CoarseOp &op, const float *T, cf *out, const cf *in, cf *Z, long vstride, long zstride) {
  const int n = op.n;
  if (!coarse_mrhs_supported(op) || !T) return false;
  // configurations: two operand sets + two landing buffers (n <= 40), two operand sets split in place (n = 48), one (n >= 56)
  const size_t limit = 226 * 1024;                                   // 227 KB per CTA minus the kernel's static shared memory
  int cfg = -1;
  const int stages_of[3] = {2, 2, 1}, raw_of[3] = {2, 0, 0};
  for (int c = 0; c < 3 && cfg < 0; c++)
    if (mrhs::layout(n, stages_of[c], raw_of[c]).total <= limit) cfg = c;
  static const char *force = getenv("DDA_MRHS_CONFIG");              // A/B runs: 0, 1, 2
  if (force && atoi(force) > cfg && atoi(force) < 3) cfg = atoi(force);
  if (cfg < 0) return false;
  const size_t smem = mrhs::layout(n, stages_of[cfg], raw_of[cfg]).total;
  static size_t attr[3] = {0, 0, 0};
  static int sms = 0;
  if (!sms) sms = dev_sm_count();
  const long grid = std::min<long>(op.V, (long)sms);                 // one CTA per SM: the CTA owns all 512 TMEM columns
  auto launch = [&](auto kern) {
    if (smem > attr[cfg]) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[cfg] = smem; }
    kern<<<(unsigned)grid, mrhs::NT, smem, g_stream>>>(op, T, out, in, Z, vstride, zstride, (int)op.V);
  };
  if (cfg == 0) launch(mrhs::k_coarse_mrhs<2, 2>);
  else if (cfg == 1) launch(mrhs::k_coarse_mrhs<2, 0>);
  else launch(mrhs::k_coarse_mrhs<1, 0>);
  g_launch_count++;
  coarse_combine_batch(op, out, in, Z, mrhs::NR, vstride, zstride);
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
}

#endif

}  // namespace dda

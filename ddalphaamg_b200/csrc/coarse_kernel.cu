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
#include "tma.cuh"
#include <cstdint>
#include <vector>

namespace dda {

#ifndef DDA_HOST_EMU

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
        blk_forward(M, vec + m * n, n, grp, ch, p, f0r, f0i, f1r, f1i);
        if (m > 0) blk_dagger(M, vec + 5 * n, n, grp, ch, p, zr[m - 1][0], zi[m - 1][0], zr[m - 1][1], zi[m - 1][1]);   // conj(M[r][c]) w[r]
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

void coarse_combine(const CoarseOp &op, cf *out, const cf *in, const cf *Z) {
  const long total = op.V * op.n;
  k_coarse_combine<<<(unsigned)((total + 127) / 128), 128, 0, g_stream>>>(op, out, in, Z, total);
  g_launch_count++;
}
// the same for nrhs vectors in one launch (vector j at out + j * vstride, in + j * vstride, Z + j * zstride)
__global__ void k_coarse_combine_batch(CoarseOp op, cf *__restrict__ out, const cf *__restrict__ in, const cf *__restrict__ Z, long total,
                                       long vstride, long zstride) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = op.n, nh = n / 2;
  const long x = i / n; const int r = (int)(i - x * n);
  const long nn = (long)n * n;
  out += blockIdx.y * vstride; in += blockIdx.y * vstride; Z += blockIdx.y * zstride;
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
void coarse_combine_batch(const CoarseOp &op, cf *out, const cf *in, const cf *Z, int nrhs, long vstride, long zstride) {
  const long total = op.V * op.n;
  k_coarse_combine_batch<<<dim3((unsigned)((total + 127) / 128), (unsigned)nrhs), 128, 0, g_stream>>>(op, out, in, Z, total, vstride, zstride);
  g_launch_count++;
}

// -------------------------------------------------------------------------------------------------------------------
// Fused SAP block solve on an intermediate level: ONE CTA per Schwarz block runs all block_iter minimal-residual steps
//   Dr = D_block r ;  alpha = <Dr,r>/<Dr,Dr> ;  e += alpha r ;  r -= alpha Dr          and finally   x += e.
// Reference counterparts: red_black_schwarz_PRECISION with the coarse block operator (schwarz_generic.c:1260-1431,
// coarse_block_operator_PRECISION coarse_operator_generic.c:208-236), local_minres_PRECISION (linsolve_generic.c:985-1029).
// The block vectors (bs x n complex each) stay in shared memory; the block operator is streamed once per MR step
// through the same TMA ring as k_coarse_full: the self coupling of every site and every in-block link exactly once
// (forward product for the link's source site, daggered product for its target site), partial sums added to Dr with
// shared-memory atomics.  The generic path needs 4 launches per MR step and reads every in-block link twice.
struct SapJob { int type, i, j, pad; };        // type 0: self coupling of local site i; 1+mu: link i -> j = i+mu

// TEAMS x 128 threads: every team of 128 streams its own share of the jobs (job q -> team q % TEAMS) through its own ring;
// all teams add into the same Dr.  TEAMS = 4 when a rank has fewer blocks than SMs (strong-scaling limit), else 1.
template <int STAGES, int TEAMS>
__global__ void __launch_bounds__(128 * TEAMS)
k_coarse_sap_mr(CoarseOp op, cf *__restrict__ x, const cf *__restrict__ rin, const int *__restrict__ blocklist, int nblk, int bs,
                int biter, const SapJob *__restrict__ jobs, int njobs, int G) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, nh = n / 2, P = n / 2, ch = n / G;
  const int len = bs * n;
  cf *Ms0 = reinterpret_cast<cf *>(smem_raw);                   // [TEAMS][STAGES][n*n]
  cf *rv = Ms0 + (size_t)TEAMS * STAGES * nn;                   // block residual
  cf *rg = rv + len;                                            // gamma5 * residual
  cf *ev = rg + len;                                            // accumulated correction
  cf *Dr = ev + len;                                            // D_block r
  float *red = reinterpret_cast<float *>(Dr + len);             // [16 warps][4]
  SapJob *sj = reinterpret_cast<SapJob *>(red + 64);            // [njobs]
  uint64_t *full0 = reinterpret_cast<uint64_t *>(sj + njobs);   // [TEAMS][STAGES]
  const int NT = 128 * TEAMS;
  const int gtid = threadIdx.x, team = gtid >> 7, tid = gtid & 127, lane = tid & 31, w = gtid >> 5;
  cf *Ms = Ms0 + (size_t)team * STAGES * nn;
  uint64_t *full = full0 + team * STAGES;
  const int grp = tid / P, p = tid - grp * P;
  const bool active = grp < G;
  const uint32_t bytes = (uint32_t)(nn * sizeof(cf));
  const int myjobs = (njobs - team + TEAMS - 1) / TEAMS;        // jobs team, team + TEAMS, ... of every MR step
  // persistent over the listed blocks: CTA b handles blocks b, b + gridDim, ...  With the grid capped so that the block
  // operators of the concurrently processed blocks fit in L2 (DDA_SAPMR_GRID), MR steps 2..biter stream from L2
  const int my_blocks = (nblk - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int per_block = biter * myjobs;
  const int total = per_block * my_blocks;
  for (int q = gtid; q < njobs; q += NT) sj[q] = jobs[q];
  if (gtid == 0) {
    for (int s = 0; s < STAGES * TEAMS; s++) mbar_init(&full0[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto team_sync = [&]() {
    if (TEAMS == 1) __syncthreads();
    else asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
  };
  auto issue = [&](int jj) {                                     // jj-th job of this team (running over all its blocks)
    const int bi = jj / per_block;
    const long base = (long)blocklist[blockIdx.x + bi * gridDim.x] * bs;
    const SapJob jb = sj[((jj - bi * per_block) % myjobs) * TEAMS + team];
    const cf *src = (jb.type == 0) ? op.S + (base + jb.i) * nn : op.F + ((base + jb.i) * 4 + (jb.type - 1)) * nn;
    const int st = jj % STAGES;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(Ms + (size_t)st * nn, src, bytes, &full[st]);
  };
  if (tid == 0) for (int jj = 0; jj < STAGES && jj < total; jj++) issue(jj);
  const float sgc = (2 * p < nh) ? 1.f : -1.f;                  // gamma5 sign of the thread's daggered output columns

  int jj = 0;
  for (int bi = 0; bi < my_blocks; bi++) {
  const long base = (long)blocklist[blockIdx.x + bi * gridDim.x] * bs;
  for (int q = gtid; q < len; q += NT) {
    const cf v = rin[base * n + q];
    rv[q] = v; rg[q] = ((q % n) < nh) ? v : -v; ev[q] = cf(0.f, 0.f); Dr[q] = cf(0.f, 0.f);
  }
  __syncthreads();
  for (int it = 0; it < biter; it++) {
    for (int q = 0; q < myjobs; q++, jj++) {
      const int st = jj % STAGES;
      const SapJob jb = sj[q * TEAMS + team];
      mbar_wait(&full[st], (uint32_t)((jj / STAGES) & 1));
      const cf *M = Ms + (size_t)st * nn;
      if (active) {
        {                                 // forward: Dr[i] += M v,  v = r[i] (self coupling) or r[j] (link)
          const int src = (jb.type == 0) ? jb.i : jb.j;
          float f0r = 0.f, f0i = 0.f, f1r = 0.f, f1i = 0.f;
          blk_forward(M, rv + src * n, n, grp, ch, p, f0r, f0i, f1r, f1i);
          float *d = reinterpret_cast<float *>(Dr + jb.i * n + 2 * p);
          atomicAdd(d, f0r); atomicAdd(d + 1, f0i); atomicAdd(d + 2, f1r); atomicAdd(d + 3, f1i);
        }
        if (jb.type > 0) {                // daggered: Dr[j] += gamma5 M^H gamma5 r[i]
          float a0r, a0i, a1r, a1i;
          blk_dagger(M, rg + jb.i * n, n, grp, ch, p, a0r, a0i, a1r, a1i);
          float *d = reinterpret_cast<float *>(Dr + jb.j * n + 2 * p);
          atomicAdd(d, sgc * a0r); atomicAdd(d + 1, sgc * a0i); atomicAdd(d + 2, sgc * a1r); atomicAdd(d + 3, sgc * a1i);
        }
      }
      team_sync();
      if (tid == 0 && jj + STAGES < total) issue(jj + STAGES);
    }
    if (TEAMS > 1) __syncthreads();                              // all teams' contributions are in Dr
    // alpha = <Dr,r>/<Dr,Dr> over the block (local_xy_over_xx, linalg_generic.c:158-169)
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
    for (int q = gtid; q < len; q += NT) {
      const cf a = Dr[q], b = rv[q];
      p0 += a.re * b.re + a.im * b.im; p1 += a.re * b.im - a.im * b.re; p2 += a.re * a.re + a.im * a.im;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    if (lane == 0) { red[4 * w] = p0; red[4 * w + 1] = p1; red[4 * w + 2] = p2; }
    __syncthreads();
    p0 = p1 = p2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * TEAMS; k++) { p0 += red[4 * k]; p1 += red[4 * k + 1]; p2 += red[4 * k + 2]; }
    cf alpha(0.f, 0.f);
    if (p2 > 1e-30f) alpha = cf(p0 / p2, p1 / p2);
    for (int q = gtid; q < len; q += NT) {
      cf e = ev[q], r = rv[q];
      const cf d = Dr[q];
      fma_(e, alpha, r); fms_(r, alpha, d);
      ev[q] = e; rv[q] = r; rg[q] = ((q % n) < nh) ? r : -r; Dr[q] = cf(0.f, 0.f);
    }
    __syncthreads();
  }
  for (int q = gtid; q < len; q += NT) x[base * n + q] += ev[q];
  __syncthreads();                                               // block vectors are re-initialised by the next block
  }
}

bool coarse_sap_mr_fast(const CoarseOp &op, cf *x, const cf *r, const int *d_blocklist, int nblk, int bs, int biter,
                        const int *d_jobs, int njobs) {
  const int n = op.n;
  if (n > 64 || n < 8 || (n & 3) || nblk <= 0 || biter <= 0 || !d_jobs || njobs <= 0) return false;
  int G = 128 / (n / 2);
  while (G > 1 && n % (2 * G) != 0) G--;
  if (n % (2 * G) != 0) return false;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const size_t nn = (size_t)n * n, len = (size_t)bs * n;
  int teams = (nblk <= 2 * sms) ? 4 : 1;            // few blocks per rank: more threads per block instead of more blocks per SM
  if (const char *e = getenv("DDA_SAP_TEAMS")) { const int t = atoi(e); if (t == 1 || t == 4) teams = t; }   // test / tuning override
  static int gridcap = -1;                          // DDA_SAPMR_GRID: persistent CTAs, at most this many (0: one CTA per block)
  if (gridcap < 0) { const char *e = getenv("DDA_SAPMR_GRID"); gridcap = e ? atoi(e) : 0; }
  const int grid = (gridcap > 0 && gridcap < nblk) ? gridcap : nblk;
  static int stages = 0;                            // depth of the TMA ring (DDA_SAPMR_STAGES = 2 | 3; tuning knob)
  if (!stages) { const char *e = getenv("DDA_SAPMR_STAGES"); stages = (e && atoi(e) == 3) ? 3 : 2; }
  const size_t smem = (size_t)teams * stages * nn * sizeof(cf) + 4 * len * sizeof(cf) + 64 * sizeof(float) + njobs * sizeof(SapJob) + 16 * sizeof(uint64_t);
  if (smem > 200 * 1024) return false;
  const SapJob *jb = reinterpret_cast<const SapJob *>(d_jobs);
  static size_t attr[4] = {0, 0, 0, 0};             // per kernel variant
  auto go = [&](auto kern, int threads, int variant) {
    if (smem > attr[variant]) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[variant] = smem; }
    kern<<<grid, threads, smem, g_stream>>>(op, x, r, d_blocklist, nblk, bs, biter, jb, njobs, G);
  };
  if (teams == 4 && stages == 2) go(k_coarse_sap_mr<2, 4>, 512, 0);
  else if (teams == 4) go(k_coarse_sap_mr<3, 4>, 512, 1);
  else if (stages == 2) go(k_coarse_sap_mr<2, 1>, 128, 2);
  else go(k_coarse_sap_mr<3, 1>, 128, 3);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
}

// -------------------------------------------------------------------------------------------------------------------
// Soo^-1 on the odd sites of the coarsest level: ONE CTA per site, the n x n block in shared memory in double, in-place
// Gauss-Jordan without pivoting (the reference factorises LU without pivoting as well, coarse_oddeven_generic.c:24-73;
// the explicit inverse replaces its forward / backward substitution, coarse_perform_fwd_bwd_subs :75-121).
__global__ void __launch_bounds__(256) k_invert_self(CoarseOp op, long n_first, long nsites) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = op.n; const long nn = (long)n * n;
  cd *A = reinterpret_cast<cd *>(smem_raw);                       // row-major [n][n]
  cd *colp = A + nn;                                              // column p of the current step
  const long o = blockIdx.x;
  if (o >= nsites) return;
  const cf *M = op.S + (n_first + o) * nn;
  for (int q = threadIdx.x; q < nn; q += blockDim.x) { const int c = q / n, r = q - c * n; const cf v = M[q]; A[(long)r * n + c] = cd(v.re, v.im); }
  __syncthreads();
  for (int p = 0; p < n; p++) {
    const cd piv = A[(long)p * n + p];
    const double d = 1.0 / norm2(piv);
    const cd ip(piv.re * d, -piv.im * d);
    for (int r = threadIdx.x; r < n; r += blockDim.x) colp[r] = A[(long)r * n + p];
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) A[(long)p * n + c] = (c == p) ? ip : A[(long)p * n + c] * ip;   // pivot row
    __syncthreads();
    for (int q = threadIdx.x; q < nn; q += blockDim.x) {
      const int r = q / n, c = q - r * n;
      if (r == p) continue;
      const cd f = colp[r];
      if (c == p) A[q] = -(f * ip);
      else fms_(A[q], f, A[(long)p * n + c]);
    }
    __syncthreads();
  }
  cf *O = op.Sinv + o * nn;
  for (int q = threadIdx.x; q < nn; q += blockDim.x) { const int c = q / n, r = q - c * n; const cd v = A[(long)r * n + c]; O[q] = cf((float)v.re, (float)v.im); }
}

bool coarse_invert_odd_self_fast(CoarseOp &op) {
  const int n = op.n; const long nodd = op.V - op.n_even;
  const size_t smem = ((size_t)n * n + n) * sizeof(cd);
  if (nodd <= 0 || smem > 200 * 1024) return false;
  static size_t attr = 0;
  if (smem > attr) { CUDA_CHECK(cudaFuncSetAttribute(k_invert_self, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  k_invert_self<<<(unsigned)nodd, 256, smem, g_stream>>>(op, op.n_even, nodd);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
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

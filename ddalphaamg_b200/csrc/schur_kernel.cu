// schur_kernel.cu -- coarsest level: even-odd Schur complement of the coarse operator as three streaming kernels (sm_100a).
//
//   out_e = S_ee in_e - N_eo Soo^-1 N_oe in_e ,   (N v)(x) = sum_mu [ F_mu(x) v(x+mu) + G5 F_mu(x-mu)^H G5 v(x-mu) ]
// Reference counterparts: coarse_apply_schur_complement_PRECISION / coarse_solve_odd_even_PRECISION
// (coarse_oddeven_generic.c:1139-1189), coarse_hopping_term / coarse_n_hopping_term (:447-728), coarse_diag_ee /
// coarse_diag_oo_inv (:123-198; LU substitution there, an explicit inverse here), dense kernels coarse_hopp /
// coarse_daggered_hopp (coarse_operator_generic.h:119-172).
//
// The work is pure streaming of dense n x n blocks.  Each half application N_oe / N_eo touches every hop matrix
// exactly ONCE (scatter form, like k_coarse_full): the sites of the TARGET parity multiply F_mu(x) with the
// neighbour's vector (forward product, accumulated over mu into dir(x)), the sites of the SOURCE parity multiply
// F_mu(x)^H with their own vector (daggered product) and leave the result for site x+mu in a scratch Z[x][mu]; the
// next kernel of the chain adds the four Z entries while it loads its input, so no separate combine pass exists:
//   k_schur_hop<.>(phase 0)  odd:  dir = sum F in_e(x+mu)           even: Z = G5 F^H G5 in_e(x)
//   k_schur_mid              odd:  t1 = -Soo^-1 (dir + sum Z)
//   k_schur_hop<.>(phase 1)  even: dir = S in_e + sum F t1(x+mu)    odd:  Z = G5 F^H G5 t1(x)
//   k_schur_fin              even: out = dir + sum Z
// Persistent CTAs of 128 threads, blocks through a ring of shared-memory stages filled by TMA bulk copies
// (cp.async.bulk + mbarrier), register tiling as in coarse_kernel.cu.  Every kernel returns at once when *skip != 0
// (steps enqueued past convergence by the device-resident GMRES, dev_gmres.h).
// Algorithmic traffic of one Schur application: 2 x (4 n^2) x V (every hop matrix twice) + n^2 x V (S_ee, Soo^-1), x 8 B.
#include "coarse_op.h"
#include "dev_gmres.h"
#include "tma.cuh"

namespace dda {

#ifndef DDA_HOST_EMU

// phase 0: forward sites = odd, daggered sites = even.  phase 1: forward sites = even (plus S(x) self(x) when `self` is
// given), daggered sites = odd.  in: vector read by both products (neighbour values for the forward sites, own value
// for the daggered sites).
template <int STAGES>
__global__ void __launch_bounds__(128)
k_schur_hop(CoarseOp op, int phase, const cf *__restrict__ in, const cf *__restrict__ self, cf *__restrict__ dir,
            cf *__restrict__ Z, int nsites, int G, int rev, const int *__restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, nh = n / 2, P = n / 2, ch = n / G;
  cf *Ms = reinterpret_cast<cf *>(smem_raw);                    // [STAGES][n*n]
  cf *vec = Ms + (size_t)STAGES * nn;                           // [5][n]: own value (self or G5 in), in(x+mu) x 4
  cf *part = vec + 5 * n;                                       // [4][G][n] partial sums
  uint64_t *full = reinterpret_cast<uint64_t *>(part + 4 * G * n);
  const int tid = threadIdx.x;
  const int grp = tid / P, p = tid - grp * P;
  const bool active = grp < G;
  const uint32_t bytes = (uint32_t)(nn * sizeof(cf));
  const int ne = (int)op.n_even;
  const int my_sites = (nsites - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool with_self = (phase == 1) && (self != nullptr);
  // rev: walk the lattice backwards.  The two half applications of a Schur complement run in opposite directions, so the
  // hop matrices streamed last by one are the first the next one needs: they are still in the 126 MB L2 (the operator of
  // a 48^3 x 96 hierarchy is 238 MB), which saves about half of the DRAM traffic.
  auto site_of = [&](int k) { const int q = (int)blockIdx.x + k * (int)gridDim.x; return rev ? nsites - 1 - q : q; };
  auto is_fwd = [&](int x) { return (x >= ne) == (phase == 0); };
  auto first_slot = [&](int x) { return (is_fwd(x) && with_self) ? 0 : 1; };   // slot 0 = S(x), slots 1..4 = F_mu(x)
  // producer (thread 0): next block to fetch = slot pm of the CTA's site number pk; pc = blocks issued so far
  int pk = 0, pm = my_sites > 0 ? first_slot(site_of(0)) : 1, pc = 0;
  auto issue = [&]() {
    const long x = site_of(pk);
    const cf *src = (pm == 0) ? op.S + x * nn : op.F + (x * 4 + (pm - 1)) * nn;
    const int st = pc % STAGES;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the stage precede the async write
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(Ms + (size_t)st * nn, src, bytes, &full[st]);
    pc++; pm++;
    if (pm == 5) { pk++; pm = pk < my_sites ? first_slot(site_of(pk)) : 1; }
  };
  if (tid == 0) for (int j = 0; j < STAGES && pk < my_sites; j++) issue();

  // input vectors of the next site are fetched into registers while the current site's blocks are processed
  cf pre[3];
  auto prefetch = [&](int k) {
    const int x = site_of(k);
    const bool fw = is_fwd(x);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int q = tid + 128 * i;
      if (q < 5 * n) {
        const int vsel = q / n, c = q - vsel * n;
        cf v(0.f, 0.f);
        if (fw) {
          if (vsel == 0) { if (with_self) v = self[(long)x * n + c]; }
          else v = in[(long)op.nb[(long)(vsel - 1) * op.V + x] * n + c];
        } else if (vsel == 0) {
          v = in[(long)x * n + c];
          if (c >= nh) v = -v;
        }
        pre[i] = v;
      }
    }
  };
  if (my_sites > 0) prefetch(0);

  int cc_ = 0;                                                     // blocks consumed so far
  for (int k = 0; k < my_sites; k++) {
    const int x = site_of(k);
    const bool fw = is_fwd(x);
    const int m0 = first_slot(x);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int q = tid + 128 * i;
      if (q < 5 * n) vec[q] = pre[i];
    }
    __syncthreads();
    if (k + 1 < my_sites) prefetch(k + 1);
    float f0r = 0.f, f0i = 0.f, f1r = 0.f, f1i = 0.f;             // forward rows 2p, 2p+1 (summed over the slots)
    float zr[4][2], zi[4][2];                                     // daggered columns 2p, 2p+1 per direction
#pragma unroll
    for (int mu = 0; mu < 4; mu++) { zr[mu][0] = zr[mu][1] = zi[mu][0] = zi[mu][1] = 0.f; }
#pragma unroll
    for (int m = 0; m < 5; m++) {
      if (m >= m0) {
        const int st = cc_ % STAGES;
        mbar_wait(&full[st], (uint32_t)((cc_ / STAGES) & 1));
        const cf *M = Ms + (size_t)st * nn;
        if (active) {
          if (fw) blk_forward(M, vec + m * n, n, grp, ch, p, f0r, f0i, f1r, f1i);
          else if (m > 0) blk_dagger(M, vec, n, grp, ch, p, zr[m - 1][0], zi[m - 1][0], zr[m - 1][1], zi[m - 1][1]);
        }
        __syncthreads();
        cc_++;
        if (tid == 0 && pk < my_sites) issue();
      }
    }
    // combine the G partial sums
    if (active) {
      if (fw) *reinterpret_cast<float4 *>(part + grp * n + 2 * p) = make_float4(f0r, f0i, f1r, f1i);
      else {
#pragma unroll
        for (int mu = 0; mu < 4; mu++)
          *reinterpret_cast<float4 *>(part + (mu * G + grp) * n + 2 * p) = make_float4(zr[mu][0], zi[mu][0], zr[mu][1], zi[mu][1]);
      }
    }
    __syncthreads();
    if (fw) {
      for (int c = tid; c < n; c += 128) {
        cf a = part[c];
        for (int g2 = 1; g2 < G; g2++) a += part[g2 * n + c];
        dir[(long)x * n + c] = a;
      }
    } else {
      for (int q = tid; q < 4 * n; q += 128) {
        const int mu = q / n, c = q - mu * n;
        cf a = part[(mu * G) * n + c];
        for (int g2 = 1; g2 < G; g2++) a += part[(mu * G + g2) * n + c];
        Z[((long)x * 4 + mu) * n + c] = (c < nh) ? a : -a;
      }
    }
    // the next iteration's first __syncthreads (after the vec fill) orders these reads of `part` before its next writes
  }
}

// odd sites:  out(x) = cS * Soo^-1(x) [ a * eta(x) + b * (dir(x) + sum_mu Z[x-mu][mu]) ]      (b == 0: no hop part read)
template <int STAGES>
__global__ void __launch_bounds__(128)
k_schur_mid(CoarseOp op, const cf *__restrict__ eta, const cf *__restrict__ dir, const cf *__restrict__ Z,
            cf *__restrict__ out, float a, float b, float cS, int G, const int *__restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = op.n, nn = n * n, P = n / 2, ch = n / G;
  cf *Ms = reinterpret_cast<cf *>(smem_raw);                    // [STAGES][n*n]
  cf *vec = Ms + (size_t)STAGES * nn;                           // [n]
  cf *part = vec + n;                                           // [G][n]
  uint64_t *full = reinterpret_cast<uint64_t *>(part + G * n);
  const int tid = threadIdx.x;
  const int grp = tid / P, p = tid - grp * P;
  const bool active = grp < G;
  const uint32_t bytes = (uint32_t)(nn * sizeof(cf));
  const int ne = (int)op.n_even, nodd = (int)(op.V - op.n_even);
  const int my_sites = (nodd - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {
    const long o = (long)blockIdx.x + (long)k * gridDim.x;      // odd-site number
    const int st = k % STAGES;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(Ms + (size_t)st * nn, op.Sinv + o * nn, bytes, &full[st]);
  };
  if (tid == 0) for (int k = 0; k < STAGES && k < my_sites; k++) issue(k);
  for (int k = 0; k < my_sites; k++) {
    const long x = ne + (long)blockIdx.x + (long)k * gridDim.x;
    if (tid < n) {
      cf t(0.f, 0.f);
      if (b != 0.f) {
        t = dir[x * n + tid];
#pragma unroll
        for (int mu = 0; mu < 4; mu++) t += Z[((long)op.nb[(long)(4 + mu) * op.V + x] * 4 + mu) * n + tid];
        t = b * t;
      }
      if (a != 0.f) { const cf e = eta[x * n + tid]; t.re = __fmaf_rn(a, e.re, t.re); t.im = __fmaf_rn(a, e.im, t.im); }
      vec[tid] = t;
    }
    __syncthreads();
    const int st = k % STAGES;
    mbar_wait(&full[st], (uint32_t)((k / STAGES) & 1));
    float f0r = 0.f, f0i = 0.f, f1r = 0.f, f1i = 0.f;
    if (active) {
      blk_forward(Ms + (size_t)st * nn, vec, n, grp, ch, p, f0r, f0i, f1r, f1i);
      *reinterpret_cast<float4 *>(part + grp * n + 2 * p) = make_float4(f0r, f0i, f1r, f1i);
    }
    __syncthreads();
    if (tid == 0 && k + STAGES < my_sites) issue(k + STAGES);
    if (tid < n) {
      cf s = part[tid];
      for (int g2 = 1; g2 < G; g2++) s += part[g2 * n + tid];
      out[x * n + tid] = cS * s;
    }
    // next iteration: vec is rewritten before its __syncthreads, part after it -- both after every thread passed the
    // barrier above, and the reads of part here precede the next barrier
  }
}

// even sites:  out(x) = a * eta(x) + b * (dir(x) + sum_mu Z[x-mu][mu])
__global__ void k_schur_fin(CoarseOp op, const cf *__restrict__ eta, const cf *__restrict__ dir, const cf *__restrict__ Z,
                            cf *__restrict__ out, float a, float b, const int *__restrict__ skip) {
  if (skip && *skip) return;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int n = op.n;
  if (i >= op.n_even * n) return;
  const long x = i / n; const int c = (int)(i - x * n);
  cf t = dir[i];
#pragma unroll
  for (int mu = 0; mu < 4; mu++) t += Z[((long)op.nb[(long)(4 + mu) * op.V + x] * 4 + mu) * n + c];
  t = b * t;
  if (a != 0.f) { const cf e = eta[i]; t.re = __fmaf_rn(a, e.re, t.re); t.im = __fmaf_rn(a, e.im, t.im); }
  out[i] = t;
}

// ---- fused Arnoldi-step kernels of the coarsest-level GMRES (dev_gmres.h hooks) ---------------------------------------
// even sites: w(x) = dir(x) + sum_mu Z[x-mu][mu]  (the last stage of the Schur complement) AND the inner products
// hb[2k], hb[2k+1] += <V_k, w> over the CTA's sites for k <= j: every warp takes basis vectors k = warp, warp + 8, ...;
// double accumulation, one double atomic per CTA and value (process_multi_inner_product, linalg_generic.c:107-154).
__global__ void __launch_bounds__(256)
k_schur_fin_dots(CoarseOp op, const cf *__restrict__ dir, const cf *__restrict__ Z, cf *__restrict__ w, const cf *__restrict__ V,
                 long stride, int j, double *__restrict__ hb, int spc, const int *__restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *ws = reinterpret_cast<cf *>(smem_raw);
  const int n = op.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long x0 = (long)blockIdx.x * spc;
  const int nloc = (int)(((op.n_even - x0) < spc ? (op.n_even - x0) : spc) * n);
  for (int e = tid; e < nloc; e += 256) {
    const long x = x0 + e / n; const int c = e % n;
    cf t = dir[x * n + c];
#pragma unroll
    for (int mu = 0; mu < 4; mu++) t += Z[((long)op.nb[(long)(4 + mu) * op.V + x] * 4 + mu) * n + c];
    w[x * n + c] = t; ws[e] = t;
  }
  __syncthreads();
  for (int k = warp; k <= j; k += 8) {
    const cf *vk = V + (long)k * stride + x0 * n;
    double ar = 0.0, ai = 0.0;
    for (int e = lane; e < nloc; e += 32) {
      const cf a = vk[e], b = ws[e];
      ar += (double)a.re * b.re + (double)a.im * b.im;
      ai += (double)a.re * b.im - (double)a.im * b.re;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
    if (lane == 0) { atomicAdd(hb + 2 * k, ar); atomicAdd(hb + 2 * k + 1, ai); }
  }
}

// w -= sum_{k<=j} h_k V_k, ||w||^2 (double atomic per CTA); the LAST CTA to finish runs the Givens / convergence step of
// the iteration and clears the accumulation buffers for the next one (vector_PRECISION_multi_saxpy + global_norm + qr_update,
// linsolve_generic.c:859-940).  No CTA waits for another one.
__global__ void __launch_bounds__(256)
k_gmres_axpy_givens(cf *__restrict__ w, const cf *__restrict__ V, long stride, int j, long nelem, double *S, int *ct, GmresOff o,
                    double tl, unsigned *counter) {
  if (ct[0]) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *hs = reinterpret_cast<cf *>(smem_raw);
  __shared__ double red[8];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k <= j; k += 256) hs[k] = cf((float)S[o.HB + 2 * k], (float)S[o.HB + 2 * k + 1]);
  __syncthreads();
  const long e = (long)blockIdx.x * 256 + tid;
  double loc = 0.0;
  if (e < nelem) {
    cf v = w[e];
    for (int k = 0; k <= j; k++) fms_(v, hs[k], V[(long)k * stride + e]);
    w[e] = v;
    loc = (double)v.re * v.re + (double)v.im * v.im;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, s);
  if (lane == 0) red[warp] = loc;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; i++) tot += red[i];
    atomicAdd(&S[o.N], tot);
    __threadfence();
    ticket = atomicAdd(counter, 1u);
  }
  __syncthreads();
  if (ticket == gridDim.x - 1 && tid == 0) {
    __threadfence();
    const double nv = atomicAdd(&S[o.N], 0.0);        // the coherent total of every CTA's contribution
    S[o.N] = nv;
    gmres_givens(S, ct, o, j, tl);
    for (int i = 0; i < 2 * (j + 2); i++) S[o.HB + i] = 0.0;
    S[o.N] = 0.0;
    *counter = 0u;
  }
}

namespace {
int pick_groups(int n) {
  int G = 128 / (n / 2);
  while (G > 1 && n % (2 * G) != 0) G--;
  return (n % (2 * G) == 0) ? G : 0;
}
int sm_count() {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  return sms;
}
const int SCHUR_STAGES = 3;
int g_schur_reverse = -1;
}  // namespace

bool schur_fast_supported(const CoarseOp &op) {
  const int n = op.n;
  if (n > 64 || n < 8 || (n & 3) || op.V <= 0 || op.n_even <= 0 || op.n_even >= op.V || !op.Sinv) return false;
  return pick_groups(n) > 0;
}

// one half application: forward sites get dir, daggered sites fill Z (see k_schur_hop)
void schur_hop(const CoarseOp &op, int phase, const cf *in, const cf *self, cf *dir, cf *Z, const int *skip) {
  if (g_schur_reverse < 0) { const char *e = getenv("DDA_SCHUR_REVERSE"); g_schur_reverse = e ? atoi(e) : 1; }
  const int n = op.n, G = pick_groups(n);
  const size_t nn = (size_t)n * n;
  const size_t smem = SCHUR_STAGES * nn * sizeof(cf) + (5 + 4 * G) * n * sizeof(cf) + 8 * sizeof(uint64_t);
  static size_t attr = 0;
  if (smem > attr) { CUDA_CHECK(cudaFuncSetAttribute(k_schur_hop<SCHUR_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  DDA_ASSERT(per_sm >= 1);
  const long grid = std::min<long>(op.V, (long)sm_count() * per_sm);
  k_schur_hop<SCHUR_STAGES><<<(unsigned)grid, 128, smem, g_stream>>>(op, phase, in, self, dir, Z, (int)op.V, G, g_schur_reverse ? phase : 0, skip);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

void schur_mid(const CoarseOp &op, const cf *eta, const cf *dir, const cf *Z, cf *out, float a, float b, float cS, const int *skip) {
  const int n = op.n, G = pick_groups(n);
  const size_t nn = (size_t)n * n;
  const size_t smem = 2 * nn * sizeof(cf) + (1 + G) * n * sizeof(cf) + 8 * sizeof(uint64_t);
  static size_t attr = 0;
  if (smem > attr) { CUDA_CHECK(cudaFuncSetAttribute(k_schur_mid<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  const long nodd = op.V - op.n_even;
  const long grid = std::min<long>(nodd, (long)sm_count() * per_sm);
  k_schur_mid<2><<<(unsigned)grid, 128, smem, g_stream>>>(op, eta, dir, Z, out, a, b, cS, G, skip);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

void schur_fin(const CoarseOp &op, const cf *eta, const cf *dir, const cf *Z, cf *out, float a, float b, const int *skip) {
  const long total = op.n_even * op.n;
  k_schur_fin<<<(unsigned)((total + 127) / 128), 128, 0, g_stream>>>(op, eta, dir, Z, out, a, b, skip);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

void schur_fin_dots(const CoarseOp &op, const cf *dir, const cf *Z, cf *w, const cf *V, long stride, int j, double *hb, const int *skip) {
  const int spc = 4;                                             // sites per CTA: 216 CTAs on the 8 x 6^3 lattice
  const size_t smem = (size_t)spc * op.n * sizeof(cf);
  const long grid = (op.n_even + spc - 1) / spc;
  k_schur_fin_dots<<<(unsigned)grid, 256, smem, g_stream>>>(op, dir, Z, w, V, stride, j, hb, spc, skip);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

void gmres_axpy_givens(cf *w, const cf *V, long stride, int j, long nelem, double *S, int *ct, const GmresOff &o, double tol, unsigned *counter) {
  const size_t smem = (size_t)(o.m + 1) * sizeof(cf);
  DDA_ASSERT(smem <= 40 * 1024);
  k_gmres_axpy_givens<<<(unsigned)((nelem + 255) / 256), 256, smem, g_stream>>>(w, V, stride, j, nelem, S, ct, o, tol, counter);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

#endif

}  // namespace dda

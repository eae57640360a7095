// sap_kernel.cu -- fused fine-level SAP block solve for sm_100a: ONE CTA per Schwarz block, one launch per colour.
//
// Reference counterparts: red_black_schwarz_PRECISION (schwarz_generic.c:1260-1431), block_solve_oddeven_PRECISION
// (oddeven_generic.c:1332-1360), apply_block_schur_complement (:1317-1329), block_hopping_term / block_n_hopping_term
// (:1051-1314), block_diag_ee / block_diag_oo_inv (:975-1046), local_minres_PRECISION (linsolve_generic.c:985-1029),
// block_PRECISION_boundary_op (schwarz_generic.c:743-856).
//
// Per block visit the reference runs ~14 separate loops over the block's sites; here the whole visit
//   r = eta - D x (block rows, couplings to neighbouring blocks included)
//   e_o = Doo^-1 r_o ; t_e = r_e - Deo e_o
//   block_iter x minimal residual on the even-odd Schur complement S = Dee - Deo Doo^-1 Doe
//   e_o = Doo^-1 (r_o - Doe e_e) ; x += e
// is one kernel.  The block's links (73.7 KB in float) stay in shared memory for the whole visit; the iteration
// vectors stay in registers; only the vector that the opposite parity has to gather goes through a 24.6 KB shared
// buffer.  Work decomposition: a PAIR of threads (lanes l and l^16) owns one even and one odd site of the block;
// thread s of the pair owns spin components {s, 2+s} of both sites, i.e. half of every half-spinor projection,
// SU(3) multiply and clover row block -- same instruction stream for both (the gamma-matrix signs that differ
// between the halves are a per-thread +-1 factor), halves recombined with 6 warp shuffles per operator.  This gives
// 8 warps per block visit and 16 warps per SM at <= 128 registers (the one-thread-per-site-pair variant ran at 7
// warps/SM and 21 % issue utilisation, see profiles/).  Clover blocks (even sites) and their inverses (odd sites) are
// streamed from L2.  MR inner products: warp shuffle + one shared-memory stage.  Shared memory: links 73.7 KB + exchange
// buffer 24.6 KB + parked r_o 12.3 KB = 110.9 KB -> 2 CTAs/SM.
#include "solver.h"
#include "fine_op.cuh"
#include "halo.h"
#include "tma.cuh"

namespace dda {

#ifndef DDA_HOST_EMU

namespace sap {

const int BS = 256;                 // sites per block (4^4), BS threads per CTA
struct Shared {
  float2 U[36][BS];                 // links of the block's sites, [9*mu + 3*row + col][site]
  float2 Vb[12][BS];                // exchange buffer [component][site in block]
  float red[2][BS / 32][4];
  float2 stash[6][BS];               // per-thread copy of r_o (needed only at the first and the last step of the block solve)
};

struct Ctx {                        // per-thread constants of the pair decomposition
  int s;                            // which half: owns spins s and 2+s
  float sg;                         // s ? +1 : -1
  int up, loA, loB;                 // component offsets: 3s, 6+3s (own lower spin), 6+3(1-s) (partner's lower spin)
};

__device__ __forceinline__ cf ld2(const float2 &v) { return cf(v.x, v.y); }

// ------------------------------------------------------------------------------------------------------------------
// CODE SIZE matters here: with every operator application inlined at its call site the visit is ~17 k instructions
// (270 KB) and the kernel stalls on instruction fetch (ncu: 55 % of the samples "no instruction",
// profiles/r1_ncu_full_k_sap_fine_a.txt).  The visit is therefore written as loops in which each operator (in-block
// hops for the odd / even site, clover, clover inverse) appears exactly ONCE: k_sap_fine's main loop below, and a
// two-pass loop with register rotation for the residual of the pair's two sites.
//
// Half hop for the thread owning spins {s, 2+s} (sg = s ? 1 : -1), S = +1 forward / -1 backward:
//   h = pu - S u pl ,  g = W h ,  up -= g ,  low += S u' g ,   W = D_mu(x) or D_mu(x-mu)^dagger
//   T: u = -1,    pl = spin 2+s, u' = -1     -> lowA        Z: u = -i,   pl = spin 3-s, u' = +i     -> lowB
//   X: u = sg i,  pl = spin 2+s, u' = -sg i  -> lowA        Y: u = sg,   pl = spin 3-s, u' = sg     -> lowB
// (gamma basis BASIS0, clifford.h:39-100; same arithmetic as project / reconstruct_sub in fine_op.cuh).  lowA is the
// thread's own lower spin, lowB the partner's (handed over by shuffle in pair_combine).

// one half of a hop with compile-time direction MU and sign S (+1 forward, -1 backward)
template <int MU, int S>
__device__ __forceinline__ void half_hop(const cf *pu, const cf *pl, const cf *M, float sg, cf *up, cf *lowA, cf *lowB) {
  cf h[3], g[3];
  const float t = (S > 0) ? sg : -sg;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    if (MU == 0) h[c] = (S > 0) ? pu[c] + pl[c] : pu[c] - pl[c];
    else if (MU == 1) h[c] = (S > 0) ? cf(pu[c].re - pl[c].im, pu[c].im + pl[c].re) : cf(pu[c].re + pl[c].im, pu[c].im - pl[c].re);
    else if (MU == 2) h[c] = cf(pu[c].re - t * pl[c].re, pu[c].im - t * pl[c].im);
    else h[c] = cf(pu[c].re + t * pl[c].im, pu[c].im - t * pl[c].re);
  }
#pragma unroll
  for (int r = 0; r < 3; r++) {
    cf a;
    if (S > 0) { a = M[3 * r] * h[0]; fma_(a, M[3 * r + 1], h[1]); fma_(a, M[3 * r + 2], h[2]); }
    else { a = conj(M[r]) * h[0]; fmac_(a, M[3 + r], h[1]); fmac_(a, M[6 + r], h[2]); }
    g[r] = a;
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    up[c] -= g[c];
    if (MU == 0) { if (S > 0) lowA[c] -= g[c]; else lowA[c] += g[c]; }
    else if (MU == 1) { if (S > 0) { lowB[c].re -= g[c].im; lowB[c].im += g[c].re; } else { lowB[c].re += g[c].im; lowB[c].im -= g[c].re; } }
    else if (MU == 2) { lowB[c].re += t * g[c].re; lowB[c].im += t * g[c].im; }
    else { lowA[c].re += t * g[c].im; lowA[c].im -= t * g[c].re; }
  }
}

// in-block hop pair of direction MU, operands in shared memory (gather form); vu / vl: the thread's component rows
template <int MU>
__device__ __forceinline__ void sm_hop_pair(const Shared &sm, const float2 *vu, const float2 *vl, float sg, int l, unsigned inmask,
                                            unsigned nbf, unsigned nbb, cf *up, cf *lowA, cf *lowB) {
  if (inmask & (1u << MU)) {
    const int n = (nbf >> (8 * MU)) & 0xFF;
    cf pu[3], pl[3], M[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { pu[c] = ld2(vu[c * BS + n]); pl[c] = ld2(vl[c * BS + n]); }
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = ld2(sm.U[9 * MU + k][l]);
    half_hop<MU, +1>(pu, pl, M, sg, up, lowA, lowB);
  }
  if (inmask & (1u << (4 + MU))) {
    const int n = (nbb >> (8 * MU)) & 0xFF;
    cf pu[3], pl[3], M[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { pu[c] = ld2(vu[c * BS + n]); pl[c] = ld2(vl[c * BS + n]); }
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = ld2(sm.U[9 * MU + k][n]);
    half_hop<MU, -1>(pu, pl, M, sg, up, lowA, lowB);
  }
}

// hop pair to neighbouring blocks, operands in global memory (tiled layout), only directions flagged in `mask`
template <int MU>
__device__ __forceinline__ void gl_hop_pair(const FineOp<float> &op, const Ctx &cx_, long s, unsigned mask, const cf *x,
                                            cf *up, cf *lowA, cf *lowB) {
  const int lo = (MU == 0 || MU == 3) ? cx_.loA : cx_.loB;
  if (mask & (1u << MU)) {
    const long n = op.nb[(long)MU * op.V + s];
    const long nv = (n >> 5) * (12L << 5) + (n & 31), su = (s >> 5) * (36L << 5) + (s & 31);
    cf pu[3], pl[3], M[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { pu[c] = x[nv + ((long)(cx_.up + c) << 5)]; pl[c] = x[nv + ((long)(lo + c) << 5)]; }
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = op.D[su + ((long)(9 * MU + k) << 5)];
    half_hop<MU, +1>(pu, pl, M, cx_.sg, up, lowA, lowB);
  }
  if (mask & (1u << (4 + MU))) {
    const long n = op.nb[(long)(4 + MU) * op.V + s];
    const long nv = (n >> 5) * (12L << 5) + (n & 31), nu = (n >> 5) * (36L << 5) + (n & 31);
    cf pu[3], pl[3], M[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { pu[c] = x[nv + ((long)(cx_.up + c) << 5)]; pl[c] = x[nv + ((long)(lo + c) << 5)]; }
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = op.D[nu + ((long)(9 * MU + k) << 5)];
    half_hop<MU, -1>(pu, pl, M, cx_.sg, up, lowA, lowB);
  }
}

// combine the pair's partial results: y[0..2] = spin s, y[3..5] = spin 2+s of N v (own lowA + partner's lowB)
__device__ __forceinline__ void pair_combine(const cf *up, const cf *lowA, const cf *lowB, cf *y) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float pr = __shfl_xor_sync(0xffffffffu, lowB[c].re, 16), pi = __shfl_xor_sync(0xffffffffu, lowB[c].im, 16);
    y[c] = up[c];
    y[3 + c] = cf(lowA[c].re + pr, lowA[c].im + pi);
  }
}

// y = N v for site l of the block (in-block hops only); v is in the shared exchange buffer
__device__ __forceinline__ void sm_hops(const Shared &sm, const Ctx &cx_, int l, unsigned inmask, unsigned nbf, unsigned nbb, cf *y) {
  cf up[3], lowA[3], lowB[3];
#pragma unroll
  for (int c = 0; c < 3; c++) { up[c] = cf(0.f, 0.f); lowA[c] = cf(0.f, 0.f); lowB[c] = cf(0.f, 0.f); }
  const float2 *vu = &sm.Vb[cx_.up][0], *va = &sm.Vb[cx_.loA][0], *vb = &sm.Vb[cx_.loB][0];
  sm_hop_pair<0>(sm, vu, va, cx_.sg, l, inmask, nbf, nbb, up, lowA, lowB);
  sm_hop_pair<1>(sm, vu, vb, cx_.sg, l, inmask, nbf, nbb, up, lowA, lowB);
  sm_hop_pair<2>(sm, vu, vb, cx_.sg, l, inmask, nbf, nbb, up, lowA, lowB);
  sm_hop_pair<3>(sm, vu, va, cx_.sg, l, inmask, nbf, nbb, up, lowA, lowB);
  pair_combine(up, lowA, lowB, y);
}

// y = N x restricted to the hops that leave the block (flags in `mask`), operands in global memory
__device__ __forceinline__ void gl_hops(const FineOp<float> &op, const Ctx &cx_, long s, unsigned mask, const cf *x, cf *y) {
  cf up[3], lowA[3], lowB[3];
#pragma unroll
  for (int c = 0; c < 3; c++) { up[c] = cf(0.f, 0.f); lowA[c] = cf(0.f, 0.f); lowB[c] = cf(0.f, 0.f); }
  if (mask) {
    gl_hop_pair<0>(op, cx_, s, mask, x, up, lowA, lowB); gl_hop_pair<1>(op, cx_, s, mask, x, up, lowA, lowB);
    gl_hop_pair<2>(op, cx_, s, mask, x, up, lowA, lowB); gl_hop_pair<3>(op, cx_, s, mask, x, up, lowA, lowB);
  }
  pair_combine(up, lowA, lowB, y);
}

__device__ __forceinline__ void put(Shared &sm, const Ctx &cx_, int l, const cf *v) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    sm.Vb[cx_.up + c][l] = make_float2(v[c].re, v[c].im);
    sm.Vb[cx_.loA + c][l] = make_float2(v[3 + c].re, v[3 + c].im);
  }
}

// index of the packed upper-triangle entry (r, c), r < c, of a Hermitian 6x6 block (row-major over r < c)
__host__ __device__ constexpr int tri(int r, int c) { return r * 6 - r * (r + 1) / 2 + (c - r - 1); }

// own half of y = M x for the packed Hermitian 2 x (6x6) site matrix at Cs (pointer already offset by tile and lane):
// x, y hold the own components (spin s, spin 2+s); the partner's components come by shuffle.  Thread s computes the
// row block s of each 6x6 block: A (rows/cols of spin s, Hermitian 3x3) times the own part plus B (s = 0) / B^H (s = 1)
// times the partner's part.  The two 6x6 blocks are a run-time loop (code size).
__device__ __forceinline__ void clov_half(const float *__restrict__ Cs, const Ctx &cx_, const cf *x, cf *y) {
  cf xm[3], xo[3], xmn[3], xon[3], ylo[3], yhi[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    xm[c] = x[c]; xmn[c] = x[3 + c];
    xo[c] = cf(__shfl_xor_sync(0xffffffffu, x[c].re, 16), __shfl_xor_sync(0xffffffffu, x[c].im, 16));
    xon[c] = cf(__shfl_xor_sync(0xffffffffu, x[3 + c].re, 16), __shfl_xor_sync(0xffffffffu, x[3 + c].im, 16));
    ylo[c] = cf(0.f, 0.f); yhi[c] = cf(0.f, 0.f);
  }
  const int s = cx_.s;
  const float cj = -cx_.sg;          // +1 for s = 0 (entries used as stored), -1 for s = 1 (conjugated)
#pragma unroll 1
  for (int b = 0; b < 2; b++) {
    const float *Cd = Cs + ((6 * b + 3 * s) << 5), *Ct = Cs + ((12 + 30 * b) << 5);
    cf yo[3];
#pragma unroll
    for (int i = 0; i < 3; i++) yo[i] = __ldg(Cd + (i << 5)) * xm[i];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i + 1; j < 3; j++) {
        const int m = s ? tri(3 + i, 3 + j) : tri(i, j);
        const cf a(__ldg(Ct + ((2 * m) << 5)), __ldg(Ct + ((2 * m + 1) << 5)));
        fma_(yo[i], a, xm[j]);
        fmac_(yo[j], a, xm[i]);
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int m = s ? tri(j, 3 + i) : tri(i, 3 + j);
        const cf a(__ldg(Ct + ((2 * m) << 5)), cj * __ldg(Ct + ((2 * m + 1) << 5)));
        fma_(yo[i], a, xo[j]);
      }
#pragma unroll
    for (int c = 0; c < 3; c++) { ylo[c] = yhi[c]; yhi[c] = yo[c]; xm[c] = xmn[c]; xo[c] = xon[c]; }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) { y[c] = ylo[c]; y[3 + c] = yhi[c]; }
}

__device__ __forceinline__ void load_own(const cf *v, long base, const Ctx &cx_, cf *out) {
#pragma unroll
  for (int c = 0; c < 3; c++) { out[c] = v[base + ((long)(cx_.up + c) << 5)]; out[3 + c] = v[base + ((long)(cx_.loA + c) << 5)]; }
}

struct SiteRef {       // everything that identifies one of the thread pair's two sites
  int l; long s, v; const float *C; unsigned f, in, nf, nb;
};

// first_zero: x == 0 on the whole lattice on entry (first colour of a zero-guess call): r = eta, x = e.
__global__ void __launch_bounds__(BS, 2)
k_sap_fine(FineOp<float> op, cf *x, const cf *__restrict__ eta, const int *__restrict__ blocklist, int biter, int first_zero) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Shared &sm = *reinterpret_cast<Shared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Ctx cx_;
  cx_.s = lane >> 4; cx_.sg = cx_.s ? 1.f : -1.f;
  cx_.up = 3 * cx_.s; cx_.loA = 6 + 3 * cx_.s; cx_.loB = 6 + 3 * (1 - cx_.s);
  const long base = (long)blocklist[blockIdx.x] * BS;
  const long V = op.V;

  // links of the block -> shared memory: thread tid copies site tid (coalesced 256 B rows)
  {
    const float2 *D2 = reinterpret_cast<const float2 *>(op.D);
    const long st = base + tid;
    const long u = (st >> 5) * (36L << 5) + (st & 31);
#pragma unroll 4
    for (int k = 0; k < 36; k++) sm.U[k][tid] = __ldg(D2 + u + ((long)k << 5));
  }
  // the pair's even site E and odd site O: block-local index, global index, offsets, in-block neighbour indices
  SiteRef E, O;
  E.l = 16 * w + (lane & 15); O.l = BS / 2 + E.l;
  {
    SiteRef *sr[2] = {&E, &O};
#pragma unroll
    for (int p = 0; p < 2; p++) {
      SiteRef &R = *sr[p];
      R.s = base + R.l;
      R.v = (R.s >> 5) * (12L << 5) + (R.s & 31);
      R.C = op.C + (R.s >> 5) * (72L << 5) + (R.s & 31);
      R.f = op.blkflag[R.s]; R.in = (~R.f) & 0xFFu; R.nf = 0; R.nb = 0;
#pragma unroll
      for (int d = 0; d < 4; d++) {
        R.nf |= (unsigned)((__ldg(op.nb + (long)d * V + R.s) - base) & 0xFF) << (8 * d);
        R.nb |= (unsigned)((__ldg(op.nb + (long)(4 + d) * V + R.s) - base) & 0xFF) << (8 * d);
      }
    }
  }
  const float *CinvO = op.Cinv + (O.s >> 5) * (72L << 5) + (O.s & 31);

  cf rE[6], rO[6];
  load_own(eta, E.v, cx_, rE); load_own(eta, O.v, cx_, rO);
  if (!first_zero) {
    // r = eta - D x on the block: clover + couplings to neighbouring blocks (global) + in-block hops (shared memory);
    // the two sites of the pair run through ONE copy of the code (register rotation at the loop end)
    SiteRef A = E, B = O;
    cf ra[6], rb[6];
#pragma unroll
    for (int c = 0; c < 6; c++) { ra[c] = rE[c]; rb[c] = rO[c]; }
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
      cf xa[6], y[6];
      load_own(x, A.v, cx_, xa);
      put(sm, cx_, A.l, xa);
      clov_half(A.C, cx_, xa, y);
#pragma unroll
      for (int c = 0; c < 6; c++) ra[c] -= y[c];
      gl_hops(op, cx_, A.s, A.f, x, y);
#pragma unroll
      for (int c = 0; c < 6; c++) { ra[c] -= y[c]; const cf t = ra[c]; ra[c] = rb[c]; rb[c] = t; }
      const SiteRef T = A; A = B; B = T;
    }
    __syncthreads();
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
      cf y[6];
      sm_hops(sm, cx_, A.l, A.in, A.nf, A.nb, y);
#pragma unroll
      for (int c = 0; c < 6; c++) { ra[c] -= y[c]; const cf t = ra[c]; ra[c] = rb[c]; rb[c] = t; }
      const SiteRef T = A; A = B; B = T;
    }
#pragma unroll
    for (int c = 0; c < 6; c++) { rE[c] = ra[c]; rO[c] = rb[c]; }
    __syncthreads();
  }

  // block solve.  k = 0: e_o = Coo^-1 r_o, t_e = r_e - N_eo e_o.   k = 1..biter: one MR step on the Schur complement
  // S = C_ee - N_eo Coo^-1 N_oe.   k = biter+1: back substitution e_o = Coo^-1 (r_o - N_oe e_e), x += e.
  // Register budget (128): r_e lives on only as the initial value of t_e, r_o is parked in shared memory.
  cf tE[6], eE[6];
#pragma unroll
  for (int c = 0; c < 6; c++) { tE[c] = rE[c]; eE[c] = cf(0.f, 0.f); sm.stash[c][tid] = make_float2(rO[c].re, rO[c].im); }
#pragma unroll 1
  for (int k = 0; k <= biter + 1; k++) {
    cf wv[6], z[6];
    if (k > 0) sm_hops(sm, cx_, O.l, O.in, O.nf, O.nb, wv);          // N_oe (t_e or e_e)
    if (k == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) wv[c] = ld2(sm.stash[c][tid]);
    } else if (k == biter + 1) {
#pragma unroll
      for (int c = 0; c < 6; c++) wv[c] = ld2(sm.stash[c][tid]) - wv[c];
    }
    clov_half(CinvO, cx_, wv, z);
    if (k == biter + 1) {
      // x += e
      if (!first_zero) {
        cf a[6], b[6];
        load_own(x, E.v, cx_, a); load_own(x, O.v, cx_, b);
#pragma unroll
        for (int c = 0; c < 6; c++) { eE[c] += a[c]; z[c] += b[c]; }
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        x[E.v + ((long)(cx_.up + c) << 5)] = eE[c]; x[E.v + ((long)(cx_.loA + c) << 5)] = eE[3 + c];
        x[O.v + ((long)(cx_.up + c) << 5)] = z[c]; x[O.v + ((long)(cx_.loA + c) << 5)] = z[3 + c];
      }
      break;
    }
    if (k > 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) z[c] = -z[c];                        // a2_o = -Coo^-1 N_oe t_e
    }
    put(sm, cx_, O.l, z);
    __syncthreads();
    cf y[6];
    sm_hops(sm, cx_, E.l, E.in, E.nf, E.nb, y);                      // N_eo (e_o or a2_o)
    if (k == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) tE[c] -= y[c];
    } else {
      cf Dr[6];
      clov_half(E.C, cx_, tE, Dr);
#pragma unroll
      for (int c = 0; c < 6; c++) Dr[c] += y[c];
      // alpha = <Dr,t>/<Dr,Dr> over the even sites of the block (local_minres, linsolve_generic.c:1013-1022)
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        p0 += Dr[c].re * tE[c].re + Dr[c].im * tE[c].im;
        p1 += Dr[c].re * tE[c].im - Dr[c].im * tE[c].re;
        p2 += Dr[c].re * Dr[c].re + Dr[c].im * Dr[c].im;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
      }
      float (*red)[4] = sm.red[k & 1];
      if (lane == 0) { red[w][0] = p0; red[w][1] = p1; red[w][2] = p2; }
      __syncthreads();
      p0 = p1 = p2 = 0.f;
#pragma unroll
      for (int i = 0; i < BS / 32; i++) { p0 += red[i][0]; p1 += red[i][1]; p2 += red[i][2]; }
      cf alpha(0.f, 0.f);
      if (p2 > 1e-30f) alpha = cf(p0 / p2, p1 / p2);
#pragma unroll
      for (int c = 0; c < 6; c++) { fma_(eE[c], alpha, tE[c]); fms_(tE[c], alpha, Dr[c]); }
    }
    put(sm, cx_, E.l, (k < biter) ? tE : eE);
    __syncthreads();
  }
}


// ===================================================================================================================
// Version 2 of the fused block visit (default; DDA_SAP_V1=1 selects k_sap_fine above).
//
// What changed against v1 (ncu of v1: 52 % issue utilisation, 150 k warp instructions per visit, top stall = L2 latency
// of the clover stream, 377 KB DRAM per visit):
//  * the clover blocks of the even sites and the inverse blocks of the odd sites are STAGED IN SHARED MEMORY by one TMA
//    bulk copy each (the 128 sites of one parity are 4 contiguous 32-site tiles = 36 864 B), issued as soon as the
//    previous contents are consumed, so the copy engine fetches the next operand block while the in-block hops run;
//  * room for that buffer: only the 768 in-block links are kept in shared memory (a 4^4 block has 1024 links, the 256
//    that leave the block are only needed once, by the residual, and are read from global memory there), and the exchange
//    buffer holds one parity at a time;
//  * the clover term of the odd sites is never applied: with r'_o = eta_o - (N x)_o the block solve needs only
//    Coo^-1 r'_o - x_o and Coo^-1 (r'_o - N_oe e_e)  (algebraically the same iterates), so C_oo is not read at all;
//  * r'_o is parked in the odd sites' slots of x itself (dead until the final write) instead of shared memory;
//  * all complex multiply-adds are 4 FFMA (common.cuh fma_).
// Shared memory: links 55 296 + exchange 12 288 + clover 36 864 + reductions 256 + barrier 8 = 104 712 B -> 2 CTAs / SM.
struct Shared2 {
  float Cb[4 * 72 * 32];            // clover (even sites) or inverse clover (odd sites) of the block, tile layout
  float2 U[4][9][192];              // in-block links: [mu][3*row+col][slot of the link's source site]
  float2 Vb[12][BS / 2];            // exchange buffer [component][site of ONE parity]
  float red[2][BS / 32][4];
  unsigned long long bar, barU;
};

struct Site2 {                      // one of the thread pair's two sites; packed bytes, one per direction
  int l;                            // block-local index
  unsigned in;                      // in-block mask (bit d: neighbour d is inside the block)
  unsigned nf, nb;                  // exchange-buffer index (0..127) of the +mu / -mu neighbour
  unsigned lf, lb;                  // link slot of U_mu(site) / U_mu(site - mu)
};

template <int MU>
__device__ __forceinline__ void sm_hop_pair2(const Shared2 &sm, const float2 *vu, const float2 *vl, float sg, const Site2 &R,
                                             cf *up, cf *lowA, cf *lowB) {
  if (R.in & (1u << MU)) {
    const int n = (R.nf >> (8 * MU)) & 0xFF, u = (R.lf >> (8 * MU)) & 0xFF;
    cf pu[3], pl[3], M[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { pu[c] = ld2(vu[c * (BS / 2) + n]); pl[c] = ld2(vl[c * (BS / 2) + n]); }
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = ld2(sm.U[MU][k][u]);
    half_hop<MU, +1>(pu, pl, M, sg, up, lowA, lowB);
  }
  if (R.in & (1u << (4 + MU))) {
    const int n = (R.nb >> (8 * MU)) & 0xFF, u = (R.lb >> (8 * MU)) & 0xFF;
    cf pu[3], pl[3], M[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { pu[c] = ld2(vu[c * (BS / 2) + n]); pl[c] = ld2(vl[c * (BS / 2) + n]); }
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = ld2(sm.U[MU][k][u]);
    half_hop<MU, -1>(pu, pl, M, sg, up, lowA, lowB);
  }
}

// y = N v for site R (in-block hops only); v (opposite parity) is in the exchange buffer
__device__ __forceinline__ void sm_hops2(const Shared2 &sm, const Ctx &cx_, const Site2 &R, cf *y) {
  cf up[3], lowA[3], lowB[3];
#pragma unroll
  for (int c = 0; c < 3; c++) { up[c] = cf(0.f, 0.f); lowA[c] = cf(0.f, 0.f); lowB[c] = cf(0.f, 0.f); }
  const float2 *vu = &sm.Vb[cx_.up][0], *va = &sm.Vb[cx_.loA][0], *vb = &sm.Vb[cx_.loB][0];
  sm_hop_pair2<0>(sm, vu, va, cx_.sg, R, up, lowA, lowB);
  sm_hop_pair2<1>(sm, vu, vb, cx_.sg, R, up, lowA, lowB);
  sm_hop_pair2<2>(sm, vu, vb, cx_.sg, R, up, lowA, lowB);
  sm_hop_pair2<3>(sm, vu, va, cx_.sg, R, up, lowA, lowB);
  pair_combine(up, lowA, lowB, y);
}

__device__ __forceinline__ void put2(Shared2 &sm, const Ctx &cx_, int h, const cf *v) {   // h: index inside the parity (0..127)
#pragma unroll
  for (int c = 0; c < 3; c++) {
    sm.Vb[cx_.up + c][h] = make_float2(v[c].re, v[c].im);
    sm.Vb[cx_.loA + c][h] = make_float2(v[3 + c].re, v[3 + c].im);
  }
}

// clov_half with the packed blocks in shared memory (Cs already offset by tile and lane, stride 32 floats per entry)
__device__ __forceinline__ void clov_half_sm(const float *Cs, const Ctx &cx_, const cf *x, cf *y) {
  cf xm[3], xo[3], xmn[3], xon[3], ylo[3], yhi[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    xm[c] = x[c]; xmn[c] = x[3 + c];
    xo[c] = cf(__shfl_xor_sync(0xffffffffu, x[c].re, 16), __shfl_xor_sync(0xffffffffu, x[c].im, 16));
    xon[c] = cf(__shfl_xor_sync(0xffffffffu, x[3 + c].re, 16), __shfl_xor_sync(0xffffffffu, x[3 + c].im, 16));
    ylo[c] = cf(0.f, 0.f); yhi[c] = cf(0.f, 0.f);
  }
  const int s = cx_.s;
  const float cj = -cx_.sg;          // +1 for s = 0 (entries used as stored), -1 for s = 1 (conjugated)
#pragma unroll 1
  for (int b = 0; b < 2; b++) {
    const float *Cd = Cs + ((6 * b + 3 * s) << 5), *Ct = Cs + ((12 + 30 * b) << 5);
    cf yo[3];
#pragma unroll
    for (int i = 0; i < 3; i++) yo[i] = Cd[i << 5] * xm[i];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i + 1; j < 3; j++) {
        const int m = s ? tri(3 + i, 3 + j) : tri(i, j);
        const cf a(Ct[(2 * m) << 5], Ct[(2 * m + 1) << 5]);
        fma_(yo[i], a, xm[j]);
        fmac_(yo[j], a, xm[i]);
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int m = s ? tri(j, 3 + i) : tri(i, 3 + j);
        const cf a(Ct[(2 * m) << 5], cj * Ct[(2 * m + 1) << 5]);
        fma_(yo[i], a, xo[j]);
      }
#pragma unroll
    for (int c = 0; c < 3; c++) { ylo[c] = yhi[c]; yhi[c] = yo[c]; xm[c] = xmn[c]; xo[c] = xon[c]; }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) { y[c] = ylo[c]; y[3 + c] = yhi[c]; }
}

__device__ __forceinline__ void store_own(cf *v, long base, const Ctx &cx_, const cf *val) {
#pragma unroll
  for (int c = 0; c < 3; c++) { v[base + ((long)(cx_.up + c) << 5)] = val[c]; v[base + ((long)(cx_.loA + c) << 5)] = val[3 + c]; }
}

// tab: per block-local site 5 words {in, nf, nb, lf, lb} (identical for every block, built on the host)
__global__ void __launch_bounds__(BS, 2)
k_sap_fine2(FineOp<float> op, const cf *__restrict__ Dblk, cf *x, const cf *__restrict__ eta, const int *__restrict__ blocklist,
            const unsigned *__restrict__ tab, int biter, int first_zero, int ext) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Shared2 &sm = *reinterpret_cast<Shared2 *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Ctx cx_;
  cx_.s = lane >> 4; cx_.sg = cx_.s ? 1.f : -1.f;
  cx_.up = 3 * cx_.s; cx_.loA = 6 + 3 * cx_.s; cx_.loB = 6 + 3 * (1 - cx_.s);
  const long base = (long)blocklist[blockIdx.x] * BS;
  uint64_t *bar = reinterpret_cast<uint64_t *>(&sm.bar);
  const uint32_t CB_BYTES = 4 * 72 * 32 * sizeof(float);
  // clover stream: load number q (0, 1, 2, ...) completes phase q of the barrier
  const float *srcCe = op.C + (base >> 5) * (72L << 5);                 // even sites: first 4 tiles of the block
  const float *srcCinvO = op.Cinv + ((base + BS / 2) >> 5) * (72L << 5);  // odd sites: last 4 tiles
  uint32_t cphase = 0;
  auto cload = [&](const float *src) {                                    // thread 0 only, after a block-wide barrier
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, CB_BYTES);
    tma_bulk_g2s(sm.Cb, src, CB_BYTES, bar);
  };
  auto cwait = [&]() { mbar_wait(bar, cphase & 1); cphase++; };
  uint64_t *barU = reinterpret_cast<uint64_t *>(&sm.barU);
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(barU, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cload(first_zero ? srcCinvO : srcCe);
    // the 768 in-block links of the block: one bulk copy of the per-block image (solver_refresh_float_op)
    const uint32_t UB = (uint32_t)sizeof(sm.U);
    mbar_expect_tx(barU, UB);
    tma_bulk_g2s(&sm.U[0][0][0], Dblk + (base / BS) * (4 * 9 * 192), UB, barU);
  }
  Site2 E, O;
  E.l = 16 * w + (lane & 15); O.l = BS / 2 + E.l;
  {
    const unsigned *te = tab + 5 * E.l, *to = tab + 5 * O.l;
    E.in = __ldg(te); E.nf = __ldg(te + 1); E.nb = __ldg(te + 2); E.lf = __ldg(te + 3); E.lb = __ldg(te + 4);
    O.in = __ldg(to); O.nf = __ldg(to + 1); O.nb = __ldg(to + 2); O.lf = __ldg(to + 3); O.lb = __ldg(to + 4);
  }
  const long sE = base + E.l, sO = base + O.l;
  const long vE = (sE >> 5) * (12L << 5) + (sE & 31), vO = (sO >> 5) * (12L << 5) + (sO & 31);
  const float *CsT = sm.Cb + (E.l >> 5) * (72 * 32) + (E.l & 31);       // same tile / lane for the even and the odd site

  cf rE[6], rO[6];
  load_own(eta, vE, cx_, rE); load_own(eta, vO, cx_, rO);
  __syncthreads();                                                      // barrier initialisation visible
  bool links_ready = false;
  if (!first_zero) {
    // r_e = eta_e - C_e x_e - (N x)_e ,  r'_o = eta_o - (N x)_o : couplings to neighbouring blocks from global memory,
    // in-block hops through the exchange buffer, one parity at a time
    cf xe[6], y[6];
    load_own(x, vE, cx_, xe);
    cwait();
    clov_half_sm(CsT, cx_, xe, y);
#pragma unroll
    for (int c = 0; c < 6; c++) rE[c] -= y[c];
    put2(sm, cx_, E.l, xe);
    if (!ext) {
      // ext: `eta` already is eta - (couplings to neighbouring blocks) x, computed by a separate full-occupancy kernel --
      // the dependent global loads of these hops are what a CTA of 8 warps hides worst (ncu: a third of the visit)
      gl_hops(op, cx_, sE, (~E.in) & 0xFFu, x, y);
#pragma unroll
      for (int c = 0; c < 6; c++) rE[c] -= y[c];
      gl_hops(op, cx_, sO, (~O.in) & 0xFFu, x, y);
#pragma unroll
      for (int c = 0; c < 6; c++) rO[c] -= y[c];
    }
    __syncthreads();                                                    // x_e in the buffer; C_e consumed by everybody
    if (tid == 0) cload(srcCinvO);
    mbar_wait(barU, 0); links_ready = true;
    sm_hops2(sm, cx_, O, y);                                            // N_oe x_e
#pragma unroll
    for (int c = 0; c < 6; c++) rO[c] -= y[c];
    __syncthreads();
    cf xo[6];
    load_own(x, vO, cx_, xo);
    put2(sm, cx_, E.l, xo);                                             // odd site O.l sits at index E.l of its parity
    __syncthreads();
    sm_hops2(sm, cx_, E, y);                                            // N_eo x_o
#pragma unroll
    for (int c = 0; c < 6; c++) rE[c] -= y[c];
    __syncthreads();
  }

  // block solve.  k = 0: e_o = Coo^-1 r'_o - x_o, t_e = r_e - N_eo e_o.   k = 1..biter: one MR step on the Schur
  // complement S = C_ee - N_eo Coo^-1 N_oe.   k = biter+1: x_o = Coo^-1 (r'_o - N_oe e_e), x_e += e_e.
  cf tE[6], eE[6];
#pragma unroll
  for (int c = 0; c < 6; c++) { tE[c] = rE[c]; eE[c] = cf(0.f, 0.f); }
#pragma unroll 1
  for (int k = 0; k <= biter + 1; k++) {
    cf wv[6], z[6];
    if (k > 0) sm_hops2(sm, cx_, O, wv);                                // N_oe (t_e or e_e)
    if (k == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) wv[c] = rO[c];
      cwait();                                                          // Coo^-1 in the buffer (stays for k = 1)
    } else if (k == biter + 1) {
      cf ro[6];
      load_own(x, vO, cx_, ro);                                         // parked r'_o
#pragma unroll
      for (int c = 0; c < 6; c++) wv[c] = ro[c] - wv[c];
      if (biter > 0) cwait();
    } else if (k > 1) cwait();
    clov_half_sm(CsT, cx_, wv, z);
    if (k == biter + 1) {
      if (!first_zero) {
        cf a[6];
        load_own(x, vE, cx_, a);
#pragma unroll
        for (int c = 0; c < 6; c++) eE[c] += a[c];
      }
      store_own(x, vE, cx_, eE);
      store_own(x, vO, cx_, z);
      break;
    }
    if (k == 0) {
      if (!first_zero) {
        cf xo[6];
        load_own(x, vO, cx_, xo);
#pragma unroll
        for (int c = 0; c < 6; c++) z[c] -= xo[c];                      // e_o = Coo^-1 r'_o - x_o
      }
      store_own(x, vO, cx_, rO);                                        // park r'_o (x_o is dead until the final write)
    } else {
#pragma unroll
      for (int c = 0; c < 6; c++) z[c] = -z[c];                         // a2_o = -Coo^-1 N_oe t_e
    }
    __syncthreads();                                                    // everybody is done with the even buffer and with Coo^-1
    if (tid == 0 && k > 0) cload(srcCe);
    put2(sm, cx_, E.l, z);
    __syncthreads();
    if (!links_ready) { mbar_wait(barU, 0); links_ready = true; }
    cf y[6];
    sm_hops2(sm, cx_, E, y);                                            // N_eo (e_o or a2_o)
    if (k == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) tE[c] -= y[c];
      __syncthreads();                                                  // odd buffer consumed
    } else {
      cf Dr[6];
      cwait();
      clov_half_sm(CsT, cx_, tE, Dr);
#pragma unroll
      for (int c = 0; c < 6; c++) Dr[c] += y[c];
      // alpha = <Dr,t>/<Dr,Dr> over the even sites of the block (local_minres, linsolve_generic.c:1013-1022)
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        p0 = __fmaf_rn(Dr[c].im, tE[c].im, __fmaf_rn(Dr[c].re, tE[c].re, p0));
        p1 = __fmaf_rn(-Dr[c].im, tE[c].re, __fmaf_rn(Dr[c].re, tE[c].im, p1));
        p2 = __fmaf_rn(Dr[c].im, Dr[c].im, __fmaf_rn(Dr[c].re, Dr[c].re, p2));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
      }
      float (*red)[4] = sm.red[k & 1];
      if (lane == 0) { red[w][0] = p0; red[w][1] = p1; red[w][2] = p2; }
      __syncthreads();                                                  // also: odd buffer and C_ee consumed by everybody
      if (tid == 0) cload(srcCinvO);
      p0 = p1 = p2 = 0.f;
#pragma unroll
      for (int i = 0; i < BS / 32; i++) { p0 += red[i][0]; p1 += red[i][1]; p2 += red[i][2]; }
      cf alpha(0.f, 0.f);
      if (p2 > 1e-30f) alpha = cf(p0 / p2, p1 / p2);
#pragma unroll
      for (int c = 0; c < 6; c++) { fma_(eE[c], alpha, tE[c]); fms_(tE[c], alpha, Dr[c]); }
    }
    put2(sm, cx_, E.l, (k < biter) ? tE : eE);
    __syncthreads();
  }
}

}  // namespace sap

// fine-level SAP with the fused block kernel; same iteration as the generic path of mg_smoother
bool sap_fine_fast_available(const Solver &s) {
  const Geometry &g = s.lev[0].geo;
  return g.sh == 5 && g.block_eo && g.bs == 256 && g.bs_even == 128;
}

void sap_fine_fast(Solver &s, cf *x, const cf *eta, int iters, bool zero_guess) {
  Level &L = s.lev[0];
  const Geometry &g = L.geo;
  const int biter = s.p.block_iter[0];
  static int version = 0;
  if (!version) {
    const char *e = getenv("DDA_SAP_V1");
    version = (e && atoi(e) != 0) ? 1 : 2;
    CUDA_CHECK(cudaFuncSetAttribute(sap::k_sap_fine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sap::Shared)));
    CUDA_CHECK(cudaFuncSetAttribute(sap::k_sap_fine2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sap::Shared2)));
  }
  const bool v2 = version == 2 && g.d_saptab && L.Dblk;
  static int split = -1;     // 1: couplings to neighbouring blocks of the block residual in a kernel of their own (default)
  if (split < 0) { const char *e = getenv("DDA_SAP_SPLIT"); split = e ? atoi(e) : 1; }
  cf *rext = L.w[0];
  auto launch = [&](const int *list, int nblk, int first) {
    if (nblk <= 0) return;
    if (v2 && split && !first) {
      dw_outer_fast(L.opf, rext, x, eta, list, nblk, g.bs);
      sap::k_sap_fine2<<<nblk, sap::BS, sizeof(sap::Shared2), g_stream>>>(L.opf, L.Dblk, x, rext, list, g.d_saptab, biter, 0, 1);
    } else if (v2) sap::k_sap_fine2<<<nblk, sap::BS, sizeof(sap::Shared2), g_stream>>>(L.opf, L.Dblk, x, eta, list, g.d_saptab, biter, first, 0);
    else sap::k_sap_fine<<<nblk, sap::BS, sizeof(sap::Shared), g_stream>>>(L.opf, x, eta, list, biter, first);
    g_launch_count++;
  };
  if (zero_guess) vzero(x, g.vlen());
  for (int cyc = 0; cyc < iters; cyc++)
    for (int col = 0; col < 2; col++) {
      const int nblk = g.nblk_color[col];
      if (nblk == 0) continue;
      const int first = (zero_guess && cyc == 0 && col == 0) ? 1 : 0;
      if (first || !g.partitioned()) {
        launch(g.d_blocklist[col], nblk, first);
      } else {
        // block residuals read x of neighbouring blocks on other ranks: exchange the ghost slabs on the second stream
        // while the blocks away from the rank boundary are solved, then the blocks that touch it
        halo_begin<cf>(g, x, 12, g.sh);
        launch(g.d_blocklist_int[col], g.nblk_int[col], 0);
        halo_end(g);
        launch(g.d_blocklist_bnd[col], g.nblk_bnd[col], 0);
      }
#ifdef DDA_DEBUG_SYNC
      CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
    }
  double rf = s.p.relax_fac[0];
  if (rf != 1.0) vscale(x, x, rf, g.vlen());
}

#endif

}  // namespace dda

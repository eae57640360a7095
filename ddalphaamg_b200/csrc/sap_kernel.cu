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
// is one kernel: the block's links (73.7 KB in float) live in shared memory for the whole visit, the iteration
// vectors live in registers (each thread owns one even and one odd site of the block, so every thread is busy in
// both half steps), and only the vector that the opposite parity has to gather is exchanged through a 24.6 KB shared
// buffer.  Clover blocks (even sites) and their inverses (odd sites) are streamed from L2.  Inner products of the MR
// step are warp-shuffle + one shared-memory stage.  98.5 KB shared memory per CTA -> 2 CTAs per SM.
//
// HBM traffic per block visit (algorithmic): links 256*288 B + clover/inverse 256*288 B (+L2 re-reads) + x, eta in,
// x out.
#include "solver.h"
#include "fine_op.cuh"
#include "halo.h"

namespace dda {

#ifndef DDA_HOST_EMU

template <int BS> struct SapShared {
  float2 U[36][BS];         // links of the block's sites, [9*mu + 3*row + col][site]
  float2 Vb[12][BS];        // exchange buffer: [component][site in block]
  float red[2][BS / 64][4];
};

__device__ __forceinline__ cf ld2(const float2 &v) { return cf(v.x, v.y); }

// hop part of the block operator restricted to in-block neighbours, gather form, operands in shared memory:
//   out -= (1-gamma_mu) D_mu(x) v(x+mu) + (1+gamma_mu) D_mu(x-mu)^dagger v(x-mu)
template <int MU, int BS>
__device__ __forceinline__ void sm_hop_pair(const SapShared<BS> &sm, int l, unsigned inmask, unsigned nbf, unsigned nbb, cf *out) {
  if (inmask & (1u << MU)) {
    const int n = (nbf >> (8 * MU)) & 0xFF;
    cf p[12], h[6], g[6], M[9];
#pragma unroll
    for (int c = 0; c < 12; c++) p[c] = ld2(sm.Vb[c][n]);
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = ld2(sm.U[9 * MU + k][l]);
    project<MU, +1>(p, h);
    su3_mul(M, h, g);
    reconstruct_sub<MU, +1>(g, out);
  }
  if (inmask & (1u << (4 + MU))) {
    const int n = (nbb >> (8 * MU)) & 0xFF;
    cf p[12], h[6], g[6], M[9];
#pragma unroll
    for (int c = 0; c < 12; c++) p[c] = ld2(sm.Vb[c][n]);
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = ld2(sm.U[9 * MU + k][n]);
    project<MU, -1>(p, h);
    su3_mul_dag(M, h, g);
    reconstruct_sub<MU, -1>(g, out);
  }
}
template <int BS>
__device__ __forceinline__ void sm_hops(const SapShared<BS> &sm, int l, unsigned inmask, unsigned nbf, unsigned nbb, cf *out) {
  sm_hop_pair<0, BS>(sm, l, inmask, nbf, nbb, out);
  sm_hop_pair<1, BS>(sm, l, inmask, nbf, nbb, out);
  sm_hop_pair<2, BS>(sm, l, inmask, nbf, nbb, out);
  sm_hop_pair<3, BS>(sm, l, inmask, nbf, nbb, out);
}

template <int BS>
__device__ __forceinline__ void put(SapShared<BS> &sm, int l, const cf *v) {
#pragma unroll
  for (int c = 0; c < 12; c++) sm.Vb[c][l] = make_float2(v[c].re, v[c].im);
}

// site-local packed Hermitian 2x(6x6) multiply from global memory (tiled layout), y = M x
__device__ __forceinline__ void clov(const float *__restrict__ C, long tile_c, int lane, const cf *x, cf *y) {
  const float *Cs = C + tile_c + lane;
#pragma unroll
  for (int b = 0; b < 2; b++) {
#pragma unroll
    for (int i = 0; i < 6; i++) y[6 * b + i] = __ldg(Cs + ((6 * b + i) << 5)) * x[6 * b + i];
    int m = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = i + 1; j < 6; j++, m++) {
        cf cij(__ldg(Cs + ((12 + 2 * (15 * b + m)) << 5)), __ldg(Cs + ((12 + 2 * (15 * b + m) + 1) << 5)));
        fma_(y[6 * b + i], cij, x[6 * b + j]);
        fmac_(y[6 * b + j], cij, x[6 * b + i]);
      }
  }
}

// first_zero: x == 0 on the whole lattice on entry (first colour of a zero-guess call): r = eta, x = e.
template <int BS>
__global__ void __launch_bounds__(BS / 2, 2)
k_sap_fine(FineOp<float> op, cf *__restrict__ x, const cf *__restrict__ eta, const int *__restrict__ blocklist,
           int biter, int first_zero) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SapShared<BS> &sm = *reinterpret_cast<SapShared<BS> *>(smem_raw);
  const int j = threadIdx.x, lane = j & 31, w = j >> 5;
  const int H = BS / 2;
  const int lE = j, lO = H + j;
  const long base = (long)blocklist[blockIdx.x] * BS;
  const long sE = base + lE, sO = base + lO;
  const long V = op.V;
  // tiled global offsets (32-site tiles, component-major inside a tile)
  const long tlE = sE >> 5, tlO = sO >> 5;
  const long vE = tlE * (12L << 5) + lane, vO = tlO * (12L << 5) + lane;
  const long cE = tlE * (72L << 5), cO = tlO * (72L << 5);

  // links of the block -> shared memory (coalesced 256 B rows)
  {
    const float2 *D2 = reinterpret_cast<const float2 *>(op.D);
    const long uE = tlE * (36L << 5) + lane, uO = tlO * (36L << 5) + lane;
#pragma unroll 4
    for (int k = 0; k < 36; k++) {
      sm.U[k][lE] = __ldg(D2 + uE + ((long)k << 5));
      sm.U[k][lO] = __ldg(D2 + uO + ((long)k << 5));
    }
  }
  // in-block neighbour indices (blocks are contiguous site ranges in the native order)
  const unsigned fE = op.blkflag[sE], fO = op.blkflag[sO];
  unsigned nfE = 0, nbE = 0, nfO = 0, nbO = 0;
#pragma unroll
  for (int d = 0; d < 4; d++) {
    nfE |= (unsigned)((__ldg(op.nb + (long)d * V + sE) - base) & 0xFF) << (8 * d);
    nbE |= (unsigned)((__ldg(op.nb + (long)(4 + d) * V + sE) - base) & 0xFF) << (8 * d);
    nfO |= (unsigned)((__ldg(op.nb + (long)d * V + sO) - base) & 0xFF) << (8 * d);
    nbO |= (unsigned)((__ldg(op.nb + (long)(4 + d) * V + sO) - base) & 0xFF) << (8 * d);
  }
  const unsigned inE = (~fE) & 0xFFu, inO = (~fO) & 0xFFu;

  cf rE[12], rO[12];
#pragma unroll
  for (int c = 0; c < 12; c++) { rE[c] = eta[vE + ((long)c << 5)]; rO[c] = eta[vO + ((long)c << 5)]; }
  if (!first_zero) {
    // r = eta - D x on the block: clover + in-block hops (shared memory) + couplings to neighbouring blocks (global)
    cf xE[12], xO[12], y[12];
#pragma unroll
    for (int c = 0; c < 12; c++) { xE[c] = x[vE + ((long)c << 5)]; xO[c] = x[vO + ((long)c << 5)]; }
    put<BS>(sm, lE, xE); put<BS>(sm, lO, xO);
    clov(op.C, cE, lane, xE, y);
#pragma unroll
    for (int c = 0; c < 12; c++) rE[c] -= y[c];
    clov(op.C, cO, lane, xO, y);
#pragma unroll
    for (int c = 0; c < 12; c++) rO[c] -= y[c];
    // cross-block hops: hop_pair subtracts from its accumulator, so accumulate -N x into y and add
#pragma unroll
    for (int c = 0; c < 12; c++) y[c] = cf(0.f, 0.f);
    if (fE) { hop_pair<0>(op, sE, fE, x, y); hop_pair<1>(op, sE, fE, x, y); hop_pair<2>(op, sE, fE, x, y); hop_pair<3>(op, sE, fE, x, y); }
#pragma unroll
    for (int c = 0; c < 12; c++) { rE[c] -= y[c]; y[c] = cf(0.f, 0.f); }
    if (fO) { hop_pair<0>(op, sO, fO, x, y); hop_pair<1>(op, sO, fO, x, y); hop_pair<2>(op, sO, fO, x, y); hop_pair<3>(op, sO, fO, x, y); }
#pragma unroll
    for (int c = 0; c < 12; c++) { rO[c] -= y[c]; y[c] = cf(0.f, 0.f); }
    __syncthreads();
    sm_hops<BS>(sm, lE, inE, nfE, nbE, y);
#pragma unroll
    for (int c = 0; c < 12; c++) { rE[c] -= y[c]; y[c] = cf(0.f, 0.f); }
    sm_hops<BS>(sm, lO, inO, nfO, nbO, y);
#pragma unroll
    for (int c = 0; c < 12; c++) rO[c] -= y[c];
  }
  __syncthreads();

  // e_o = Doo^-1 r_o ; t_e = r_e - N_eo e_o        (hop accumulators hold N v with the operator's sign: D = C + N)
  cf tE[12], eE[12];
  {
    cf eO[12];
    clov(op.Cinv, cO, lane, rO, eO);
    put<BS>(sm, lO, eO);
  }
  __syncthreads();
  {
    cf y[12];
#pragma unroll
    for (int c = 0; c < 12; c++) y[c] = cf(0.f, 0.f);
    sm_hops<BS>(sm, lE, inE, nfE, nbE, y);
#pragma unroll
    for (int c = 0; c < 12; c++) { tE[c] = rE[c] - y[c]; eE[c] = cf(0.f, 0.f); }
  }
  // minimal residual on S = C_ee - N_eo Coo^-1 N_oe
  for (int it = 0; it < biter; it++) {
    put<BS>(sm, lE, tE);
    __syncthreads();
    {
      cf a[12], a2[12];
#pragma unroll
      for (int c = 0; c < 12; c++) a[c] = cf(0.f, 0.f);
      sm_hops<BS>(sm, lO, inO, nfO, nbO, a);          // a_o = N_oe t_e
      clov(op.Cinv, cO, lane, a, a2);
#pragma unroll
      for (int c = 0; c < 12; c++) a2[c] = -a2[c];      // a2_o = -Coo^-1 a_o
      put<BS>(sm, lO, a2);
    }
    __syncthreads();
    cf Dr[12];
    clov(op.C, cE, lane, tE, Dr);
    {
      cf y[12];
#pragma unroll
      for (int c = 0; c < 12; c++) y[c] = cf(0.f, 0.f);
      sm_hops<BS>(sm, lE, inE, nfE, nbE, y);           // N_eo a2_o
#pragma unroll
      for (int c = 0; c < 12; c++) Dr[c] += y[c];
    }
    // alpha = <Dr,t>/<Dr,Dr> over the even sites of the block
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int c = 0; c < 12; c++) {
      p0 += Dr[c].re * tE[c].re + Dr[c].im * tE[c].im;
      p1 += Dr[c].re * tE[c].im - Dr[c].im * tE[c].re;
      p2 += Dr[c].re * Dr[c].re + Dr[c].im * Dr[c].im;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    float (*red)[4] = sm.red[it & 1];
    if (lane == 0) { red[w][0] = p0; red[w][1] = p1; red[w][2] = p2; }
    __syncthreads();
    p0 = p1 = p2 = 0.f;
#pragma unroll
    for (int k = 0; k < BS / 64; k++) { p0 += red[k][0]; p1 += red[k][1]; p2 += red[k][2]; }
    cf alpha(0.f, 0.f);
    if (p2 > 1e-30f) alpha = cf(p0 / p2, p1 / p2);
#pragma unroll
    for (int c = 0; c < 12; c++) { fma_(eE[c], alpha, tE[c]); fms_(tE[c], alpha, Dr[c]); }
  }
  // back substitution: e_o = Coo^-1 (r_o - N_oe e_e) ; x += e
  put<BS>(sm, lE, eE);
  __syncthreads();
  {
    cf y[12], eO[12];
#pragma unroll
    for (int c = 0; c < 12; c++) y[c] = cf(0.f, 0.f);
    sm_hops<BS>(sm, lO, inO, nfO, nbO, y);
#pragma unroll
    for (int c = 0; c < 12; c++) y[c] = rO[c] - y[c];
    clov(op.Cinv, cO, lane, y, eO);
    if (first_zero) {
#pragma unroll
      for (int c = 0; c < 12; c++) { x[vE + ((long)c << 5)] = eE[c]; x[vO + ((long)c << 5)] = eO[c]; }
    } else {
#pragma unroll
      for (int c = 0; c < 12; c++) {
        cf a = x[vE + ((long)c << 5)], b = x[vO + ((long)c << 5)];
        x[vE + ((long)c << 5)] = a + eE[c]; x[vO + ((long)c << 5)] = b + eO[c];
      }
    }
  }
}

// fine-level SAP with the fused block kernel; same iteration as the generic path of mg_smoother
bool sap_fine_fast_available(const Solver &s) {
  const Geometry &g = s.lev[0].geo;
  return g.sh == 5 && g.block_eo && g.bs == 256 && g.bs_even == 128;
}

void sap_fine_fast(Solver &s, cf *x, const cf *eta, int iters, bool zero_guess) {
  Level &L = s.lev[0];
  const Geometry &g = L.geo;
  const int biter = s.p.block_iter[0];
  static bool attr_set = false;
  const int smem = (int)sizeof(SapShared<256>);
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_sap_fine<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (zero_guess) vzero(x, g.vlen());
  for (int cyc = 0; cyc < iters; cyc++)
    for (int col = 0; col < 2; col++) {
      const int nblk = g.nblk_color[col];
      if (nblk == 0) continue;
      const int first = (zero_guess && cyc == 0 && col == 0) ? 1 : 0;
      if (!first) halo_exchange<cf>(g, x, 12, g.sh);   // block residuals read x of neighbouring blocks on other ranks
      k_sap_fine<256><<<nblk, 128, smem, g_stream>>>(L.opf, x, eta, g.d_blocklist[col], biter, first);
      g_launch_count++;
#ifdef DDA_DEBUG_SYNC
      CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
    }
  double rf = s.p.relax_fac[0];
  if (rf != 1.0) vscale(x, x, rf, g.vlen());
}

#endif

}  // namespace dda

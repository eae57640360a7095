// galerkin_kernel.cu -- fused Galerkin construction of the level-1 operator from the fine Wilson-Clover operator (sm_100a):
//   S(a)[:, j]    = P_a^H (C + N_inside-aggregate) P_a e_j ,      F_mu(a)[:, j] = P_a^H N_{+mu, leaving the aggregate} P e_j
// for every aggregate a and coarse column j = (chirality, test vector).
//
// Reference counterparts: coarse_operator_PRECISION_setup / set_coarse_self_coupling / set_coarse_neighbor_coupling
// (coarse_operator_generic.c:53-205) with the aggregate-restricted operator pieces of dirac_generic.c:308-462.  The generic
// path (mg_setup.cu) does, per column, five masked operator applications and five restrictions (2 x 5 x 40 launches, each
// restriction re-reading all Nv prolongator vectors): 1.4 s per rebuild at 48^3 x 96, and a setup runs four rebuilds.
// Here ONE launch handles a range of columns: a CTA owns an aggregate (4^4 sites, one thread per site), computes the five
// pieces of D P e_j for its site in registers (clover + in-aggregate hops -> S piece; forward hops that leave the aggregate
// -> F_mu pieces), and contracts them with the Nv prolongator vectors read once per column (warp-shuffle reductions, partial
// sums per warp in shared memory).  No intermediate vectors touch HBM.
#include <type_traits>
#include "solver.h"
#include "fine_op.cuh"

namespace dda {

#ifndef DDA_HOST_EMU

namespace {

const int GAS = 256;       // sites per aggregate = threads per CTA

__device__ __forceinline__ cf ldcf(const cf *p) { const float2 v = __ldg(reinterpret_cast<const float2 *>(p)); return cf(v.x, v.y); }

// spinor of site n restricted to chirality ch (dofs 6 ch .. 6 ch + 5), other dofs zero
__device__ __forceinline__ void load_chiral(const cf *__restrict__ v, long n, int ch, cf *p) {
  const long nt = (n >> 5) * (12L << 5) + (n & 31);
#pragma unroll
  for (int c = 0; c < 12; c++) p[c] = cf(0.f, 0.f);
  if (ch == 0) {
#pragma unroll
    for (int c = 0; c < 6; c++) p[c] = ldcf(v + nt + ((long)c << 5));
  } else {
#pragma unroll
    for (int c = 6; c < 12; c++) p[c] = ldcf(v + nt + ((long)c << 5));
  }
}

template <int MU>
__device__ __forceinline__ void hop_fwd(const cf *__restrict__ D, const cf *__restrict__ v, long s, long n, int ch, cf *out) {
  cf p[12], h[6], g[6], M[9];
  load_chiral(v, n, ch, p);
  const long su = (s >> 5) * (36L << 5) + (s & 31);
#pragma unroll
  for (int k = 0; k < 9; k++) M[k] = ldcf(D + su + ((long)(9 * MU + k) << 5));
  project<MU, +1>(p, h);
  su3_mul(M, h, g);
  reconstruct_sub<MU, +1>(g, out);
}
template <int MU>
__device__ __forceinline__ void hop_bwd(const cf *__restrict__ D, const cf *__restrict__ v, long n, int ch, cf *out) {
  cf p[12], h[6], g[6], M[9];
  load_chiral(v, n, ch, p);
  const long nu = (n >> 5) * (36L << 5) + (n & 31);
#pragma unroll
  for (int k = 0; k < 9; k++) M[k] = ldcf(D + nu + ((long)(9 * MU + k) << 5));
  project<MU, -1>(p, h);
  su3_mul_dag(M, h, g);
  reconstruct_sub<MU, -1>(g, out);
}

template <int MU>
__device__ __forceinline__ void hop_dir(const FineOp<float> &op, const cf *__restrict__ v, long s, unsigned af, int ch, cf *wS, cf *wmu) {
  const long nf = __ldg(op.nb + (long)MU * op.V + s);
  if (af & (1u << MU)) hop_fwd<MU>(op.D, v, s, nf, ch, wmu);              // leaves the aggregate: F_mu piece
  else hop_fwd<MU>(op.D, v, s, nf, ch, wS);
  if (!(af & (1u << (4 + MU)))) hop_bwd<MU>(op.D, v, __ldg(op.nb + (long)(4 + MU) * op.V + s), ch, wS);
}

__global__ void __launch_bounds__(GAS, 1)
k_galerkin_fine(FineOp<float> op, Transfer t, cf *__restrict__ S, cf *__restrict__ F, int j0, int j1) {
  __shared__ float part[GAS / 32][MAX_NV][20];
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = t.nv, n = 2 * nv;
  const long nn = (long)n * n;
  const long s = (long)a * GAS + tid;
  const long st = (s >> 5) * (12L << 5) + (s & 31);
  const unsigned af = op.aggflag[s];
  const unsigned any_face = __ballot_sync(0xffffffffu, (af & 0xFu) != 0) ? 1u : 0u;
  unsigned face_warp[4];
#pragma unroll
  for (int mu = 0; mu < 4; mu++) face_warp[mu] = __ballot_sync(0xffffffffu, (af >> mu) & 1u);
  (void)any_face;
  const long ca = t.agg2coarse[a];
  for (int j = j0; j < j1; j++) {
    const int ch = j / nv, kj = j - ch * nv;
    const cf *__restrict__ src = t.P[kj];
    cf w[5][12];                                                       // S piece, F_0 .. F_3 pieces of D P e_j at this site
#pragma unroll
    for (int q = 0; q < 5; q++)
#pragma unroll
      for (int c = 0; c < 12; c++) w[q][c] = cf(0.f, 0.f);
    {
      cf x[12];
      load_chiral(src, s, ch, x);
      clover_mul(op.C, 5, s, x, w[0]);
    }
    hop_dir<0>(op, src, s, af, ch, w[0], w[1]);
    hop_dir<1>(op, src, s, af, ch, w[0], w[2]);
    hop_dir<2>(op, src, s, af, ch, w[0], w[3]);
    hop_dir<3>(op, src, s, af, ch, w[0], w[4]);
    // contraction with the prolongator vectors: value (piece q, chirality c2) = sum over the aggregate of conj(P_k) w_q
    for (int k = 0; k < nv; k++) {
      const cf *__restrict__ Pk = t.P[k];
      cf p[12];
#pragma unroll
      for (int c = 0; c < 12; c++) p[c] = ldcf(Pk + st + ((long)c << 5));
#pragma unroll
      for (int q = 0; q < 5; q++) {
        if (q > 0 && face_warp[q - 1] == 0) continue;                    // no site of this warp on the +mu face: piece is zero
        float v0r = 0.f, v0i = 0.f, v1r = 0.f, v1i = 0.f;
#pragma unroll
        for (int c = 0; c < 6; c++) {
          v0r = __fmaf_rn(p[c].im, w[q][c].im, __fmaf_rn(p[c].re, w[q][c].re, v0r));
          v0i = __fmaf_rn(-p[c].im, w[q][c].re, __fmaf_rn(p[c].re, w[q][c].im, v0i));
          v1r = __fmaf_rn(p[6 + c].im, w[q][6 + c].im, __fmaf_rn(p[6 + c].re, w[q][6 + c].re, v1r));
          v1i = __fmaf_rn(-p[6 + c].im, w[q][6 + c].re, __fmaf_rn(p[6 + c].re, w[q][6 + c].im, v1i));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v0r += __shfl_xor_sync(0xffffffffu, v0r, o); v0i += __shfl_xor_sync(0xffffffffu, v0i, o);
          v1r += __shfl_xor_sync(0xffffffffu, v1r, o); v1i += __shfl_xor_sync(0xffffffffu, v1i, o);
        }
        if (lane == 0) { part[warp][k][4 * q] = v0r; part[warp][k][4 * q + 1] = v0i; part[warp][k][4 * q + 2] = v1r; part[warp][k][4 * q + 3] = v1i; }
      }
#pragma unroll
      for (int q = 1; q < 5; q++)
        if (face_warp[q - 1] == 0 && lane == 0) { part[warp][k][4 * q] = 0.f; part[warp][k][4 * q + 1] = 0.f; part[warp][k][4 * q + 2] = 0.f; part[warp][k][4 * q + 3] = 0.f; }
    }
    __syncthreads();
    // outputs of this column: 5 pieces x 2 chiralities x nv complex
    for (int o = tid; o < 10 * nv; o += GAS) {
      const int q = o / (2 * nv), r = o - q * 2 * nv, c2 = r / nv, k = r - c2 * nv;
      float re = 0.f, im = 0.f;
#pragma unroll
      for (int wq = 0; wq < GAS / 32; wq++) { re += part[wq][k][4 * q + 2 * c2]; im += part[wq][k][4 * q + 2 * c2 + 1]; }
      cf *dst = (q == 0) ? S + ca * nn : F + (ca * 4 + (q - 1)) * nn;
      dst[(long)j * n + c2 * nv + k] = cf(re, im);
    }
    __syncthreads();
  }
}


// Version 2 (default): the pieces F_mu of a column are non-zero only on the +mu face of the aggregate (64 of 256 sites).
// Version 1 above keeps all five pieces of a site in registers (120 of 181 registers, 1 CTA per SM) and contracts every
// piece in every warp that owns a face site (ncu: 28 G warp instructions per rebuild at 32^3 x 64, 26 % of them the
// SHFL + FADD of the warp reductions, issue 36 % at 8 warps per SM).  Here a thread keeps only the S piece of its site;
// the F_mu pieces go through a compact shared-memory buffer [mu][face site] and are contracted by a second thread
// assignment (thread u -> face u / 64, face site u % 64: the four faces occupy exactly the 256 threads), so that per
// prolongator vector a thread does two 12-component contractions and two warp reductions instead of up to five.
__global__ void __launch_bounds__(GAS, 2)
k_galerkin_fine2(FineOp<float> op, Transfer t, cf *__restrict__ S, cf *__restrict__ F, int j0, int j1) {
  __shared__ float partS[GAS / 32][MAX_NV][4];
  __shared__ float partF[4][2][MAX_NV][4];
  __shared__ float2 Fbuf[4][GAS / 4][12];
  __shared__ short fsite[4][GAS / 4];
  __shared__ int fcnt[GAS / 32][4];
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = t.nv, n = 2 * nv;
  const long nn = (long)n * n;
  const long s = (long)a * GAS + tid;
  const long st = (s >> 5) * (12L << 5) + (s & 31);
  const unsigned af = op.aggflag[s];
  // index of this site inside each +mu face it belongs to (site order inside the aggregate is whatever the geometry uses)
  int fidx[4];
  {
    unsigned bal[4];
#pragma unroll
    for (int mu = 0; mu < 4; mu++) {
      bal[mu] = __ballot_sync(0xffffffffu, (af >> mu) & 1u);
      if (lane == 0) fcnt[warp][mu] = __popc(bal[mu]);
    }
    __syncthreads();
#pragma unroll
    for (int mu = 0; mu < 4; mu++) {
      int off = 0;
      for (int w2 = 0; w2 < warp; w2++) off += fcnt[w2][mu];
      fidx[mu] = ((af >> mu) & 1u) ? off + __popc(bal[mu] & ((1u << lane) - 1u)) : -1;
      if (fidx[mu] >= GAS / 4) fidx[mu] = -1;                            // cannot happen for a 4^4 aggregate (host check)
      if (fidx[mu] >= 0) fsite[mu][fidx[mu]] = (short)tid;
    }
    __syncthreads();
  }
  const int muF = tid >> 6, fi = tid & 63;                               // second assignment: face muF, face site fi
  const long sF = (long)a * GAS + fsite[muF][fi];
  const long stF = (sF >> 5) * (12L << 5) + (sF & 31);
  const long ca = t.agg2coarse[a];
  for (int j = j0; j < j1; j++) {
    const int ch = j / nv, kj = j - ch * nv;
    const cf *__restrict__ src = t.P[kj];
    cf w0[12];                                                           // S piece of D P e_j at this site
    {
      cf x[12];
      load_chiral(src, s, ch, x);
      clover_mul(op.C, 5, s, x, w0);
    }
    auto dir = [&](auto MUc) {
      constexpr int MU = decltype(MUc)::value;
      const long nf = __ldg(op.nb + (long)MU * op.V + s);
      if (af & (1u << MU)) {                                             // leaves the aggregate: F_mu piece -> shared memory
        cf tmp[12];
#pragma unroll
        for (int c = 0; c < 12; c++) tmp[c] = cf(0.f, 0.f);
        hop_fwd<MU>(op.D, src, s, nf, ch, tmp);
        if (fidx[MU] >= 0) {
#pragma unroll
          for (int c = 0; c < 12; c++) Fbuf[MU][fidx[MU]][c] = make_float2(tmp[c].re, tmp[c].im);
        }
      } else hop_fwd<MU>(op.D, src, s, nf, ch, w0);
      if (!(af & (1u << (4 + MU)))) hop_bwd<MU>(op.D, src, __ldg(op.nb + (long)(4 + MU) * op.V + s), ch, w0);
    };
    dir(std::integral_constant<int, 0>()); dir(std::integral_constant<int, 1>());
    dir(std::integral_constant<int, 2>()); dir(std::integral_constant<int, 3>());
    __syncthreads();
    cf wf[12];
#pragma unroll
    for (int c = 0; c < 12; c++) { const float2 v = Fbuf[muF][fi][c]; wf[c] = cf(v.x, v.y); }
    // contraction with the prolongator vectors
    for (int k = 0; k < nv; k++) {
      const cf *__restrict__ Pk = t.P[k];
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = 0.f;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const cf p0 = ldcf(Pk + st + ((long)c << 5)), p1 = ldcf(Pk + st + ((long)(6 + c) << 5));
        v[0] = __fmaf_rn(p0.im, w0[c].im, __fmaf_rn(p0.re, w0[c].re, v[0]));
        v[1] = __fmaf_rn(-p0.im, w0[c].re, __fmaf_rn(p0.re, w0[c].im, v[1]));
        v[2] = __fmaf_rn(p1.im, w0[6 + c].im, __fmaf_rn(p1.re, w0[6 + c].re, v[2]));
        v[3] = __fmaf_rn(-p1.im, w0[6 + c].re, __fmaf_rn(p1.re, w0[6 + c].im, v[3]));
        const cf q0 = ldcf(Pk + stF + ((long)c << 5)), q1 = ldcf(Pk + stF + ((long)(6 + c) << 5));
        v[4] = __fmaf_rn(q0.im, wf[c].im, __fmaf_rn(q0.re, wf[c].re, v[4]));
        v[5] = __fmaf_rn(-q0.im, wf[c].re, __fmaf_rn(q0.re, wf[c].im, v[5]));
        v[6] = __fmaf_rn(q1.im, wf[6 + c].im, __fmaf_rn(q1.re, wf[6 + c].re, v[6]));
        v[7] = __fmaf_rn(-q1.im, wf[6 + c].re, __fmaf_rn(q1.re, wf[6 + c].im, v[7]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) { partS[warp][k][q] = v[q]; partF[muF][warp & 1][k][q] = v[4 + q]; }
      }
    }
    __syncthreads();
    // outputs of this column: 5 pieces x 2 chiralities x nv complex
    for (int o = tid; o < 10 * nv; o += GAS) {
      const int q = o / (2 * nv), r = o - q * 2 * nv, c2 = r / nv, k = r - c2 * nv;
      float re = 0.f, im = 0.f;
      if (q == 0) {
#pragma unroll
        for (int wq = 0; wq < GAS / 32; wq++) { re += partS[wq][k][2 * c2]; im += partS[wq][k][2 * c2 + 1]; }
      } else {
        re = partF[q - 1][0][k][2 * c2] + partF[q - 1][1][k][2 * c2];
        im = partF[q - 1][0][k][2 * c2 + 1] + partF[q - 1][1][k][2 * c2 + 1];
      }
      cf *dst = (q == 0) ? S + ca * nn : F + (ca * 4 + (q - 1)) * nn;
      dst[(long)j * n + c2 * nv + k] = cf(re, im);
    }
    __syncthreads();
  }
}

}  // namespace

// Galerkin operator of level 1 (all columns).  The prolongator vectors must have current ghost slabs.  Returns false when the
// shape is not supported (the caller falls back to the generic column-by-column construction).
bool galerkin_fine_fast(const FineOp<float> &op, const Transfer &t, cf *S, cf *F) {
  if (!(t.lay.sh == 5 && t.nc == 12 && t.as == GAS && t.nv <= MAX_NV && op.sh == 5)) return false;
  static const int v1 = []() { const char *e = getenv("DDA_GALERKIN_V1"); return e ? atoi(e) : 0; }();
  if (v1) k_galerkin_fine<<<t.nagg, GAS, 0, g_stream>>>(op, t, S, F, 0, 2 * t.nv);
  else k_galerkin_fine2<<<t.nagg, GAS, 0, g_stream>>>(op, t, S, F, 0, 2 * t.nv);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
}

#endif

}  // namespace dda

// fine_op.cuh -- per-site device functions of the fine Wilson-Clover operator.
//
//   D_W psi(x) = C(x) psi(x) - sum_mu [ (1-gamma_mu) D_mu(x) psi(x+mu) + (1+gamma_mu) D_mu(x-mu)^dagger psi(x-mu) ],
//   D_mu = U_mu/2, C = (4+m0) + clover term.
// Behaviour follows the reference's d_plus_clover_PRECISION (dirac_generic.c:159-277), site_clover
// (dirac_generic.h:723-799), the SU(3) mvm/mvmh kernels (dirac_generic.h:58-107) and the gamma basis BASIS0
// (clifford.h:39-100); the implementation is a single gather-form pass (one thread reads 8 neighbours, 8 links,
// the clover block and writes once) instead of the reference's five passes with materialised half spinors.
#pragma once
#include "common.cuh"
#include "fine_op.h"

namespace dda {

// gamma_mu (BASIS0, clifford.h:39-100): row r has its single non-zero in column CO[r] with value VAL[r],
// encoded 0:+1 1:-1 2:+i 3:-i.
template <int MU> struct Gam;
template <> struct Gam<0> { enum { c0 = 2, c1 = 3, c2 = 0, c3 = 1, v0 = 1, v1 = 1, v2 = 1, v3 = 1 }; };   // T
template <> struct Gam<1> { enum { c0 = 3, c1 = 2, c2 = 1, c3 = 0, v0 = 3, v1 = 3, v2 = 2, v3 = 2 }; };   // Z
template <> struct Gam<2> { enum { c0 = 3, c1 = 2, c2 = 1, c3 = 0, v0 = 1, v1 = 0, v2 = 0, v3 = 1 }; };   // Y
template <> struct Gam<3> { enum { c0 = 2, c1 = 3, c2 = 0, c3 = 1, v0 = 3, v1 = 2, v2 = 2, v3 = 3 }; };   // X

// half-spinor projection h = upper two spin rows of (1 - S*gamma_mu) phi   (S = +1 forward hop, -1 backward hop)
template <int MU, int S, class T> HD void project(const cx<T> *p, cx<T> *h) {
  typedef Gam<MU> G;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    cx<T> a = mul_unit<G::v0>(p[3 * G::c0 + c]), b = mul_unit<G::v1>(p[3 * G::c1 + c]);
    if (S > 0) { h[c] = p[c] - a; h[3 + c] = p[3 + c] - b; }
    else       { h[c] = p[c] + a; h[3 + c] = p[3 + c] + b; }
  }
}
// out -= (1 - S*gamma_mu) reconstructed from the link-multiplied half spinor g
template <int MU, int S, class T> HD void reconstruct_sub(const cx<T> *g, cx<T> *out) {
  typedef Gam<MU> G;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    out[c] -= g[c];
    out[3 + c] -= g[3 + c];
    cx<T> a = mul_unit<G::v2>(g[3 * G::c2 + c]), b = mul_unit<G::v3>(g[3 * G::c3 + c]);
    if (S > 0) { out[6 + c] += a; out[9 + c] += b; }
    else       { out[6 + c] -= a; out[9 + c] -= b; }
  }
}
// g = M h (two colour vectors), M row-major 3x3
template <class T> HD void su3_mul(const cx<T> *M, const cx<T> *h, cx<T> *g) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      cx<T> a = M[3 * r] * h[3 * s];
      fma_(a, M[3 * r + 1], h[3 * s + 1]);
      fma_(a, M[3 * r + 2], h[3 * s + 2]);
      g[3 * s + r] = a;
    }
}
// g = M^dagger h
template <class T> HD void su3_mul_dag(const cx<T> *M, const cx<T> *h, cx<T> *g) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      cx<T> a(T(0), T(0));
      fmac_(a, M[r], h[3 * s]);
      fmac_(a, M[3 + r], h[3 * s + 1]);
      fmac_(a, M[6 + r], h[3 * s + 2]);
      g[3 * s + r] = a;
    }
}

template <int MU, class T> HD void hop_pair(const FineOp<T> &op, long s, unsigned mask, const cx<T> *in, cx<T> *out) {
  const Lay ls = {12, op.sh}, lu = {36, op.sh};
  if (mask & (1u << MU)) {             // forward: (1-gamma) D_mu(x) psi(x+mu)
    long n = op.nb[(long)MU * op.V + s];
    cx<T> p[12], h[6], g[6], M[9];
#pragma unroll
    for (int c = 0; c < 12; c++) p[c] = in[ls.idx(n, c)];
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = op.D[lu.idx(s, 9 * MU + k)];
    project<MU, +1>(p, h);
    su3_mul(M, h, g);
    reconstruct_sub<MU, +1>(g, out);
  }
  if (mask & (1u << (4 + MU))) {       // backward: (1+gamma) D_mu(x-mu)^dagger psi(x-mu)
    long n = op.nb[(long)(4 + MU) * op.V + s];
    cx<T> p[12], h[6], g[6], M[9];
#pragma unroll
    for (int c = 0; c < 12; c++) p[c] = in[ls.idx(n, c)];
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = op.D[lu.idx(n, 9 * MU + k)];
    project<MU, -1>(p, h);
    su3_mul_dag(M, h, g);
    reconstruct_sub<MU, -1>(g, out);
  }
}

// y = Cl x for the packed Hermitian 2 x (6x6) block matrix stored at site s of array C (Lay{72,sh})
template <class T> HD void clover_mul(const T *C, int sh, long s, const cx<T> *x, cx<T> *y) {
  const Lay lc = {72, sh};
#pragma unroll
  for (int b = 0; b < 2; b++) {
#pragma unroll
    for (int i = 0; i < 6; i++) y[6 * b + i] = C[lc.idx(s, 6 * b + i)] * x[6 * b + i];
    int m = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = i + 1; j < 6; j++, m++) {
        cx<T> cij(C[lc.idx(s, 12 + 2 * (15 * b + m))], C[lc.idx(s, 12 + 2 * (15 * b + m) + 1)]);
        fma_(y[6 * b + i], cij, x[6 * b + j]);
        fmac_(y[6 * b + j], cij, x[6 * b + i]);
      }
  }
}

}  // namespace dda

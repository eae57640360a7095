// dev_gmres.h -- device-resident restarted GMRES for the coarsest level: the Hessenberg matrix, the Givens rotations,
// the convergence test and the back substitution live on the DEVICE, so an Arnoldi step is a fixed sequence of
// launches without any host synchronisation.  The host only decides how many steps to enqueue: it polls a 4-int control
// block after the number of steps the previous solve needed (then every few steps), and the kernels of steps enqueued
// past convergence return at once (the `skip` flag handed to the operator).
//
// Arithmetic and control flow are those of Fgmres<float> (krylov.h), i.e. the reference's fgmres_PRECISION /
// arnoldi_step / qr_update / compute_solution (linsolve_generic.c:219-413, 809-982): classical Gram-Schmidt with all
// inner products of a step in one reduction, a second reduction for the norm, double-precision scalars.  With several
// ranks the reductions are NCCL all-reduces on the device buffer (every rank sees identical values, hence identical
// control flow).
#pragma once
#include "common.cuh"
#include "blas.h"

namespace dda {

// layout of the device scalars of DevGmres (doubles): Hessenberg matrix (column-major, m+1 rows), gamma, Givens c / s,
// y, the inner products of the current step, then the norm slot and a few scalars behind it
struct GmresOff { long H, G, C, S, Y, HB, N; int m; };
inline GmresOff gmres_offsets(int m) {
  GmresOff o; o.m = m;
  o.H = 0; o.G = o.H + 2L * (m + 1) * m; o.C = o.G + 2L * (m + 1); o.S = o.C + 2L * m; o.Y = o.S + 2L * m;
  o.HB = o.Y + 2L * m; o.N = o.HB + 2L * (m + 1);
  return o;
}
enum { GM_INVSCALE = 1, GM_NORM_R0 = 2, GM_RELRES = 3, GM_SCALARS = 8 };   // st[o.N] = ||.||^2 of the last norm reduction

// Hessenberg column j from the inner products st[o.HB ..] and the norm st[o.N], Givens rotations, convergence test
// (arnoldi_step / qr_update, linsolve_generic.c:809-940).  One thread.  Shared by the generic kernel and the fused one.
HD void gmres_givens(double *S, int *ct, const GmresOff &o, int j, double tl) {
  if (ct[0]) return;
  const int M1 = o.m + 1;
  double *H = S + o.H + 2L * j * M1;            // column j
  for (int i = 0; i <= j; i++) { H[2 * i] = S[o.HB + 2 * i]; H[2 * i + 1] = S[o.HB + 2 * i + 1]; }
  const double hn = sqrt(S[o.N]);
  H[2 * (j + 1)] = hn; H[2 * (j + 1) + 1] = 0.0;
  ct[1] += 1; ct[2] = j + 1;
  S[o.N + GM_INVSCALE] = hn > 1e-15 ? 1.0 / hn : 0.0;
  if (hn > tl / 10) {
    for (int i = 0; i < j; i++) {
      const cd ci(S[o.C + 2 * i], S[o.C + 2 * i + 1]), si(S[o.S + 2 * i], S[o.S + 2 * i + 1]);
      const cd h0(H[2 * i], H[2 * i + 1]), h1(H[2 * i + 2], H[2 * i + 3]);
      const cd beta = (-si) * h0 + ci * h1, hh = conj(ci) * h0 + conj(si) * h1;
      H[2 * i] = hh.re; H[2 * i + 1] = hh.im; H[2 * i + 2] = beta.re; H[2 * i + 3] = beta.im;
    }
    const cd hj(H[2 * j], H[2 * j + 1]), hj1(H[2 * j + 2], H[2 * j + 3]);
    const double bn = sqrt(norm2(hj) + norm2(hj1));
    const cd sj(hj1.re / bn, hj1.im / bn), cj(hj.re / bn, hj.im / bn);
    S[o.S + 2 * j] = sj.re; S[o.S + 2 * j + 1] = sj.im; S[o.C + 2 * j] = cj.re; S[o.C + 2 * j + 1] = cj.im;
    const cd gj(S[o.G + 2 * j], S[o.G + 2 * j + 1]);
    const cd gj1 = (-sj) * gj, gjn = conj(cj) * gj;
    S[o.G + 2 * (j + 1)] = gj1.re; S[o.G + 2 * (j + 1) + 1] = gj1.im; S[o.G + 2 * j] = gjn.re; S[o.G + 2 * j + 1] = gjn.im;
    H[2 * j] = bn; H[2 * j + 1] = 0.0; H[2 * j + 2] = 0.0; H[2 * j + 3] = 0.0;
    const double rel = sqrt(norm2(gj1)) / S[o.N + GM_NORM_R0];
    S[o.N + GM_RELRES] = rel;
    if (rel < tl || rel > 1e5) { ct[0] = 1; ct[3] = 1; }
  } else { ct[0] = 1; ct[3] = 1; }
}

struct DevGmres {
  long n = 0, stride = 0;
  int m = 0, max_restart = 0;
  double tol = 0;
  bool allocated = false;
  bool reduce_over_ranks = true;   // false: every rank holds the WHOLE vectors (gathered coarsest level): local reductions are global
  cf *V = nullptr;        // m+1 basis vectors, stride `stride` (>= n: room for ghost slabs)
  cf *w = nullptr;
  double *st = nullptr;   // device scalars, see offsets in dev_gmres.cu
  int *ctrl = nullptr;    // device: [0] current restart cycle finished, [1] iterations, [2] valid Arnoldi columns, [3] solve finished
  // out = A in; `skip` (device pointer, may be null): the kernels may return immediately when *skip != 0
  std::function<void(cf *out, const cf *in, const int *skip)> op;
  // optional fused steps (sm_100a kernels of the coarsest level; both or none):
  //   op_dots(w, v_j, j): w = A v_j AND st[HB + 2k ..] += <V_k, w> for k <= j (the buffer is zero on entry)
  //   axpy_givens(j):     w -= sum_k h_k V_k, ||w||^2, then gmres_givens by the last CTA, which also clears the buffers
  std::function<void(cf *w, const cf *vj, int j)> op_dots;
  std::function<void(int j)> axpy_givens;
  std::function<bool()> fused_gate;       // optional: the fused steps are used only while this returns true
  int last_iter = 0, predicted = 8;
  long polls = 0;
  double last_relres = 0;

  void alloc(long n_, int m_, int max_restart_, double tol_, long nalloc_);
  void release();
  int solve(cf *x, const cf *b);   // zero initial guess; returns the number of iterations
  cf *basis(int k) const { return V + (long)k * stride; }
};

}  // namespace dda

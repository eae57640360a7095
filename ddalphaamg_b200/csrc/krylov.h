// krylov.h -- restarted, right-preconditioned flexible GMRES driven from the host, vectors on the device.
// Behaviour follows the reference's fgmres_PRECISION (linsolve_generic.c:219-413), arnoldi_step (default variant,
// :809-895), qr_update (:898-940) and compute_solution (:943-982).  One fused multi-dot launch + one norm launch
// per Arnoldi step; Hessenberg / Givens scalars live on the host in double.
#pragma once
#include "common.cuh"
#include "blas.h"

namespace dda {

template <class T> struct Fgmres {
  typedef cx<T> C;
  long n = 0;
  int m = 0, max_restart = 0;
  double tol = 0;
  bool flexible = false, allocated = false;
  // opt-in (env DDA_SINGLE_REDUCTION=1, coarse-level solvers only): one fused reduction per Arnoldi step, the norm of the
  // new direction from ||w||^2 - sum |h_i|^2 like the reference's SINGLE_ALLREDUCE_ARNOLDI variant
  // (linsolve_generic.c:773-805); falls back to an explicit norm when the subtraction cancels
  bool single_reduction = false;
  std::vector<C *> V, Z;
  C *w = nullptr, *r = nullptr;
  std::vector<cd> H, gamma, c, s, y;   // H column-major: H[j*(m+1)+i]
  std::function<void(C *, const C *)> op, prec;
  int last_iter = 0;
  double last_relres = 0;

  // n_: local vector length the solver works on; nalloc_ >= n_: allocated length (ghost slabs of a partitioned level)
  void alloc(long n_, int m_, int max_restart_, double tol_, bool flexible_, long nalloc_ = 0) {
    release();
    n = n_; m = m_; max_restart = max_restart_; tol = tol_; flexible = flexible_;
    const long na = nalloc_ > n_ ? nalloc_ : n_;
    V.resize(m + 1); for (auto &v : V) v = dev_alloc<C>(na);
    if (flexible) { Z.resize(m); for (auto &z : Z) z = dev_alloc<C>(na); }
    w = dev_alloc<C>(na); r = dev_alloc<C>(na);
    H.assign((size_t)(m + 1) * m, cd(0, 0)); gamma.assign(m + 1, cd(0, 0)); c.assign(m, cd(0, 0)); s.assign(m, cd(0, 0)); y.assign(m, cd(0, 0));
    allocated = true;
  }
  void release() {
    for (auto v : V) dev_free(v);
    for (auto z : Z) dev_free(z);
    V.clear(); Z.clear();
    dev_free(w); dev_free(r); w = r = nullptr; allocated = false;
  }
  cd &h(int i, int j) { return H[(size_t)j * (m + 1) + i]; }

  // solves op x = b.  zero_guess: x is taken as 0 on entry.  Returns the number of iterations.
  int solve(C *x, const C *b, bool zero_guess) {
    DDA_ASSERT(allocated && op);
    int iter = 0, finish = 0, j = -1;
    double norm_r0 = 1, gamma_jp1 = 1;
    for (int ol = 0; ol < max_restart && !finish; ol++) {
      bool nores = (ol == 0 && zero_guess);
      if (nores) vcopy(r, b, n);
      else { op(w, x); vsub(r, b, w, n); }
      double g0 = std::sqrt(vnorm2(r, n));
      gamma[0] = cd(g0, 0);
      if (ol == 0) norm_r0 = g0;
      if (g0 == 0.0) { if (nores) vzero(x, n); last_relres = 0; break; }
      vscale(V[0], r, 1.0 / g0, n);
      j = -1;
      for (int il = 0; il < m && !finish; il++) {
        j = il; iter++;
        const C *zj = V[j];
        if (prec) { prec(Z[j], V[j]); zj = Z[j]; }
        op(w, zj);
        std::vector<cd> hcol(j + 2);
        double hn;
        if (single_reduction && j + 2 < 64) {
          vmulti_dot_norm(hcol.data(), V.data(), j + 1, w, n);          // hcol[j+1] = <w,w>
          const double ww = hcol[j + 1].re;
          double hh = 0;
          for (int i = 0; i <= j; i++) hh += norm2(hcol[i]);
          vmulti_axpy(w, V.data(), hcol.data(), j + 1, -1, n);
          hn = (ww - hh > 1e-4 * ww) ? std::sqrt(ww - hh) : std::sqrt(vnorm2(w, n));
        } else {
          vmulti_dot(hcol.data(), V.data(), j + 1, w, n);
          hn = std::sqrt(vmulti_axpy_norm2(w, V.data(), hcol.data(), j + 1, -1, n));
        }
        for (int i = 0; i <= j; i++) h(i, j) = hcol[i];
        h(j + 1, j) = cd(hn, 0);
        if (hn > 1e-15) vscale(V[j + 1], w, 1.0 / hn, n);
        if (hn > tol / 10) {
          // Givens update
          for (int i = 0; i < j; i++) {
            cd beta = (-s[i]) * h(i, j) + c[i] * h(i + 1, j);
            h(i, j) = conj(c[i]) * h(i, j) + conj(s[i]) * h(i + 1, j);
            h(i + 1, j) = beta;
          }
          double bn = std::sqrt(norm2(h(j, j)) + norm2(h(j + 1, j)));
          s[j] = cd(h(j + 1, j).re / bn, h(j + 1, j).im / bn); c[j] = cd(h(j, j).re / bn, h(j, j).im / bn);
          gamma[j + 1] = (-s[j]) * gamma[j]; gamma[j] = conj(c[j]) * gamma[j];
          h(j, j) = cd(bn, 0); h(j + 1, j) = cd(0, 0);
          gamma_jp1 = std::sqrt(norm2(gamma[j + 1]));
          if (gamma_jp1 / norm_r0 < tol || gamma_jp1 / norm_r0 > 1e5) {
            finish = 1;
            if (gamma_jp1 / norm_r0 > 1e5) fprintf(stderr, "dd_alpha_amg_b200: divergence of fgmres, iter = %d\n", iter);
          }
        } else { finish = 1; break; }
      }
      // back substitution + solution update (x = or += sum y_i Z_i)
      if (j >= 0) {
        for (int i = j; i >= 0; i--) {
          cd yi = gamma[i];
          for (int k = i + 1; k <= j; k++) yi -= h(i, k) * y[k];
          double d = norm2(h(i, i)); cd hi = h(i, i);
          y[i] = cd((yi.re * hi.re + yi.im * hi.im) / d, (yi.im * hi.re - yi.re * hi.im) / d);
        }
        std::vector<C *> &B = prec ? Z : V;
        if (nores) vzero(x, n);
        vmulti_axpy(x, B.data(), y.data(), j + 1, +1, n);
      }
      last_relres = gamma_jp1 / norm_r0;
    }
    last_iter = iter;
    return iter;
  }
};

// Mixed-precision FGMRES ("mixed precision: 2"): restarts, true residual and solution in double; Arnoldi basis, operator
// and preconditioner in float with double-accumulated inner products; an inner cycle ends after a residual reduction of
// max(tol, 1e-5).  Behaviour follows fgmres_MP / arnoldi_step_MP / compute_solution_MP (linsolve.c:153-424) and
// fgmres_MP_struct_alloc (linsolve.c:31-50).  The float vectors are already in the level's native order, so the
// reference's trans / trans_back permutations reduce to precision casts.
struct FgmresMP {
  long n = 0;
  int m = 0, max_restart = 0;
  double tol = 0, sp_tol = 0;
  bool allocated = false;
  std::vector<cf *> V, Z;
  cf *w = nullptr;
  cd *r = nullptr;
  std::vector<cd> H, gamma, c, s, y;
  std::function<void(cd *, const cd *)> op_d;
  std::function<void(cf *, const cf *)> op_f, prec_f;
  int last_iter = 0;
  double last_relres = 0;

  void alloc(long n_, int m_, int max_restart_, double tol_, bool flexible, long nalloc_ = 0) {
    release();
    n = n_; m = m_; max_restart = max_restart_; tol = tol_; sp_tol = std::max(tol_, 1e-5);
    const long na = nalloc_ > n_ ? nalloc_ : n_;
    V.resize(m + 1); for (auto &v : V) v = dev_alloc<cf>(na);
    if (flexible) { Z.resize(m); for (auto &z : Z) z = dev_alloc<cf>(na); }
    w = dev_alloc<cf>(na); r = dev_alloc<cd>(na);
    H.assign((size_t)(m + 1) * m, cd(0, 0)); gamma.assign(m + 1, cd(0, 0)); c.assign(m, cd(0, 0)); s.assign(m, cd(0, 0)); y.assign(m, cd(0, 0));
    allocated = true;
  }
  void release() {
    for (auto v : V) dev_free(v);
    for (auto z : Z) dev_free(z);
    V.clear(); Z.clear();
    dev_free(w); dev_free(r); w = nullptr; r = nullptr; allocated = false;
  }
  cd &h(int i, int j) { return H[(size_t)j * (m + 1) + i]; }

  int solve(cd *x, const cd *b, bool zero_guess) {
    DDA_ASSERT(allocated && op_d && op_f);
    sp_tol = std::max(tol, 1e-5);
    int iter = 0, finish = 0, j = -1;
    double norm_r0 = 1, gamma_jp1 = 1;
    for (int ol = 0; ol < max_restart && !finish; ol++) {
      if (ol == 0 && zero_guess) vcopy(r, b, n);
      else { op_d(r, x); vsub(r, b, r, n); }
      const double g0 = std::sqrt(vnorm2(r, n));
      gamma[0] = cd(g0, 0);
      if (ol == 0) norm_r0 = g0;
      if (g0 == 0.0) { if (ol == 0 && zero_guess) vzero(x, n); last_relres = 0; break; }
      vcast(V[0], r, n);
      vscale(V[0], V[0], 1.0 / g0, n);
      j = -1;
      for (int il = 0; il < m && !finish; il++) {
        j = il; iter++;
        const cf *zj = V[j];
        if (prec_f) { prec_f(Z[j], V[j]); zj = Z[j]; }
        op_f(w, zj);
        std::vector<cd> hcol(j + 2);
        vmulti_dot(hcol.data(), V.data(), j + 1, w, n);
        for (int i = 0; i <= j; i++) h(i, j) = hcol[i];
        const double hn = std::sqrt(vmulti_axpy_norm2(w, V.data(), hcol.data(), j + 1, -1, n));
        h(j + 1, j) = cd(hn, 0);
        if (hn > 1e-15) {
          vscale(V[j + 1], w, 1.0 / hn, n);
          for (int i = 0; i < j; i++) {
            cd beta = (-s[i]) * h(i, j) + c[i] * h(i + 1, j);
            h(i, j) = conj(c[i]) * h(i, j) + conj(s[i]) * h(i + 1, j);
            h(i + 1, j) = beta;
          }
          const double bn = std::sqrt(norm2(h(j, j)) + norm2(h(j + 1, j)));
          s[j] = cd(h(j + 1, j).re / bn, h(j + 1, j).im / bn); c[j] = cd(h(j, j).re / bn, h(j, j).im / bn);
          gamma[j + 1] = (-s[j]) * gamma[j]; gamma[j] = conj(c[j]) * gamma[j];
          h(j, j) = cd(bn, 0); h(j + 1, j) = cd(0, 0);
          gamma_jp1 = std::sqrt(norm2(gamma[j + 1]));
          if (gamma_jp1 / norm_r0 < tol || gamma_jp1 / norm_r0 > 1e5) {
            finish = 1;
            if (gamma_jp1 / norm_r0 > 1e5) fprintf(stderr, "dd_alpha_amg_b200: divergence of fgmres_MP, iter = %d\n", iter);
          }
          if (gamma_jp1 / g0 < sp_tol) break;
        } else { finish = 1; }
      }
      if (j >= 0) {
        for (int i = j; i >= 0; i--) {
          cd yi = gamma[i];
          for (int k = i + 1; k <= j; k++) yi -= h(i, k) * y[k];
          const double d = norm2(h(i, i)); const cd hi = h(i, i);
          y[i] = cd((yi.re * hi.re + yi.im * hi.im) / d, (yi.im * hi.re - yi.re * hi.im) / d);
        }
        std::vector<cf *> &B = prec_f ? Z : V;
        vzero(w, n);
        vmulti_axpy(w, B.data(), y.data(), j + 1, +1, n);
        vcast(r, w, n);
        // the reference assigns in the first restart cycle whatever the initial guess was (linsolve.c fgmres_MP); a caller's
        // non-zero guess must be kept, so only a zero guess is overwritten
        if (ol == 0 && zero_guess) vcopy(x, r, n); else vadd(x, x, r, n);
      }
      last_relres = gamma_jp1 / norm_r0;
    }
    last_iter = iter;
    return iter;
  }
};

}  // namespace dda

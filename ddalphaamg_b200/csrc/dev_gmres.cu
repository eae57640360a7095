// dev_gmres.cu -- see dev_gmres.h.  All kernels here are generic (launch_n / launch_reduce), so the host-emulation build
// runs the very same device-side Givens / convergence logic in the CPU tests.
#include "dev_gmres.h"
#include "comm.h"

namespace dda {

namespace {
typedef GmresOff Off;
inline Off offsets(int m) { return gmres_offsets(m); }
enum { SC_INVSCALE = GM_INVSCALE, SC_NORM_R0 = GM_NORM_R0, SC_RELRES = GM_RELRES, SC_TOTAL = GM_SCALARS };

HD cd ldc2(const double *p) { return cd(p[0], p[1]); }
HD void stc2(double *p, cd v) { p[0] = v.re; p[1] = v.im; }
}  // namespace

void DevGmres::alloc(long n_, int m_, int max_restart_, double tol_, long nalloc_) {
  release();
  n = n_; m = m_; max_restart = max_restart_; tol = tol_;
  stride = nalloc_ > n_ ? nalloc_ : n_;
  V = dev_alloc<cf>((size_t)(m + 1) * stride);
  w = dev_alloc<cf>(stride);
  const Off o = offsets(m);
  st = dev_alloc<double>(o.N + SC_TOTAL);
  ctrl = dev_alloc<int>(4);
  dev_zero(V, sizeof(cf) * (size_t)(m + 1) * stride); dev_zero(w, sizeof(cf) * stride);
  dev_zero(st, sizeof(double) * (o.N + SC_TOTAL)); dev_zero(ctrl, sizeof(int) * 4);
  predicted = 8;
  allocated = true;
}

void DevGmres::release() {
  dev_free(V); dev_free(w); dev_free(st); dev_free(ctrl);
  V = w = nullptr; st = nullptr; ctrl = nullptr; allocated = false;
}

int DevGmres::solve(cf *x, const cf *b) {
  DDA_ASSERT(allocated && op);
  const Off o = offsets(m);
  double *S = st; int *ct = ctrl;
  cf *Vb = V, *ww = w;
  const long nn = n, sd = stride;
  const double tl = tol;
  int hc[4] = {0, 0, 0, 0};
  int first_cycle_iters = -1;
  for (int ol = 0; ol < max_restart; ol++) {
    // r = b - A x (r = b in the first cycle: zero initial guess), kept unnormalised in w
    if (ol == 0) vcopy(ww, b, nn);
    else { op(ww, x, nullptr); launch_n(nn, DLAMBDA(long i) { ww[i] = b[i] - ww[i]; }); }
    launch_reduce<1>(1, nn, DLAMBDA(long seg, long i, double *acc) { (void)seg; cf a = ww[i]; acc[0] += (double)a.re * a.re + (double)a.im * a.im; }, S + o.N);
    if (reduce_over_ranks) comm_allreduce_sum(S + o.N, 1);
    launch_n(1, DLAMBDA(long) {
      const double g0 = sqrt(S[o.N]);
      S[o.G] = g0; S[o.G + 1] = 0.0;
      if (ol == 0) { S[o.N + SC_NORM_R0] = g0; ct[1] = 0; }
      ct[0] = 0; ct[2] = 0; ct[3] = 0;
      if (g0 == 0.0) { ct[0] = 1; ct[3] = 1; S[o.N + SC_RELRES] = 0.0; S[o.N + SC_INVSCALE] = 0.0; }
      else S[o.N + SC_INVSCALE] = 1.0 / g0;
      for (int i = 0; i < 2 * (o.m + 1); i++) S[o.HB + i] = 0.0;      // the fused step kernels accumulate into these
      S[o.N] = 0.0;
    });
    launch_n(nn, DLAMBDA(long i) { const float f = (float)S[o.N + SC_INVSCALE]; Vb[i] = f * ww[i]; });
    int next_poll = std::min(m, std::max(1, predicted));
    bool done = false;
    for (int j = 0; j < m && !done; j++) {
      const cf *vj = Vb + (long)j * sd;
      if (op_dots && axpy_givens && (!fused_gate || fused_gate())) {
        op_dots(ww, vj, j);
        axpy_givens(j);
      } else {
      op(ww, vj, ct);
      // hbuf[k] = <V_k, w>, k <= j   (process_multi_inner_product, linalg_generic.c:107-154)
      launch_reduce<2>(j + 1, nn, DLAMBDA(long seg, long i, double *acc) {
        const cf a = Vb[seg * sd + i], c = ww[i];
        acc[0] += (double)a.re * c.re + (double)a.im * c.im;
        acc[1] += (double)a.re * c.im - (double)a.im * c.re;
      }, S + o.HB);
      if (reduce_over_ranks) comm_allreduce_sum(S + o.HB, 2 * (j + 1));
      // w -= sum_k hbuf[k] V_k and ||w||^2 in one pass
      launch_reduce<1>(1, nn, DLAMBDA(long seg, long i, double *acc) {
        (void)seg;
        cf v = ww[i];
        for (int k = 0; k <= j; k++) fms_(v, cf((float)S[o.HB + 2 * k], (float)S[o.HB + 2 * k + 1]), Vb[(long)k * sd + i]);
        ww[i] = v;
        acc[0] += (double)v.re * v.re + (double)v.im * v.im;
      }, S + o.N);
      if (reduce_over_ranks) comm_allreduce_sum(S + o.N, 1);
      // Hessenberg column, Givens rotations, convergence test (qr_update, linsolve_generic.c:898-940)
      launch_n(1, DLAMBDA(long) { gmres_givens(S, ct, o, j, tl); });
      }
      if (j + 1 < m) {
        cf *vn = Vb + (long)(j + 1) * sd;
        launch_n(nn, DLAMBDA(long i) { if (ct[0]) return; const float f = (float)S[o.N + SC_INVSCALE]; vn[i] = f * ww[i]; });
      }
      if (j + 1 == next_poll || j + 1 == m) {
        d2h(hc, ct, sizeof(int) * 4); polls++;
        done = hc[0] != 0;
        next_poll += 4;
      }
    }
    if (first_cycle_iters < 0) first_cycle_iters = hc[2];
    // back substitution and x (+)= sum_i y_i V_i   (compute_solution, linsolve_generic.c:943-982)
    launch_n(1, DLAMBDA(long) {
      const int jc = ct[2], M1 = o.m + 1;
      for (int i = jc - 1; i >= 0; i--) {
        cd yi = ldc2(S + o.G + 2 * i);
        for (int k = i + 1; k < jc; k++) yi -= ldc2(S + o.H + 2L * k * M1 + 2 * i) * ldc2(S + o.Y + 2 * k);
        const cd hi = ldc2(S + o.H + 2L * i * M1 + 2 * i); const double d = norm2(hi);
        stc2(S + o.Y + 2 * i, cd((yi.re * hi.re + yi.im * hi.im) / d, (yi.im * hi.re - yi.re * hi.im) / d));
      }
    });
    const bool firstx = (ol == 0);
    launch_n(nn, DLAMBDA(long i) {
      const int jc = ct[2];
      cf v = firstx ? cf(0.f, 0.f) : x[i];
      for (int k = 0; k < jc; k++) fma_(v, cf((float)S[o.Y + 2 * k], (float)S[o.Y + 2 * k + 1]), Vb[(long)k * sd + i]);
      x[i] = v;
    });
    if (hc[3]) break;
  }
  last_iter = hc[1];
  if (first_cycle_iters > 0) predicted = std::min(m, std::max(4, first_cycle_iters));
  return hc[1];
}

}  // namespace dda

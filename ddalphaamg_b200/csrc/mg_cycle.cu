// mg_cycle.cu -- the solve phase: SAP smoother, V-/K-cycle, coarsest-level even-odd solve, preconditioned outer
// FGMRES.  Generic kernels only (launch_n / launch_reduce); the hand-tuned sm_100a kernels are selected inside
// solver_apply_dw / mg_apply_op / mg_smoother when Solver::use_fast is set.
//
// Reference counterparts: red_black_schwarz_PRECISION (schwarz_generic.c:1260-1431), block_solve_oddeven_PRECISION
// (oddeven_generic.c:1332-1360), apply_block_schur_complement (:1317-1329), local_minres_PRECISION
// (linsolve_generic.c:985-1029), vcycle_PRECISION / smoother_PRECISION (vcycle_generic.c:25-141), preconditioner
// (preconditioner.c:25-69), coarse_solve_odd_even / coarse_apply_schur_complement (coarse_oddeven_generic.c:1139-1189),
// fgmres_PRECISION (linsolve_generic.c:219-413), wilson_driver (top_level.c:64-104).
#include "solver.h"
#include "halo.h"
#include <chrono>

namespace dda {

void lv_halo(Level &L, const cf *v) { halo_exchange<cf>(L.geo, const_cast<cf *>(v), L.geo.nc, L.geo.sh); }

void lv_apply(Level &L, cf *out, const cf *in, SiteSel sel, int hop, int dir, int self, int outmode, const cf *eta,
              const cf *in_self) {
  if (L.depth == 0) fine_apply<float>(L.opf, out, in, sel, hop, dir, self, outmode, eta, in_self);
  else coarse_apply(L.cop, out, in, sel, hop, dir, self, outmode, eta, in_self);
}


void mg_apply_op(Solver &s, int depth, cf *out, const cf *in) {
  Level &L = s.lev[depth];
  ProfScope ps(s, &s.t_op[depth]);
  if (depth == 0) { solver_apply_dw<float>(s, out, in); return; }
  lv_halo(L, in);
#ifndef DDA_HOST_EMU
  if (s.use_fast && coarse_apply_fast(L.cop, out, in, L.copZ)) return;
#endif
  lv_apply(L, out, in, sel_all(L.geo.V), HOP_ALL, 0, SELF_C, OUT_SET);
}

// ---------------------------------------------------------------------------------------------------------------
// SAP: multiplicative (red-black) Schwarz.  For colour c = 0,1:  x_c += B_c^{-1} (eta - D x)_c  where B_c^{-1} is
// the approximate block inverse: block_iter minimal-residual steps, on the fine level applied to the even-odd Schur
// complement of the block operator.  The block residual is recomputed from x (gather form) instead of being updated
// incrementally as in the reference; the iterates are the same up to rounding.
// per-block MR step: alpha = <Dr,r>/<Dr,Dr>; lphi += alpha r; r -= alpha Dr      (local_minres, linsolve_generic.c:1013-1022)
static void block_mr_step(Level &L, const int *list, int nblk, int cnt, cf *lphi, cf *r, const cf *Dr, bool first) {
  const Geometry &g = L.geo;
  const Lay lay = g.lay(); const int nc = g.nc, bs = g.bs;
  double *red = L.blockred;
  launch_reduce<3>(nblk, (long)cnt * nc, DLAMBDA(long seg, long i, double *acc) {
    long sl; int c;
    if ((cnt & 31) == 0) sel_decode(cnt, nc, i, sl, c); else { sl = i / nc; c = (int)(i - sl * nc); }
    long q = lay.idx((long)list[seg] * bs + sl, c);
    cf a = Dr[q], b = r[q];
    acc[0] += (double)a.re * b.re + (double)a.im * b.im;
    acc[1] += (double)a.re * b.im - (double)a.im * b.re;
    acc[2] += (double)a.re * a.re + (double)a.im * a.im;
  }, red);
  long n = (long)nblk * cnt;
  launch_n(sel_threads(n, nc), DLAMBDA(long i) {
    long si; int c;
    if (!sel_decode(n, nc, i, si, c)) return;
    long b = si / cnt; long sl = si - b * cnt;
    long q = lay.idx((long)list[b] * bs + sl, c);
    double den = red[3 * b + 2];
    cf alpha(0.f, 0.f);
    if (den > 1e-30) alpha = cf((float)(red[3 * b] / den), (float)(red[3 * b + 1] / den));
    cf rv = r[q];
    cf lp = first ? cf(0.f, 0.f) : lphi[q];
    fma_(lp, alpha, rv); lphi[q] = lp;
    fms_(rv, alpha, Dr[q]); r[q] = rv;
  });
}

void sap_fine_fast(Solver &s, cf *x, const cf *eta, int iters, bool zero_guess);
bool sap_fine_fast_available(const Solver &s);

void mg_smoother(Solver &s, int depth, cf *phi, const cf *eta, int iters, bool zero_guess) {
  Level &L = s.lev[depth];
  ProfScope ps(s, &s.t_smooth[depth]);
  const Geometry &g = L.geo;
#ifndef DDA_HOST_EMU
  if (depth == 0 && s.use_fast && sap_fine_fast_available(s)) { sap_fine_fast(s, phi, eta, iters, zero_guess); return; }
#endif
  const int nc = g.nc, bs = g.bs, be = g.bs_even, bo = g.bs - g.bs_even;
  const long n = g.vlen();
  const int biter = s.p.block_iter[depth];
  cf *x = phi;
  if (zero_guess) vzero(x, n);
  cf *r = L.w[0], *e = L.w[1], *t = L.w[2], *a = L.w[3], *Dr = L.w[4], *a2 = L.w[5];
  const bool eo = g.block_eo;
  for (int cyc = 0; cyc < iters; cyc++) {
    for (int col = 0; col < 2; col++) {
      const int *list = g.d_blocklist[col]; const int nblk = g.nblk_color[col];
      if (nblk == 0) continue;
      SiteSel sb = sel_blocks(list, nblk, bs, 0, bs);
      // block residual r = eta - D x on the blocks of this colour
      if (zero_guess && cyc == 0 && col == 0) {
        const Lay lay = g.lay(); long ns = (long)nblk * bs;
        launch_n(sel_threads(ns, nc), DLAMBDA(long i) {
          long si; int c; if (!sel_decode(ns, nc, i, si, c)) return;
          long b = si / bs; long q = lay.idx((long)list[b] * bs + (si - b * bs), c);
          r[q] = eta[q];
        });
      } else {
        lv_halo(L, x);
#ifndef DDA_HOST_EMU
        // coarse levels: the scatter-form full-lattice kernel reads 5 blocks per site, the masked gather kernel 9 per
        // selected site -- the full apply costs the same traffic as the half-lattice gather and runs at 2x the rate
        if (depth > 0 && s.use_fast && coarse_apply_fast(L.cop, Dr, x, L.copZ)) vsub(r, eta, Dr, n);
        else
#endif
        lv_apply(L, r, x, sb, HOP_ALL, 0, SELF_C, OUT_ETA_MINUS, eta);
      }
      if (eo) {
        SiteSel se = sel_blocks(list, nblk, bs, 0, be), so = sel_blocks(list, nblk, bs, be, bo);
        // odd -> even: e_o = Doo^-1 r_o ; t_e = r_e - Deo e_o
        lv_apply(L, e, r, so, HOP_NONE, 0, SELF_CINV, OUT_SET);
        lv_apply(L, t, e, se, HOP_INBLOCK, 0, SELF_NONE, OUT_ETA_MINUS, r);
        // MR on the Schur complement S = Dee - Deo Doo^-1 Doe ; solution accumulates in e_e, residual in t_e
        for (int it = 0; it < biter; it++) {
          // lv_apply's hop part is the operator's off-diagonal part N (sign included): D = C + N
          lv_apply(L, a, t, so, HOP_INBLOCK, 0, SELF_NONE, OUT_SET);             // a_o  = N_oe t_e
          lv_apply(L, a2, a, so, HOP_NONE, 0, SELF_CINV, OUT_NEG);               // a2_o = -Doo^-1 a_o
          lv_apply(L, Dr, a2, se, HOP_INBLOCK, 0, SELF_C, OUT_SET, nullptr, t);  // Dr_e = C t_e + N_eo a2_o = S t_e
          block_mr_step(L, list, nblk, be, e, t, Dr, it == 0);
        }
        if (biter == 0) {
          const Lay lay = g.lay(); long ns = (long)nblk * be;
          launch_n(sel_threads(ns, nc), DLAMBDA(long i) {
            long si; int c; if (!sel_decode(ns, nc, i, si, c)) return;
            long b = si / be; long q = lay.idx((long)list[b] * bs + (si - b * be), c);
            e[q] = cf(0.f, 0.f);
          });
        }
        // even -> odd: e_o = Doo^-1 (r_o - Doe e_e)
        lv_apply(L, a, e, so, HOP_INBLOCK, 0, SELF_NONE, OUT_ETA_MINUS, r);
        lv_apply(L, e, a, so, HOP_NONE, 0, SELF_CINV, OUT_SET);
      } else {
        // plain MR on the block operator (coarse_block_operator, coarse_operator_generic.c:208-236)
#ifndef DDA_HOST_EMU
        if (depth > 0 && s.use_fast && biter > 0 &&
            coarse_sap_mr_fast(L.cop, x, r, list, nblk, bs, biter, g.d_sapjobs, g.nsapjobs)) continue;   // fused: MR steps and x += e
#endif
        for (int it = 0; it < biter; it++) {
          lv_apply(L, Dr, r, sb, HOP_INBLOCK, 0, SELF_C, OUT_SET);
          block_mr_step(L, list, nblk, bs, e, r, Dr, it == 0);
        }
        if (biter == 0) continue;
      }
      // x += e on the blocks of this colour
      {
        const Lay lay = g.lay(); long ns = (long)nblk * bs;
        launch_n(sel_threads(ns, nc), DLAMBDA(long i) {
          long si; int c; if (!sel_decode(ns, nc, i, si, c)) return;
          long b = si / bs; long q = lay.idx((long)list[b] * bs + (si - b * bs), c);
          x[q] += e[q];
        });
      }
    }
  }
  double rf = s.p.relax_fac[depth];
  if (rf != 1.0) vscale(phi, phi, rf, n);
}

// ---------------------------------------------------------------------------------------------------------------
void mg_vcycle(Solver &s, int depth, cf *phi, const cf *eta, bool zero_guess) {
  Level &L = s.lev[depth];
  if (L.last || s.p.interpolation == 0) {
    mg_smoother(s, depth, phi, eta, depth == 0 ? s.p.ncycle[0] : s.p.post_smooth_iter[depth], zero_guess);
    return;
  }
  Level &N = s.lev[depth + 1];
  const long n = L.geo.vlen();
  for (int i = 0; i < s.p.ncycle[depth]; i++) {
    if (i == 0 && zero_guess) {
      ProfScope ps(s, &s.t_restrict);
      tr_restrict(L.tr, N.vb, N.geo.nc, 0, eta, L.tr_scratch);
    } else {
      mg_apply_op(s, depth, L.w[5], phi);
      vsub(L.w[6], eta, L.w[5], n);
      ProfScope ps(s, &s.t_restrict);
      tr_restrict(L.tr, N.vb, N.geo.nc, 0, L.w[6], L.tr_scratch);
    }
    if (!N.last) {
      if (s.p.kcycle) N.kc.solve(N.vx, N.vb, true);
      else mg_vcycle(s, depth + 1, N.vx, N.vb, true);
    } else {
      mg_coarsest_solve(s);
    }
    {
      ProfScope ps(s, &s.t_interp);
      tr_interpolate(L.tr, phi, N.vx, !(i == 0 && zero_guess));
    }
    mg_smoother(s, depth, phi, eta, s.p.post_smooth_iter[depth], false);
    zero_guess = false;
  }
}

void mg_preconditioner(Solver &s, cd *out, const cd *in) {
  Level &L = s.lev[0];
  const long n = L.geo.vlen();
  if (s.p.method == 0 || s.nlev == 1) { vcopy(out, in, n); return; }
  vcast(L.vb, in, n);
  mg_vcycle(s, 0, L.vx, L.vb, true);
  vcast(out, L.vx, n);
}

double mg_solve(Solver &s, cd *x, const cd *b, double tol, int *status) {
  s.coarse_iter_count = 0;
  int it;
  if (s.outer_mp.allocated) { s.outer_mp.tol = tol; it = s.outer_mp.solve(x, b, true); }
  else { s.outer.tol = tol; it = s.outer.solve(x, b, true); }
  s.iter_count = it;
  // true residual (reference -DFGMRES_RESTEST, linsolve_generic.c:351-357)
  Level &L = s.lev[0];
  const long n = L.geo.vlen();
  cd *w = s.outer_mp.allocated ? s.outer_mp.r : s.outer.w;
  solver_apply_dw<double>(s, w, x);
  vsub(w, b, w, n);
  double nr = std::sqrt(vnorm2(w, n)), nb = std::sqrt(vnorm2(b, n));
  s.norm_res = nb > 0 ? nr / nb : nr;
  if (status) { status[0] = (s.norm_res > tol) ? -1 : it; status[1] = (int)s.coarse_iter_count; }
  return s.norm_res;
}

}  // namespace dda

// coarsest.cu -- the coarsest-level solve: even-odd preconditioned GMRES on the Schur complement of the coarsest
// operator, device resident (dev_gmres.h), on the gathered (replicated) coarsest lattice when the level is partitioned
// and small.
//
// Reference counterparts: coarse_solve_odd_even_PRECISION / coarse_apply_schur_complement_PRECISION
// (coarse_oddeven_generic.c:1139-1189), the fgmres call of the coarsest level (vcycle_generic.c:58-73), and the
// coarse-level gathering of gathering_generic.c:44-194, 285-346 (vector_PRECISION_gather / distribute).
#include "solver.h"
#include "halo.h"

namespace dda {

static long replicate_max_sites() {
  const char *e = getenv("DDA_COARSEST_REPLICATE_MAX");     // 0 switches the replication off
  return e ? atol(e) : 4096;
}

static void c_halo(Coarsest &C, const cf *v) { halo_exchange<cf>(*C.geo, const_cast<cf *>(v), C.geo->nc, 0); }

void coarsest_alloc(Solver &s) {
  Coarsest &C = s.cst;
  DDA_ASSERT(!C.active && s.nlev >= 2);
  Level &L = s.lev[s.nlev - 1];
  const Params &p = s.p;
  const int nc = L.geo.nc;
  long Vglob = 1;
  for (int m = 0; m < 4; m++) Vglob *= p.global_lattice[s.nlev - 1][m];
  C.replicated = p.odd_even && L.geo.partitioned() && Vglob <= replicate_max_sites();
  { const char *e = getenv("DDA_COARSEST_HOST"); C.host_driven = e && atoi(e) != 0; }
  if (C.replicated) {
    Geometry &g = C.rgeo;
    g = Geometry();
    for (int m = 0; m < 4; m++) { g.L[m] = p.global_lattice[s.nlev - 1][m]; g.B[m] = 0; g.A[m] = 0; }
    g.nc = nc; g.sh = 0; g.block_eo = false; g.global_eo = true;
    g.build();
    CoarseOp &c = C.rop;
    c = CoarseOp();
    c.n = nc; c.V = g.V; c.n_even = g.n_even;
    const long nn = (long)nc * nc;
    c.F = dev_alloc<cf>(g.V * 4 * nn); c.S = dev_alloc<cf>(g.V * nn); c.Sinv = dev_alloc<cf>((g.V - g.n_even) * nn);
    c.nb = g.d_nb; c.blkflag = g.d_blkflag; c.aggflag = g.d_aggflag;
    C.geo = &g; C.op = &c;
    // site maps between the ranks' local even-odd orders and the global one
    const Geometry &lg = L.geo;
    const long Vl = lg.V;
    std::vector<int> natidx[2];
    for (int po = 0; po < 2; po++) {
      natidx[po].assign(Vl, 0);
      long ne = 0;
      for (int t = 0; t < lg.L[0]; t++) for (int z = 0; z < lg.L[1]; z++) for (int y = 0; y < lg.L[2]; y++) for (int x = 0; x < lg.L[3]; x++)
        if (((t + z + y + x + po) & 1) == 0) ne++;
      long ce = 0, co = 0;
      for (int t = 0; t < lg.L[0]; t++) for (int z = 0; z < lg.L[1]; z++) for (int y = 0; y < lg.L[2]; y++) for (int x = 0; x < lg.L[3]; x++)
        natidx[po][lg.lex(t, z, y, x)] = (int)((((t + z + y + x + po) & 1) == 0) ? ce++ : ne + co++);
    }
    std::vector<int> src(g.V), own(Vl);
    for (long gi = 0; gi < g.V; gi++) {
      long lx = g.nat2lex[gi];
      int c4[4];
      for (int m = 3; m >= 0; m--) { c4[m] = (int)(lx % g.L[m]); lx /= g.L[m]; }
      int rc[4], lc[4], po = 0;
      for (int m = 0; m < 4; m++) { rc[m] = c4[m] / lg.L[m]; lc[m] = c4[m] % lg.L[m]; po += rc[m] * lg.L[m]; }
      const int r = ((rc[0] * lg.P[1] + rc[1]) * lg.P[2] + rc[2]) * lg.P[3] + rc[3];
      const int k = natidx[po & 1][lg.lex(lc[0], lc[1], lc[2], lc[3])];
      src[gi] = (int)(r * Vl + k);
      bool mine = true;
      for (int m = 0; m < 4; m++) mine = mine && rc[m] == lg.pc[m];
      if (mine) { DDA_ASSERT(k == lg.lex2nat[lg.lex(lc[0], lc[1], lc[2], lc[3])]); own[k] = (int)gi; }
    }
    C.d_src = dev_upload(src); C.d_own = dev_upload(own);
    C.gbuf = dev_alloc<cf>(g.V * nc);
  } else {
    C.geo = &L.geo; C.op = &L.cop;
  }
  const long na = C.geo->valloc();
  C.b = dev_alloc<cf>(na); C.x = dev_alloc<cf>(na);
  for (int i = 0; i < 4; i++) C.t[i] = dev_alloc<cf>(na);
#ifndef DDA_HOST_EMU
  C.fast = p.odd_even && !C.geo->partitioned() && schur_fast_supported(*C.op);
  { const char *e = getenv("DDA_SCHUR_FAST"); if (e && atoi(e) == 0) C.fast = false; }
  if (C.fast) { C.dir = dev_alloc<cf>(na); C.Z = dev_alloc<cf>(C.geo->V * 4 * nc); }
#endif
  const long nsolve = p.odd_even ? C.geo->n_even * nc : C.geo->vlen();
  Solver *sp = &s;
  if (C.host_driven) {
    C.hostk.alloc(nsolve, p.coarse_iter, p.coarse_restart, p.coarse_tol, false, na);
    if (p.odd_even) C.hostk.op = [sp](cf *out, const cf *in) { mg_coarsest_schur(*sp, out, in, nullptr); };
    else C.hostk.op = [sp](cf *out, const cf *in) { mg_apply_op(*sp, sp->nlev - 1, out, in); };
  } else {
    C.dg.alloc(nsolve, p.coarse_iter, p.coarse_restart, p.coarse_tol, na);
    C.dg.reduce_over_ranks = C.geo->partitioned();        // gathered / single-rank lattice: every rank has the whole vectors
    if (p.odd_even) C.dg.op = [sp](cf *out, const cf *in, const int *skip) { mg_coarsest_schur(*sp, out, in, skip); };
    else C.dg.op = [sp](cf *out, const cf *in, const int *) { mg_apply_op(*sp, sp->nlev - 1, out, in); };
#ifndef DDA_HOST_EMU
    bool fused = C.fast;
    { const char *e = getenv("DDA_GMRES_FUSED"); if (e && atoi(e) == 0) fused = false; }
    if (fused) {
      // Arnoldi step = 5 launches: the three streaming Schur kernels, the last Schur stage fused with the inner products,
      // orthogonalisation + norm + Givens (last CTA) -- instead of 10 with the generic reductions
      C.counter = dev_alloc<unsigned>(1);
      dev_zero(C.counter, sizeof(unsigned));
      C.dg.op_dots = [sp](cf *w, const cf *vj, int j) {
        Coarsest &K = sp->cst;
        const CoarseOp &o_ = *K.op; const int *skip = K.dg.ctrl;
        schur_hop(o_, 0, vj, nullptr, K.dir, K.Z, skip);
        schur_mid(o_, nullptr, K.dir, K.Z, K.t[1], 0.f, 1.f, -1.f, skip);
        schur_hop(o_, 1, K.t[1], vj, K.dir, K.Z, skip);
        schur_fin_dots(o_, K.dir, K.Z, w, K.dg.V, K.dg.stride, j, K.dg.st + gmres_offsets(K.dg.m).HB, skip);
      };
      C.dg.fused_gate = [sp]() { return sp->use_fast != 0; };
      C.dg.axpy_givens = [sp](int j) {
        Coarsest &K = sp->cst;
        gmres_axpy_givens(K.dg.w, K.dg.V, K.dg.stride, j, K.dg.n, K.dg.st, K.dg.ctrl, gmres_offsets(K.dg.m), K.dg.tol, K.counter);
      };
    }
#endif
  }
  C.active = true;
}

void coarsest_free(Solver &s) {
  Coarsest &C = s.cst;
  if (!C.active) return;
  dev_free(C.b); dev_free(C.x); C.b = C.x = nullptr;
  for (int i = 0; i < 4; i++) { dev_free(C.t[i]); C.t[i] = nullptr; }
  dev_free(C.dir); dev_free(C.Z); dev_free(C.gbuf); dev_free(C.d_src); dev_free(C.d_own); dev_free(C.counter); C.counter = nullptr;
  C.dg.op_dots = nullptr; C.dg.axpy_givens = nullptr; C.dg.fused_gate = nullptr;
  C.dir = C.Z = C.gbuf = nullptr; C.d_src = C.d_own = nullptr;
  C.dg.release(); C.hostk.release();
  if (C.replicated) {
    dev_free(C.rop.F); dev_free(C.rop.S); dev_free(C.rop.Sinv);
    C.rop = CoarseOp();
    C.rgeo.destroy();
  }
  C.geo = nullptr; C.op = nullptr;
  C.active = false; C.replicated = false; C.fast = false;
}

// dst[g][0..len) = gathered[src[g]][0..len)   (len complex per site)
static void permute_sites(cf *dst, const cf *gathered, const int *src, long nsites, long len) {
  launch_n(nsites * len, DLAMBDA(long i) { const long g = i / len; dst[i] = gathered[(long)src[g] * len + (i - g * len)]; });
}

void coarsest_refresh(Solver &s) {
  Coarsest &C = s.cst;
  if (!C.active) return;
  Level &L = s.lev[s.nlev - 1];
  if (C.replicated) {
    const long nn = (long)L.cop.n * L.cop.n, Vl = L.geo.V, Vg = C.rgeo.V;
    cf *stage = dev_alloc<cf>(Vg * 4 * nn);
    comm_allgather(L.cop.S, stage, sizeof(cf) * Vl * nn);
    permute_sites(C.rop.S, stage, C.d_src, Vg, nn);
    comm_allgather(L.cop.F, stage, sizeof(cf) * Vl * 4 * nn);     // the local sites' hops (the ghost sites' follow behind them)
    permute_sites(C.rop.F, stage, C.d_src, Vg, 4 * nn);
    dev_sync();
    dev_free(stage);
  }
  if (s.p.odd_even) coarse_invert_odd_self(*C.op);
}

// out_e = S_ee in_e - N_eo Soo^-1 N_oe in_e   (coarse_apply_schur_complement_PRECISION, coarse_oddeven_generic.c:1162-1189)
void mg_coarsest_schur(Solver &s, cf *out, const cf *in, const int *skip) {
  Coarsest &C = s.cst;
  const CoarseOp &op = *C.op;
  const long ne = op.n_even, no = op.V - op.n_even;
#ifndef DDA_HOST_EMU
  if (C.fast && s.use_fast) {
    schur_hop(op, 0, in, nullptr, C.dir, C.Z, skip);
    schur_mid(op, nullptr, C.dir, C.Z, C.t[1], 0.f, 1.f, -1.f, skip);
    schur_hop(op, 1, C.t[1], in, C.dir, C.Z, skip);
    schur_fin(op, nullptr, C.dir, C.Z, out, 0.f, 1.f, skip);
    return;
  }
#endif
  (void)skip;
  cf *t0 = C.t[0], *t1 = C.t[1];
  c_halo(C, in);
  coarse_apply(op, t0, in, sel_range(ne, no), HOP_ALL, 0, SELF_NONE, OUT_SET);                 // t0_o = N_oe in_e
  coarse_apply(op, t1, t0, sel_range(ne, no), HOP_NONE, 0, SELF_CINV, OUT_NEG);                // t1_o = -Soo^-1 t0_o
  c_halo(C, t1);
  coarse_apply(op, out, t1, sel_range(0, ne), HOP_ALL, 0, SELF_C, OUT_SET, nullptr, in);       // out_e = See in_e + N_eo t1_o
}

void mg_coarsest_solve(Solver &s) {
  Level &L = s.lev[s.nlev - 1];
  Coarsest &C = s.cst;
  ProfScope ps(s, &s.t_coarse_solve);
  DDA_ASSERT(C.active);
  const CoarseOp &op = *C.op;
  const int nc = op.n;
  const long ne = op.n_even, no = op.V - op.n_even;
  cf *b = L.vb, *x = L.vx;
  if (C.replicated) {
    comm_allgather(L.vb, C.gbuf, sizeof(cf) * L.geo.V * nc);
    permute_sites(C.b, C.gbuf, C.d_src, C.rgeo.V, nc);
    b = C.b; x = C.x;
  }
  auto krylov = [&](cf *xx, const cf *bb) { return C.host_driven ? C.hostk.solve(xx, bb, true) : C.dg.solve(xx, bb); };
  if (!s.p.odd_even) {
    s.coarse_iter_count += krylov(x, b);
  } else {
    cf *rhs = C.t[2];
#ifndef DDA_HOST_EMU
    if (C.fast && s.use_fast) {
      // x_o = Soo^-1 b_o ; rhs_e = b_e - N_eo x_o ; solve S x_e = rhs_e ; x_o = Soo^-1 (b_o - N_oe x_e)
      schur_mid(op, b, nullptr, nullptr, x, 1.f, 0.f, 1.f, nullptr);
      schur_hop(op, 1, x, nullptr, C.dir, C.Z, nullptr);
      schur_fin(op, b, C.dir, C.Z, rhs, 1.f, -1.f, nullptr);
      s.coarse_iter_count += krylov(x, rhs);
      schur_hop(op, 0, x, nullptr, C.dir, C.Z, nullptr);
      schur_mid(op, b, C.dir, C.Z, x, 1.f, -1.f, 1.f, nullptr);
    } else
#endif
    {
      cf *t2 = C.t[3];
      // coarse_solve_odd_even_PRECISION, coarse_oddeven_generic.c:1139-1160
      coarse_apply(op, x, b, sel_range(ne, no), HOP_NONE, 0, SELF_CINV, OUT_SET);
      c_halo(C, x);
      coarse_apply(op, rhs, x, sel_range(0, ne), HOP_ALL, 0, SELF_NONE, OUT_ETA_MINUS, b);
      s.coarse_iter_count += krylov(x, rhs);
      c_halo(C, x);
      coarse_apply(op, t2, x, sel_range(ne, no), HOP_ALL, 0, SELF_NONE, OUT_ETA_MINUS, b);
      coarse_apply(op, x, t2, sel_range(ne, no), HOP_NONE, 0, SELF_CINV, OUT_SET);
    }
  }
  if (C.replicated) {
    const int *own = C.d_own; cf *lx = L.vx; const cf *gx = C.x;
    launch_n(L.geo.V * nc, DLAMBDA(long i) { const long k = i / nc; lx[i] = gx[(long)own[k] * nc + (i - k * nc)]; });
  }
}

}  // namespace dda

// solver_fine.cu -- fine-level operator storage: gauge upload, clover construction, float copy, mass shift.
// Reference counterparts: dd_alpha_amg_set_conf (dd_alpha_amg.c:188-250), dirac_setup (dirac.c:60-170),
// schwarz_PRECISION_setup (schwarz_generic.c:1037-1074, double -> float Schwarz-ordered copy),
// schwarz_PRECISION_oddeven_setup (oddeven_generic.c:918-971), shift_update (dirac.c:669-691).
#include "solver.h"
#include "halo.h"

namespace dda {

Solver *g_solver = nullptr;

// process grid of level `depth` from the parameters (global / local lattice); identical on all levels (no idle ranks);
// rank -> grid coordinates with T slowest (the order MPI_Cart_create gives the reference, ghost.c:47-66)
void solver_process_grid(Solver &s, int depth, Geometry &g) {
  const Params &p = s.p;
  long np = 1;
  for (int m = 0; m < 4; m++) { g.P[m] = p.global_lattice[depth][m] / p.local_lattice[depth][m]; np *= g.P[m]; }
  if (np != g_comm.size) {
    fprintf(stderr, "dd_alpha_amg_b200: level %d process grid %d x %d x %d x %d needs %ld ranks, communicator has %d "
            "(call dda_comm_init before dd_alpha_amg_init)\n", depth, g.P[0], g.P[1], g.P[2], g.P[3], np, g_comm.size);
    fatal("geometry", __FILE__, __LINE__);
  }
  for (int m = 0; m < 4; m++) DDA_ASSERT(g.P[m] == p.global_lattice[0][m] / p.local_lattice[0][m]);
  int r = g_comm.rank;
  for (int m = 3; m >= 0; m--) { g.pc[m] = r % g.P[m]; r /= g.P[m]; }
  // DDA_FORCE_SPLIT=<letters of TZYX>: ghost slabs in these directions even where the process grid has extent 1
  if (const char *e = getenv("DDA_FORCE_SPLIT")) {
    const char *names = "TZYX";
    for (int m = 0; m < 4; m++) g.fg[m] = (strchr(e, names[m]) || strchr(e, names[m] + 32)) ? 1 : 0;
  }
}

void solver_alloc_fine(Solver &s) {
  DDA_ASSERT(!s.fine_alloc);
  Level &L = s.lev[0];
  const Params &p = s.p;
  L.depth = 0;
  Geometry &g = L.geo;
  for (int m = 0; m < 4; m++) {
    g.L[m] = p.local_lattice[0][m];
    if (p.num_levels > 1) {
      g.B[m] = p.block_lattice[0][m];
      g.A[m] = p.global_lattice[0][m] / p.global_lattice[1][m];
    } else { g.B[m] = 0; g.A[m] = 0; }
  }
  g.nc = 12;
  solver_process_grid(s, 0, g);
  long V = (long)g.L[0] * g.L[1] * g.L[2] * g.L[3];
  g.sh = (V % 32 == 0) ? 5 : 0;
  g.block_eo = (p.num_levels > 1) && p.odd_even;   // reference: block_solve_oddeven only with "odd even preconditioning: 1" (schwarz_generic.c:1269-1273)
  g.global_eo = false;
  g.build();
  const long VA = V + g.Vg;     // local + ghost sites (links and vectors carry ghost slabs)
  L.Dd = dev_alloc<cd>(VA * 36); L.Cd = dev_alloc<double>(V * 72);
  L.Df = dev_alloc<cf>(VA * 36); L.Cf = dev_alloc<float>(V * 72); L.Cinvf = dev_alloc<float>(V * 72);
  s.lexbuf = dev_alloc<cd>(V * 36);
  s.xb = dev_alloc<cd>(VA * 12); s.xx = dev_alloc<cd>(VA * 12);
  L.opd.D = L.Dd; L.opd.C = L.Cd; L.opd.Cinv = nullptr; L.opd.nb = g.d_nb; L.opd.blkflag = g.d_blkflag; L.opd.aggflag = g.d_aggflag; L.opd.V = V; L.opd.sh = g.sh;
  L.opf.D = L.Df; L.opf.C = L.Cf; L.opf.Cinv = L.Cinvf; L.opf.nb = g.d_nb; L.opf.blkflag = g.d_blkflag; L.opf.aggflag = g.d_aggflag; L.opf.V = V; L.opf.sh = g.sh;
  s.fine_alloc = true;
}

void solver_free_fine(Solver &s) {
  if (!s.fine_alloc) return;
  Level &L = s.lev[0];
  dev_free(L.Dd); dev_free(L.Cd); dev_free(L.Df); dev_free(L.Cf); dev_free(L.Cinvf); dev_free(L.Dblk); L.Dblk = nullptr;
  dev_free(s.lexbuf); dev_free(s.xb); dev_free(s.xx);
  L.Dd = nullptr; L.Cd = nullptr; L.Df = nullptr; L.Cf = L.Cinvf = nullptr; s.lexbuf = s.xb = s.xx = nullptr;
  L.geo.destroy();
  s.fine_alloc = false;
}

void solver_upload_conf(Solver &s, const double *gauge_lex) {
  Level &L = s.lev[0];
  long V = L.geo.V;
  h2d(s.lexbuf, gauge_lex, sizeof(cd) * 36 * V);
  spinor_from_lex<double>(L.geo, L.Dd, s.lexbuf, 36);
  vscale(L.Dd, L.Dd, 0.5, V * 36);                       // reference stores D = U/2 (dirac.c:80)
  halo_exchange<cd>(L.geo, L.Dd, 36, L.geo.sh);          // links of the ghost slabs (reference: dirac.c:405-496)
  fine_build_clover(L.geo, L.Dd, L.Cd, s.p.m0, s.p.csw, &s.plaq);
  s.m0_op = s.p.m0;
  solver_refresh_float_op(s);
  s.conf_set = true;
  // pointers handed out by dd_alpha_amg_get_gauge_pointer / get_clover_pointer stay live and current, like the
  // reference's pointers into op_double (dirac.c:171-176)
  if (!s.h_gauge.empty()) solver_sync_host_mirrors(s, false);
}

void solver_refresh_float_op(Solver &s) {
  Level &L = s.lev[0];
  long V = L.geo.V;
  cast_links(L.Dd, L.Df, (V + L.geo.Vg) * 36);
  if (L.geo.d_sapslotsite) {
    // per-block image of the in-block links for the fused SAP kernel (layout of sap::Shared2::U)
    if (!L.Dblk) L.Dblk = dev_alloc<cf>((long)L.geo.nblocks * 4 * 9 * 192);
    const cf *Df = L.Df; cf *Db = L.Dblk; const int *ss = L.geo.d_sapslotsite;
    launch_n((long)L.geo.nblocks * 4 * 9 * 192, DLAMBDA(long i) {
      const long b = i / (4 * 9 * 192); const int r = (int)(i - b * (4 * 9 * 192));
      const int mu = r / (9 * 192), k = (r / 192) % 9, sl = r % 192;
      const long site = b * 256 + ss[mu * 192 + sl];
      Db[i] = Df[(site >> 5) * (36L << 5) + ((long)(9 * mu + k) << 5) + (site & 31)];
    });
  }
  cast_reals(L.Cd, L.Cf, V * 72);
  double *tmp = dev_alloc<double>(V * 72);
  fine_invert_clover(L.geo, L.Cd, tmp);
  cast_reals(tmp, L.Cinvf, V * 72);
  dev_sync();
  dev_free(tmp);
}

// host mirrors in the reference's array formats: D[36*site + 9*mu + 3*r + c] complex, clover[42*site + k] complex
void solver_sync_host_mirrors(Solver &s, bool to_device) {
  Level &L = s.lev[0];
  long V = L.geo.V;
  s.h_gauge.resize((size_t)V * 72); s.h_clover.resize((size_t)V * 84);
  std::vector<double> c72((size_t)V * 72);
  if (!to_device) {
    spinor_to_lex<double>(L.geo, s.lexbuf, L.Dd, 36);
    d2h(s.h_gauge.data(), s.lexbuf, sizeof(cd) * 36 * V);
    reals_to_lex(L.geo, (double *)s.lexbuf, L.Cd, 72);
    d2h(c72.data(), s.lexbuf, sizeof(double) * 72 * V);
    for (long i = 0; i < V; i++) {
      for (int k = 0; k < 12; k++) { s.h_clover[84 * i + 2 * k] = c72[72 * i + k]; s.h_clover[84 * i + 2 * k + 1] = 0.0; }
      for (int k = 0; k < 60; k++) s.h_clover[84 * i + 24 + k] = c72[72 * i + 12 + k];
    }
  } else {
    h2d(s.lexbuf, s.h_gauge.data(), sizeof(cd) * 36 * V);
    spinor_from_lex<double>(L.geo, L.Dd, s.lexbuf, 36);
    halo_exchange<cd>(L.geo, L.Dd, 36, L.geo.sh);
    for (long i = 0; i < V; i++) {
      for (int k = 0; k < 12; k++) c72[72 * i + k] = s.h_clover[84 * i + 2 * k];
      for (int k = 0; k < 60; k++) c72[72 * i + 12 + k] = s.h_clover[84 * i + 24 + k];
    }
    h2d(s.lexbuf, c72.data(), sizeof(double) * 72 * V);
    reals_from_lex(L.geo, L.Cd, (const double *)s.lexbuf, 72);
    solver_refresh_float_op(s);
  }
}

void solver_shift_mass(Solver &s, double new_m0) {
  double delta = new_m0 - s.m0_op;
  if (delta == 0.0) return;
  Level &L = s.lev[0];
  fine_shift_clover(L.geo, L.Cd, delta);
  solver_refresh_float_op(s);
  if (s.setup_done) {
    for (int d = 1; d < s.nlev; d++) {
      CoarseOp &c = s.lev[d].cop;
      cf *S = c.S; int n = c.n; float df = (float)delta;
      launch_n(c.V * n, DLAMBDA(long i) { long site = i / n; int r = (int)(i - site * n); S[site * (long)n * n + (long)r * n + r].re += df; });
      if (s.lev[d].last) coarsest_refresh(s);
    }
  }
  s.m0_op = new_m0;
  if (!s.h_gauge.empty()) solver_sync_host_mirrors(s, false);   // the clover mirror follows the shifted diagonal
}


template <> void solver_apply_dw<double>(Solver &s, cd *out, const cd *in) {
  Level &L = s.lev[0];
#ifndef DDA_HOST_EMU
  if (s.use_fast && L.geo.sh == 5) {
    if (!L.geo.partitioned()) { dw_apply_fast<double>(L.opd, out, in); return; }
    // interior sites while the ghost slabs travel (second stream), then the rank-boundary sites
    halo_begin<cd>(L.geo, const_cast<cd *>(in), 12, L.geo.sh);
    dw_apply_fast<double>(L.opd, out, in, 1);
    halo_end(L.geo);
    dw_apply_fast<double>(L.opd, out, in, 2, L.geo.d_bnd, L.geo.nbnd);
    return;
  }
#endif
  halo_exchange<cd>(L.geo, const_cast<cd *>(in), 12, L.geo.sh);
  fine_apply<double>(L.opd, out, in, sel_all(L.geo.V), HOP_ALL, 0, SELF_C, OUT_SET);
}
template <> void solver_apply_dw<float>(Solver &s, cf *out, const cf *in) {
  Level &L = s.lev[0];
#ifndef DDA_HOST_EMU
  if (s.use_fast && L.geo.sh == 5) {
    if (!L.geo.partitioned()) { dw_apply_fast<float>(L.opf, out, in); return; }
    // interior sites while the ghost slabs travel (second stream), then the rank-boundary sites
    halo_begin<cf>(L.geo, const_cast<cf *>(in), 12, L.geo.sh);
    dw_apply_fast<float>(L.opf, out, in, 1);
    halo_end(L.geo);
    dw_apply_fast<float>(L.opf, out, in, 2, L.geo.d_bnd, L.geo.nbnd);
    return;
  }
#endif
  halo_exchange<cf>(L.geo, const_cast<cf *>(in), 12, L.geo.sh);
  fine_apply<float>(L.opf, out, in, sel_all(L.geo.V), HOP_ALL, 0, SELF_C, OUT_SET);
}

}  // namespace dda

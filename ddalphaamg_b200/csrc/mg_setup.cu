// mg_setup.cu -- multigrid hierarchy: level allocation, test vectors, aggregation interpolation, Galerkin coarse
// operators, bootstrap setup iterations.
//
// Reference counterparts: method_setup / next_level_setup (init.c:32-116,134-283), interpolation_PRECISION_define
// (setup_generic.c:191-275), coarse_grid_correction_PRECISION_setup (:29-108), coarse_operator_PRECISION_setup
// (coarse_operator_generic.c:53-205), re_setup_PRECISION (setup_generic.c:278-321), inv_iter_inv_fcycle_PRECISION
// (:441-503), gram_schmidt_PRECISION (linalg_generic.c:356-397), coarse_oddeven_setup (coarse_oddeven_generic.c:200-406).
#include "solver.h"
#include "halo.h"

namespace dda {

// deterministic counter-based uniform numbers in [-0.5, 0.5) (the reference uses libc rand(), data_generic.c:42-56)
static HD float u01(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; x ^= x >> 31;
  return (float)((x >> 40) * (1.0 / 16777216.0)) - 0.5f;
}
static void vrandom(cf *v, long n, unsigned long long seed) {
  launch_n(n, DLAMBDA(long i) { v[i] = cf(u01(seed + 2ULL * i), u01(seed + 2ULL * i + 1ULL)); });
}

#ifndef DDA_HOST_EMU
bool galerkin_fine_fast(const FineOp<float> &op, const Transfer &t, cf *S, cf *F);
static int g_galerkin_fast = []() { const char *e = getenv("DDA_GALERKIN_FAST"); return e ? atoi(e) : 1; }();
#endif

// setup phase timers (DDA_SETUP_PROFILE=1: device-synchronising, printed at the end of mg_setup)
static int g_setup_profile = []() { const char *e = getenv("DDA_SETUP_PROFILE"); return e ? atoi(e) : 0; }();
static double g_t_setup[6] = {0, 0, 0, 0, 0, 0};   // 0 test-vector smoothing, 1 interpolation (Gram-Schmidt), 2 Galerkin, 3 bootstrap cycles, 4 global GS, 5 other
struct SetupTimer {
  int k; double t0;
  explicit SetupTimer(int k_) : k(k_), t0(0) { if (g_setup_profile) { dev_sync(); t0 = now_s(); } }
  ~SetupTimer() { if (g_setup_profile) { dev_sync(); g_t_setup[k] += now_s() - t0; } }
};

static void build_transfer(Level &L, Level &N) {
  Transfer &t = L.tr;
  t.lay = L.geo.lay(); t.V = L.geo.V; t.nc = L.geo.nc; t.nv = L.nv; t.nagg = L.geo.nagg; t.as = L.geo.as;
  t.agg2coarse = L.geo.d_agg2coarse;
  for (int k = 0; k < L.nv; k++) t.P[k] = L.P[k];
  (void)N;
}

void mg_alloc(Solver &s) {
  const Params &p = s.p;
  s.nlev = p.num_levels;
  double t_mark = now_s();
  auto mark = [&](const char *what) {
    if (g_setup_profile) { const double t = now_s(); fprintf(stderr, "dd_alpha_amg_b200 mg_alloc: %s %.3f s\n", what, t - t_mark); t_mark = t; }
  };
  for (int d = 0; d < s.nlev; d++) {
    Level &L = s.lev[d];
    L.depth = d; L.last = (d == s.nlev - 1);
    L.nv = L.last ? 0 : p.num_eig_vect[d];
    DDA_ASSERT(L.nv <= MAX_NV);
    if (d > 0) {
      Geometry &g = L.geo;
      for (int m = 0; m < 4; m++) {
        g.L[m] = p.local_lattice[d][m];
        if (!L.last) { g.B[m] = p.block_lattice[d][m]; g.A[m] = p.global_lattice[d][m] / p.global_lattice[d + 1][m]; }
        else { g.B[m] = 0; g.A[m] = 0; }
      }
      g.nc = 2 * s.lev[d - 1].nv; g.sh = 0; g.block_eo = false; g.global_eo = L.last && p.odd_even;
      solver_process_grid(s, d, g);
      g.build();
      CoarseOp &c = L.cop;
      c.n = g.nc; c.V = g.V; c.n_even = L.last && p.odd_even ? g.n_even : g.V;
      long nn = (long)c.n * c.n;
      c.F = dev_alloc<cf>((g.V + g.Vg) * 4 * nn); c.S = dev_alloc<cf>(g.V * nn);   // hops of the ghost sites too
      c.Sinv = (L.last && p.odd_even) ? dev_alloc<cf>((g.V - c.n_even) * nn) : nullptr;
      c.nb = g.d_nb; c.blkflag = g.d_blkflag; c.aggflag = g.d_aggflag;
      L.copZ = dev_alloc<cf>(g.V * 4 * c.n);
    }
  }
  mark("coarse geometries and operators");
  for (int d = 0; d < s.nlev; d++) {
    Level &L = s.lev[d];
    const long n = L.geo.valloc();
    if (!L.last) {
      Level &N = s.lev[d + 1];
      std::vector<int> a2c(L.geo.nagg);
      for (int a = 0; a < L.geo.nagg; a++) a2c[a] = N.geo.lex2nat[a];
      DDA_ASSERT(N.geo.V == L.geo.nagg);
      L.geo.d_agg2coarse = dev_upload(a2c);
      L.tv.resize(L.nv); L.P.resize(L.nv);
      for (int k = 0; k < L.nv; k++) { L.tv[k] = dev_alloc<cf>(n); L.P[k] = dev_alloc<cf>(n); }
      build_transfer(L, N);
      L.tr_scratch = dev_alloc<double>(tr_scratch_doubles(L.tr));
      L.blockred = dev_alloc<double>(3L * L.geo.nblocks + 8);
    }
    L.vb = dev_alloc<cf>(n); L.vx = dev_alloc<cf>(n);
    int nw = L.last ? 4 : NWORK;
    for (int i = 0; i < nw; i++) L.w[i] = dev_alloc<cf>(n);
    // Krylov wrappers
    if (d > 0 && !L.last) {
      L.kc.alloc(L.geo.vlen(), p.kcycle_restart, p.kcycle_max_restart, p.kcycle_tol, true, n);
      Solver *sp = &s; int dd = d;
      { const char *e = getenv("DDA_SINGLE_REDUCTION"); L.kc.single_reduction = e && atoi(e) != 0; }
      L.kc.op = [sp, dd](cf *out, const cf *in) { mg_apply_op(*sp, dd, out, in); };
      L.kc.prec = [sp, dd](cf *out, const cf *in) { mg_vcycle(*sp, dd, out, in, true); };
    }
  }
  mark("level vectors");
  coarsest_alloc(s);     // coarsest-level solver (device-resident GMRES, gathered lattice when the level is partitioned)
  mark("coarsest solver");
}

void mg_free(Solver &s) {
  coarsest_free(s);
  for (int d = 0; d < s.nlev; d++) {
    Level &L = s.lev[d];
    for (auto v : L.tv) dev_free(v);
    for (auto v : L.P) dev_free(v);
    L.tv.clear(); L.P.clear();
    dev_free(L.tr_scratch); L.tr_scratch = nullptr;
    dev_free(L.blockred); L.blockred = nullptr;
    dev_free(L.vb); dev_free(L.vx); L.vb = L.vx = nullptr;
    for (int i = 0; i < NWORK; i++) { dev_free(L.w[i]); L.w[i] = nullptr; }
    L.kc.release();
    if (d > 0) {
      dev_free(L.cop.F); dev_free(L.cop.S); dev_free(L.cop.Sinv); dev_free(L.copZ); L.copZ = nullptr;
      L.cop.F = L.cop.S = L.cop.Sinv = nullptr;
      L.geo.destroy();
    } else {
      dev_free(L.geo.d_agg2coarse); L.geo.d_agg2coarse = nullptr;
    }
  }
  s.setup_done = false;
}

// Galerkin operator of level depth+1:  S = P^H (C + N_inside-aggregate) P,  F_mu = P^H N_{+mu, across aggregates} P.
// One operator application + one restriction per coarse column and coupling (the reference assembles the same
// products column by column, coarse_operator_generic.c:53-205).
void mg_rebuild_coarse(Solver &s, int depth) {
  SetupTimer st_(2);
  Level &L = s.lev[depth], &N = s.lev[depth + 1];
  CoarseOp &c = N.cop;
  const int n = c.n, nv = L.nv; const long nn = (long)n * n;
  cf *w0 = L.w[6], *w1 = L.w[7];
  SiteSel all = sel_all(L.geo.V);
  bool done = false;
#ifndef DDA_HOST_EMU
  if (depth == 0 && s.use_fast && g_galerkin_fast) {
    // fused construction (galerkin_kernel.cu): all columns in one launch, no intermediate vectors
    for (int k = 0; k < nv; k++) lv_halo(L, L.P[k]);
    // (4^4 aggregates: the kernel maps the four +mu faces of 64 sites onto its 256 threads)
    if (L.geo.A[0] == 4 && L.geo.A[1] == 4 && L.geo.A[2] == 4 && L.geo.A[3] == 4) done = galerkin_fine_fast(L.opf, L.tr, c.S, c.F);
  }
#endif
  for (int j = 0; j < n && !done; j++) {
    int ch = j / nv, k = j - ch * nv;
    tr_chirality_part(L.tr, w0, L.P[k], ch);
    lv_halo(L, w0);
    lv_apply(L, w1, w0, all, HOP_INAGG, 0, SELF_C, OUT_SET);
    tr_restrict(L.tr, c.S, nn, (long)j * n, w1, L.tr_scratch);
    for (int mu = 0; mu < 4; mu++) {
      lv_apply(L, w1, w0, all, HOP_CROSSAGG, mu, SELF_NONE, OUT_SET);
      tr_restrict(L.tr, c.F + mu * nn, 4 * nn, (long)j * n, w1, L.tr_scratch);
    }
  }
  halo_exchange<cf>(N.geo, c.F, 4 * (int)nn, 0);     // forward hops of the -mu ghost sites (used by the backward hop)
  if (N.last) coarsest_refresh(s);
}

static void define_interpolation(Solver &s, int depth) {
  SetupTimer st_(1);
  Level &L = s.lev[depth];
  const long n = L.geo.vlen();
  for (int k = 0; k < L.nv; k++) vcopy(L.P[k], L.tv[k], n);
  tr_gram_schmidt_aggregates(L.tr, L.P.data(), L.tr_scratch);
  if (depth > 0) tr_gram_schmidt_aggregates(L.tr, L.P.data(), L.tr_scratch);
}

// re_setup_PRECISION (setup_generic.c:278-321): P from the current test vectors, then the coarse operator, recursively
static void re_setup(Solver &s, int depth) {
  if (s.lev[depth].last) return;
  define_interpolation(s, depth);
  mg_rebuild_coarse(s, depth);
  re_setup(s, depth + 1);
}

static void normalise(cf *v, long n) {
  double nr = std::sqrt(vnorm2(v, n));
  if (nr > 0) vscale(v, v, 1.0 / nr, n);
}

// initial test vectors of a level: random vectors smoothed by 1+2+3 SAP iterations (setup_generic.c:215-231)
// (on every level: the restricted finer test vectors are overwritten by random ones in the reference, too)
static void initial_test_vectors(Solver &s, int depth) {
  SetupTimer st_(0);
  Level &L = s.lev[depth];
  const long n = L.geo.vlen();
  cf *buf = L.w[8];
  for (int k = 0; k < L.nv; k++) {
    vrandom(L.tv[k], n, s.seed + ((unsigned long long)(depth * 1000 + k) << 40));
    for (int it = 1; it <= 3; it++) {
      mg_smoother(s, depth, buf, L.tv[k], it, true);
      vcopy(L.tv[k], buf, n);
    }
  }
  for (int k = 0; k < L.nv; k++) normalise(L.tv[k], n);
}

// global Gram-Schmidt on the test vectors (gram_schmidt_PRECISION, linalg_generic.c:356-397)
static void gram_schmidt_global(Level &L) {
  SetupTimer st_(4);
  const long n = L.geo.vlen();
  std::vector<cd> co(L.nv);
  for (int k = 0; k < L.nv; k++) {
    if (k > 0) {
      vmulti_dot(co.data(), L.tv.data(), k, L.tv[k], n);
      vmulti_axpy(L.tv[k], L.tv.data(), co.data(), k, -1, n);
    }
    normalise(L.tv[k], n);
  }
}

// inv_iter_inv_fcycle_PRECISION (setup_generic.c:441-503)
static void bootstrap(Solver &s, int depth, int setup_iter) {
  Level &L = s.lev[depth];
  if (L.last) return;
  const long n = L.geo.vlen();
  Level &N = s.lev[depth + 1];
  for (int j = 0; j < setup_iter; j++) {
    gram_schmidt_global(L);
    for (int i = 0; i < L.nv; i++) {
      // one cycle with the test vector as right-hand side; every level keeps its iterate as new test vector
      { SetupTimer st_(3); mg_vcycle(s, depth, L.vx, L.tv[i], true); }
      for (int d = s.nlev - 2; d > depth; d--) {
        // test_vector_PRECISION_update (setup_generic.c:428-438): deeper levels take the solution of their last solve
        Level &D = s.lev[d];
        if (i < D.nv) { vcopy(D.tv[i], D.vx, D.geo.vlen()); normalise(D.tv[i], D.geo.vlen()); }
      }
      vcopy(L.tv[i], L.vx, n); normalise(L.tv[i], n);
    }
    re_setup(s, depth);
    if (depth == 0 && !N.last) {
      int it = (int)std::lround(((double)(j + 1) * s.p.setup_iter[1]) / (double)setup_iter);
      bootstrap(s, depth + 1, std::max(1, it));
    }
  }
  if (depth > 0 && !N.last) {
    int it = (int)std::lround((double)(s.p.setup_iter[depth + 1] * setup_iter) / (double)s.p.setup_iter[depth]);
    bootstrap(s, depth + 1, std::max(1, it));
  }
}

static void set_kcycle_tol(Solver &s, double tol) {
  for (int d = 1; d < s.nlev - 1; d++) s.lev[d].kc.tol = tol;
}

// initial setup (method_setup -> next_level_setup, init.c:134-283) followed by `setup_iters` bootstrap iterations
// (method_update -> iterative_PRECISION_setup, init.c:326-373)
void mg_setup(Solver &s, int setup_iters) {
  DDA_ASSERT(s.conf_set);
  if (s.setup_done) mg_free(s);
  s.coarse_iter_count = 0;
  const Params &p = s.p;
  Level &L0 = s.lev[0];
  // outer solver
  {
    Solver *sp = &s;
    const bool mg = p.method > 0 && p.num_levels > 1;
    if (p.mixed_precision == 2) {
      s.outer.release();
      s.outer_mp.alloc(L0.geo.vlen(), p.restart, p.max_restart, p.tol, mg, L0.geo.valloc());
      s.outer_mp.op_d = [sp](cd *out, const cd *in) { solver_apply_dw<double>(*sp, out, in); };
      s.outer_mp.op_f = [sp](cf *out, const cf *in) { solver_apply_dw<float>(*sp, out, in); };
      if (mg) s.outer_mp.prec_f = [sp](cf *out, const cf *in) { mg_vcycle(*sp, 0, out, in, true); };
      else s.outer_mp.prec_f = nullptr;
    } else {
      s.outer_mp.release();
      s.outer.alloc(L0.geo.vlen(), p.restart, p.max_restart, p.tol, mg, L0.geo.valloc());
      s.outer.op = [sp](cd *out, const cd *in) { solver_apply_dw<double>(*sp, out, in); };
      if (mg) s.outer.prec = [sp](cd *out, const cd *in) { mg_preconditioner(*sp, out, in); };
      else s.outer.prec = nullptr;
    }
  }
  if (p.method <= 0 || p.num_levels < 2) { s.nlev = 1; s.setup_done = true; return; }
  const double t_begin = now_s();
  { SetupTimer st_(5); mg_alloc(s); }
  double m_solve = s.m0_op;
  if (p.setup_m0 != s.m0_op) solver_shift_mass(s, p.setup_m0);
  for (int d = 0; d + 1 < s.nlev; d++) {
    initial_test_vectors(s, d);
    define_interpolation(s, d);
    mg_rebuild_coarse(s, d);
  }
  s.setup_done = true;
  if (setup_iters > 0 && p.interpolation == 4) {
    // iterative_PRECISION_setup case 4 (setup_generic.c:111-118): the setup iterations are replaced by test vectors from files
    tv_read(s, p.tv_file.c_str());
    re_setup(s, 0);
  } else if (setup_iters > 0) {
    set_kcycle_tol(s, p.coarse_tol);
    bootstrap(s, 0, setup_iters);
    set_kcycle_tol(s, p.kcycle_tol);
  }
  if (m_solve != s.m0_op) solver_shift_mass(s, m_solve);
  dev_sync();
  if (g_setup_profile) {
    fprintf(stderr, "dd_alpha_amg_b200 setup phases [s]: test-vector smoothing %.3f, aggregate Gram-Schmidt %.3f, Galerkin %.3f, bootstrap cycles %.3f (nested levels included), global Gram-Schmidt %.3f, level allocation %.3f, total %.3f\n",
            g_t_setup[0], g_t_setup[1], g_t_setup[2], g_t_setup[3], g_t_setup[4], g_t_setup[5], now_s() - t_begin);
    for (double &t : g_t_setup) t = 0;
  }
}

void mg_resetup_from_test_vectors(Solver &s) {
  DDA_ASSERT(s.setup_done && s.nlev > 1);
  re_setup(s, 0);
  dev_sync();
}

void mg_setup_update(Solver &s, int setup_iters) {
  if (!s.setup_done) { mg_setup(s, setup_iters); return; }
  if (s.nlev < 2) return;
  s.coarse_iter_count = 0;
  double m_solve = s.m0_op;
  if (s.p.setup_m0 != s.m0_op) solver_shift_mass(s, s.p.setup_m0);
  re_setup(s, 0);   // operators follow the current gauge field
  if (setup_iters > 0) {
    set_kcycle_tol(s, s.p.coarse_tol);
    bootstrap(s, 0, setup_iters);
    set_kcycle_tol(s, s.p.kcycle_tol);
  }
  if (m_solve != s.m0_op) solver_shift_mass(s, m_solve);
  dev_sync();
}

}  // namespace dda

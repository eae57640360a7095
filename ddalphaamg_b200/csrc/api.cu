// api.cu -- C ABI of libdd_alpha_amg: the reference's library interface (include/dd_alpha_amg.h, implemented in the
// reference by src/dd_alpha_amg.c:95-404) plus the operator-level entry points of include/dd_alpha_amg_b200.h that
// tests and benchmarks use to reach individual hot-path operators.
#include "solver.h"
#include "halo.h"
#include "../../include/dd_alpha_amg.h"
#include "../../include/dd_alpha_amg_b200.h"
#include <chrono>

using namespace dda;

namespace {
struct ApiState {
  Solver s;
  int (*conf_index_fct)(int, int, int, int, int) = nullptr;
  int (*vector_index_fct)(int, int, int, int) = nullptr;
  int (*global_time)(int) = nullptr;
  int bc = 1;
  bool initialised = false;
  std::vector<double> stage;     // host staging (lexicographic)
  // struct-route setup policy (dd_alpha_amg_parameters.h:36-38, dd_alpha_amg_setup_status.h)
  dd_alpha_amg_parameters amg = {};
  bool have_amg = false;
  int updates_since_setup = 0, updates_since_setup_update = 0;
};
ApiState *A = nullptr;

void need_init(const char *fn) {
  if (!A || !A->initialised) { fprintf(stderr, "%s called before dd_alpha_amg_init\n", fn); fatal("API misuse", __FILE__, __LINE__); }
}

#ifndef DDA_HOST_EMU
void select_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fprintf(stderr, "dd_alpha_amg_b200: no CUDA device available (%s); this library has no CPU fallback\n", cudaGetErrorString(e));
    fatal("no GPU", __FILE__, __LINE__);
  }
  if (g_comm.active() && g_stream) return;   // dda_comm_init already bound this process to its GPU
  const char *lr = getenv("DDA_DEVICE");
  if (!lr) lr = getenv("LOCAL_RANK");
  int dev = lr ? atoi(lr) % n : -1;
  if (dev >= 0) CUDA_CHECK(cudaSetDevice(dev));
  if (!g_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
}
#else
void select_device() {}
#endif

void init_common(dd_alpha_amg_par &p, bool from_struct) {
  if (A && A->initialised) dd_alpha_amg_free();
  select_device();
  A = new ApiState();
  Solver &s = A->s;
  g_solver = &s;
  if (!from_struct) {
    params_from_ini(s.p, p.param_file_path);
  } else {
    // struct route (init.c:817-901): lattice arrays arrive in X,Y,Z,T order and are reversed
    const dd_alpha_amg_parameters &ap = p.amg_params;
    Params &q = s.p;
    q.num_levels = ap.number_of_levels;
    int ls = q.num_levels < 2 ? 2 : q.num_levels;
    for (int i = 0; i < ls && i < MAX_LEVELS; i++) {
      for (int m = 0; m < 4; m++) {
        q.global_lattice[i][m] = ap.global_lattice[i][3 - m];
        q.local_lattice[i][m] = ap.local_lattice[i][3 - m];
        q.block_lattice[i][m] = ap.block_lattice[i][3 - m];
      }
      q.num_eig_vect[i] = ap.mg_basis_vectors[i];
      q.post_smooth_iter[i] = ap.post_smooth_iterations[i];
      q.block_iter[i] = ap.post_smooth_block_iterations[i];
      q.setup_iter[i] = ap.setup_iterations[i];
    }
    q.mixed_precision = 1; q.interpolation = 2; q.randomize = 0;
    q.coarse_iter = ap.coarse_grid_iterations; q.coarse_restart = ap.coarse_grid_maximum_number_of_restarts;
    q.coarse_tol = ap.coarse_grid_tolerance; q.odd_even = 1;
    q.m0 = ap.solver_mass; q.setup_m0 = ap.setup_mass; q.csw = ap.c_sw;
    q.method = 2; q.max_restart = 100; q.tol = 1e-10; q.print = 1;
    q.restart = 10;   // the reference allocates no outer solver on this route (init.c:893); we keep a small one
  }
  s.p.csw = p.csw; s.p.m0 = p.m0; s.p.setup_m0 = p.setup_m0;
  if (p.bc == 2) s.p.anti_pbc = 1;
  params_finalize(s.p);
  A->conf_index_fct = p.conf_index_fct; A->vector_index_fct = p.vector_index_fct; A->global_time = p.global_time;
  A->bc = p.bc;
  if (from_struct) {
    // init.c:899-900: the counters start at their thresholds, so the first "if necessary" check runs a setup
    A->amg = p.amg_params; A->have_amg = true;
    A->updates_since_setup = p.amg_params.discard_setup_after; A->updates_since_setup_update = p.amg_params.update_setup_after;
  }
  s.seed = 20261018ULL + 7919ULL * (unsigned long long)g_comm.rank;
  solver_alloc_fine(s);
  A->initialised = true;
}

inline long lexsite(const int *L, int t, int z, int y, int x) { return x + (long)L[3] * (y + (long)L[2] * (z + (long)L[1] * t)); }

// host lexicographic float/complex64 <-> device native vector of a level
void level_upload(Level &L, cf *dst, const float *src_lex) {
  long n = L.geo.vlen();
  std::vector<double> tmp((size_t)2 * n);
  for (long i = 0; i < 2 * n; i++) tmp[i] = src_lex[i];
  cd *d = dev_alloc<cd>(n);
  h2d(d, tmp.data(), sizeof(cd) * n);
  spinor_from_lex<float>(L.geo, dst, d, L.geo.nc);
  dev_sync(); dev_free(d);
}
void level_download(Level &L, float *dst_lex, const cf *src) {
  long n = L.geo.vlen();
  std::vector<double> tmp((size_t)2 * n);
  cd *d = dev_alloc<cd>(n);
  spinor_to_lex<float>(L.geo, d, src, L.geo.nc);
  d2h(tmp.data(), d, sizeof(cd) * n);
  dev_free(d);
  for (long i = 0; i < 2 * n; i++) dst_lex[i] = (float)tmp[i];
}
}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ reference interface
void dd_alpha_amg_init(dd_alpha_amg_par p) { init_common(p, false); }
void dd_alpha_amg_init_external_threading(dd_alpha_amg_par p, int n_core, int n_thread) { (void)n_core; (void)n_thread; init_common(p, true); }

double dd_alpha_amg_set_conf(double *gauge_field) {
  need_init("dd_alpha_amg_set_conf");
  Solver &s = A->s;
  const int *L = s.lev[0].geo.L;
  long V = s.lev[0].geo.V;
  std::vector<double> &h = A->stage;
  h.resize((size_t)V * 72);
  std::vector<double> hc;
  const bool dirichlet = (A->bc == 0);
  if (dirichlet) hc.resize((size_t)V * 72);
  long j = 0; int ifail = 0;
  const int Tg = s.p.global_lattice[0][0];
  for (int t = 0; t < L[0]; t++) {
    // without a callback the global time follows from the rank's process coordinate (the reference requires the callback)
    int tg = A->global_time ? A->global_time(t) : s.lev[0].geo.pc[0] * L[0] + t;
    for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++) for (int x = 0; x < L[3]; x++)
      for (int mu = 0; mu < 4; mu++) {
        long i = A->conf_index_fct ? (long)A->conf_index_fct(t, z, y, x, mu) : 18 * (4 * lexsite(L, t, z, y, x) + mu);
        // Dirichlet in time (bc 0, dd_alpha_amg.c:206-235): time links on slices 0, T-2, T-1 enter the clover term only
        bool cut = dirichlet && mu == 0 && (tg == 0 || tg >= Tg - 2);
        for (int k = 0; k < 18; k++, j++) {
          double v = gauge_field[i + k];
          if (dirichlet) { hc[j] = v; if (cut && tg == Tg - 1 && v != 0.0) ifail++; }
          h[j] = cut ? 0.0 : v;
        }
      }
  }
  if (ifail) { fprintf(stderr, "Error in \"dd_alpha_amg_set_conf\": Gauge field does not fit expected boundary conditions.\n"); fatal("set_conf", __FILE__, __LINE__); }
  if (dirichlet) {
    solver_upload_conf(s, hc.data());            // clover from the full links
    Level &L0 = s.lev[0];
    h2d(s.lexbuf, h.data(), sizeof(cd) * 36 * V);
    spinor_from_lex<double>(L0.geo, L0.Dd, s.lexbuf, 36);
    vscale(L0.Dd, L0.Dd, 0.5, V * 36);
    halo_exchange<cd>(L0.geo, L0.Dd, 36, L0.geo.sh);
    solver_refresh_float_op(s);
    if (!s.h_gauge.empty()) solver_sync_host_mirrors(s, false);
  } else {
    solver_upload_conf(s, h.data());
  }
  return s.plaq;
}

double *dd_alpha_amg_get_gauge_pointer(void) {
  need_init("dd_alpha_amg_get_gauge_pointer");
  if (A->s.h_gauge.empty()) solver_sync_host_mirrors(A->s, false);
  return A->s.h_gauge.data();
}
double *dd_alpha_amg_get_clover_pointer(void) {
  need_init("dd_alpha_amg_get_clover_pointer");
  if (A->s.h_clover.empty()) solver_sync_host_mirrors(A->s, false);
  return A->s.h_clover.data();
}
void dd_alpha_amg_fields_updated(void) {
  need_init("dd_alpha_amg_fields_updated");
  A->updates_since_setup++; A->updates_since_setup_update++;      // dd_alpha_amg.c:182-185
  if (!A->s.h_gauge.empty()) { solver_sync_host_mirrors(A->s, true); A->s.conf_set = true; }
}

void dd_alpha_amg_update_parameters(const struct dd_alpha_amg_parameters *ap) {
  need_init("dd_alpha_amg_update_parameters");
  Params &q = A->s.p;
  int ls = q.num_levels < 2 ? 2 : q.num_levels;
  for (int i = 0; i < ls && i < MAX_LEVELS; i++) {
    q.post_smooth_iter[i] = ap->post_smooth_iterations[i];
    q.block_iter[i] = ap->post_smooth_block_iterations[i];
    q.setup_iter[i] = ap->setup_iterations[i];
  }
  if (ap->solver_mass != A->s.m0_op && A->s.conf_set) solver_shift_mass(A->s, ap->solver_mass);
  q.m0 = ap->solver_mass;
}

void dd_alpha_amg_setup(int iterations, int *status) {
  need_init("dd_alpha_amg_setup");
  mg_setup(A->s, iterations);
  A->updates_since_setup = 0; A->updates_since_setup_update = 0;  // init.c:277-278
  if (status) { status[0] = 1; status[1] = (int)A->s.coarse_iter_count; }
}
void dd_alpha_amg_setup_external_threading(int iterations, int *status, int core, int thread, void *bd, void (*bf)(void *, int)) {
  (void)thread; (void)bd; (void)bf;
  if (core == 0) dd_alpha_amg_setup(iterations, status);
}
void dd_alpha_amg_setup_update(int iterations, int *status) {
  need_init("dd_alpha_amg_setup_update");
  mg_setup_update(A->s, iterations);
  A->updates_since_setup_update = 0;                               // init.c:366
  if (status) { status[0] = 1; status[1] = (int)A->s.coarse_iter_count; }
}
void dd_alpha_amg_setup_update_external_threading(int iterations, int *status, int core, int thread, void *bd, void (*bf)(void *, int)) {
  (void)thread; (void)bd; (void)bf;
  if (core == 0) dd_alpha_amg_setup_update(iterations, status);
}

// offsets of the caller's vector layout, evaluated once (the callback is a per-site function call in the reference,
// dd_alpha_amg.c:345-352, on every solve) and reused by all later solves
static const std::vector<long> &vector_offsets() {
  static std::vector<long> off;
  static void *for_fct = nullptr; static long for_V = -1;
  Level &L0 = A->s.lev[0];
  const int *L = L0.geo.L; const long V = L0.geo.V;
  if (for_fct != (void *)A->vector_index_fct || for_V != V) {
    off.resize(V);
    long j = 0;
    for (int t = 0; t < L[0]; t++) for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++) for (int x = 0; x < L[3]; x++)
      off[j++] = A->vector_index_fct(t, z, y, x);
    for_fct = (void *)A->vector_index_fct; for_V = V;
  }
  return off;
}
static void gather_source(const double *in, cd *dev_native) {
  Solver &s = A->s; Level &L0 = s.lev[0];
  long V = L0.geo.V;
  if (A->vector_index_fct) {
    std::vector<double> &h = A->stage;
    h.resize((size_t)V * 24);
    const std::vector<long> &off = vector_offsets();
#pragma omp parallel for schedule(static)
    for (long j = 0; j < V; j++) memcpy(&h[24 * j], in + off[j], 24 * sizeof(double));
    h2d(s.lexbuf, h.data(), sizeof(cd) * 12 * V);
  } else h2d(s.lexbuf, in, sizeof(cd) * 12 * V);
  spinor_from_lex<double>(L0.geo, dev_native, s.lexbuf, 12);
}
static void scatter_solution(double *out, const cd *dev_native) {
  Solver &s = A->s; Level &L0 = s.lev[0];
  long V = L0.geo.V;
  spinor_to_lex<double>(L0.geo, s.lexbuf, dev_native, 12);
  if (A->vector_index_fct) {
    std::vector<double> &h = A->stage;
    h.resize((size_t)V * 24);
    d2h(h.data(), s.lexbuf, sizeof(cd) * 12 * V);
    const std::vector<long> &off = vector_offsets();
#pragma omp parallel for schedule(static)
    for (long j = 0; j < V; j++) memcpy(out + off[j], &h[24 * j], 24 * sizeof(double));
  } else d2h(out, s.lexbuf, sizeof(cd) * 12 * V);
}

// clover scaling by site parity as the reference's wilson_solve does (dd_alpha_amg.c:354-373); the hierarchy follows.
// The unscaled clover term is kept in a device backup and restored afterwards (dd_alpha_amg.c:368-373).
static double *clover_backup = nullptr;
static void apply_clover_scale(double se, double so) {
  Solver &s = A->s; Level &L0 = s.lev[0];
  clover_backup = dev_alloc<double>(L0.geo.V * 72);
  d2d(clover_backup, L0.Cd, sizeof(double) * 72 * L0.geo.V);
  fine_scale_clover(L0.geo, L0.Cd, se, so);
  solver_refresh_float_op(s);
  if (s.setup_done) for (int d = 0; d + 1 < s.nlev; d++) mg_rebuild_coarse(s, d);
}
static void restore_clover() {
  Solver &s = A->s; Level &L0 = s.lev[0];
  d2d(L0.Cd, clover_backup, sizeof(double) * 72 * L0.geo.V);
  solver_refresh_float_op(s);
  if (s.setup_done) for (int d = 0; d + 1 < s.nlev; d++) mg_rebuild_coarse(s, d);
  dev_sync(); dev_free(clover_backup); clover_backup = nullptr;
}

double dd_alpha_amg_wilson_solve(double *vector_out, double *vector_in, double tol, double scale_even, double scale_odd, int *status) {
  need_init("dd_alpha_amg_wilson_solve");
  Solver &s = A->s;
  if (!s.setup_done) { fprintf(stderr, "dd_alpha_amg_wilson_solve: dd_alpha_amg_setup has not been called\n"); fatal("API misuse", __FILE__, __LINE__); }
  gather_source(vector_in, s.xb);
  const bool scaled = (scale_even != 1.0 || scale_odd != 1.0);
  if (scaled) apply_clover_scale(scale_even, scale_odd);
  double res = mg_solve(s, s.xx, s.xb, tol, status);
  if (scaled) restore_clover();
  scatter_solution(vector_out, s.xx);
  if (s.p.print > 0) printf("dd_alpha_amg_b200: solve: %ld iterations, %ld coarsest iterations, relative residual %.6e\n", s.iter_count, s.coarse_iter_count, res);
  return res;
}

void dd_alpha_amg_preconditioner(double *vector_out, double *vector_in, double scale_even, double scale_odd, int *status) {
  need_init("dd_alpha_amg_preconditioner");
  Solver &s = A->s;
  if (!s.setup_done) { fprintf(stderr, "dd_alpha_amg_preconditioner: dd_alpha_amg_setup has not been called\n"); fatal("API misuse", __FILE__, __LINE__); }
  gather_source(vector_in, s.xb);
  const bool scaled = (scale_even != 1.0 || scale_odd != 1.0);
  if (scaled) apply_clover_scale(scale_even, scale_odd);
  s.coarse_iter_count = 0;
  mg_preconditioner(s, s.xx, s.xb);
  if (scaled) restore_clover();
  scatter_solution(vector_out, s.xx);
  if (status) { status[0] = 1; status[1] = (int)s.coarse_iter_count; }
}
void dd_alpha_amg_preconditioner_external_threading(double *vector_out, double *vector_in, int *status, int core, int thread, void *bd, void (*bf)(void *, int)) {
  (void)thread; (void)bd; (void)bf;
  if (core == 0) dd_alpha_amg_preconditioner(vector_out, vector_in, 1.0, 1.0, status);
}

void dd_alpha_amg_free(void) {
  if (!A) return;
  Solver &s = A->s;
  if (s.setup_done && s.nlev > 1) mg_free(s);
  s.outer.release();
  s.outer_mp.release();
  solver_free_fine(s);
  halo_finalize();
  delete A; A = nullptr; g_solver = nullptr;
}

void DDalphaAMG_initialize(dd_alpha_amg_par p) { dd_alpha_amg_init(p); }
void DDalphaAMG_update_parameters(const struct dd_alpha_amg_parameters *ap) { dd_alpha_amg_update_parameters(ap); }
void DDalphaAMG_setup(int iterations, int *status) { dd_alpha_amg_setup(iterations, status); }
double DDalphaAMG_solve(double *out, double *in, double tol, int *status) { return dd_alpha_amg_wilson_solve(out, in, tol, 1.0, 1.0, status); }
void DDalphaAMG_finalize(void) { dd_alpha_amg_free(); }

// ------------------------------------------------------------------------------------ setup policy / test-vector files
int dda_setup_if_necessary(void) {
  need_init("dda_setup_if_necessary");
  Solver &s = A->s;
  int did = 0, st[2];
  if (A->have_amg) {
    if (!s.setup_done || A->updates_since_setup >= A->amg.discard_setup_after) { dd_alpha_amg_setup(s.p.setup_iter[0], st); did = 2; }
    else if (A->updates_since_setup_update >= A->amg.update_setup_after) { dd_alpha_amg_setup_update(A->amg.update_setup_iterations[0], st); did = 1; }
  }
  if (s.conf_set && s.p.m0 != s.m0_op) solver_shift_mass(s, s.p.m0);
  return did;
}
void dda_write_test_vectors(const char *basename) { need_init("dda_write_test_vectors"); tv_write(A->s, basename); }
void dda_read_test_vectors(const char *basename) {
  need_init("dda_read_test_vectors");
  tv_read(A->s, basename);
  mg_resetup_from_test_vectors(A->s);
}

// ------------------------------------------------------------------------------------ operator-level entry points
int dda_is_emulation(void) {
#ifdef DDA_HOST_EMU
  return 1;
#else
  return 0;
#endif
}

int dda_info(int what, int depth) {
  need_init("dda_info");
  Solver &s = A->s;
  int nl = s.setup_done ? s.nlev : 1;
  if (what == DDA_INFO_NUM_LEVELS) return s.setup_done ? s.nlev : s.p.num_levels;
  if (what == DDA_INFO_COARSEST_REPLICATED) return (s.setup_done && s.cst.active && s.cst.replicated) ? 1 : 0;
  if (what == DDA_INFO_EMULATION) {
#ifdef DDA_HOST_EMU
    return 1;
#else
    return 0;
#endif
  }
  if (depth < 0 || depth >= nl) return -1;
  Level &L = s.lev[depth];
  switch (what) {
    case DDA_INFO_SITES: return (int)L.geo.V;
    case DDA_INFO_SITE_VARS: return L.geo.nc;
    case DDA_INFO_TEST_VECTORS: return L.nv;
    case DDA_INFO_BLOCK_SITES: return L.geo.bs;
    case DDA_INFO_NUM_BLOCKS: return L.geo.nblocks;
    default: return -1;
  }
}

void dda_set_option(int what, double value) {
  need_init("dda_set_option");
  Solver &s = A->s;
  switch (what) {
    case DDA_OPT_USE_FAST: s.use_fast = (int)value; g_transfer_fast = (int)value; break;
    case DDA_OPT_PROFILE: s.profile = (int)value; break;
    case DDA_OPT_SEED: s.seed = (unsigned long long)value; break;
    case DDA_OPT_PRINT: s.p.print = (int)value; break;
    default: fprintf(stderr, "dda_set_option: unknown option %d\n", what); fatal("API misuse", __FILE__, __LINE__);
  }
}

double dda_get_stat(int what) {
  Solver *s = A ? &A->s : nullptr;
  switch (what) {
    case DDA_STAT_LAUNCHES: return (double)g_launch_count;
    case DDA_STAT_DEVICE_BYTES: return (double)dev_bytes_in_use();
    case DDA_STAT_PLAQUETTE: return s ? s->plaq : 0.0;
    case DDA_STAT_ITER: return s ? (double)s->iter_count : 0.0;
    case DDA_STAT_COARSE_ITER: return s ? (double)s->coarse_iter_count : 0.0;
    case DDA_STAT_T_SMOOTH0: case DDA_STAT_T_SMOOTH0 + 1: case DDA_STAT_T_SMOOTH0 + 2: case DDA_STAT_T_SMOOTH0 + 3:
      return s ? s->t_smooth[what - DDA_STAT_T_SMOOTH0] : 0.0;
    case DDA_STAT_T_OP0: case DDA_STAT_T_OP0 + 1: case DDA_STAT_T_OP0 + 2: case DDA_STAT_T_OP0 + 3:
      return s ? s->t_op[what - DDA_STAT_T_OP0] : 0.0;
    case DDA_STAT_T_COARSEST: return s ? s->t_coarse_solve : 0.0;
    case DDA_STAT_T_RESTRICT: return s ? s->t_restrict : 0.0;
    case DDA_STAT_T_INTERPOLATE: return s ? s->t_interp : 0.0;
    default: return 0.0;
  }
}

void dda_reset_stats(void) {
  g_launch_count = 0;
  if (!A) return;
  Solver &s = A->s;
  for (int i = 0; i < MAX_LEVELS; i++) { s.t_smooth[i] = 0; s.t_op[i] = 0; }
  s.t_coarse_solve = s.t_restrict = s.t_interp = 0;
}

void dda_apply_dw(int precision, double *out_lex, const double *in_lex) {
  need_init("dda_apply_dw");
  Solver &s = A->s; Level &L0 = s.lev[0];
  DDA_ASSERT(s.conf_set);
  long V = L0.geo.V, n = 12 * V;
  h2d(s.lexbuf, in_lex, sizeof(cd) * n);
  if (precision == DDA_DOUBLE) {
    spinor_from_lex<double>(L0.geo, s.xb, s.lexbuf, 12);
    solver_apply_dw<double>(s, s.xx, s.xb);
    spinor_to_lex<double>(L0.geo, s.lexbuf, s.xx, 12);
  } else {
    cf *a = (cf *)s.xb, *b = (cf *)s.xx;
    spinor_from_lex<float>(L0.geo, a, s.lexbuf, 12);
    solver_apply_dw<float>(s, b, a);
    spinor_to_lex<float>(L0.geo, s.lexbuf, b, 12);
  }
  d2h(out_lex, s.lexbuf, sizeof(cd) * n);
}

void dda_get_operator(double *D_lex, double *clover_lex) {
  need_init("dda_get_operator");
  Solver &s = A->s;
  solver_sync_host_mirrors(s, false);
  if (D_lex) memcpy(D_lex, s.h_gauge.data(), s.h_gauge.size() * sizeof(double));
  if (clover_lex) memcpy(clover_lex, s.h_clover.data(), s.h_clover.size() * sizeof(double));
}

static Level &setup_level(int depth, const char *fn) {
  need_init(fn);
  Solver &s = A->s;
  if (!s.setup_done || depth < 0 || depth >= s.nlev) { fprintf(stderr, "%s: level %d not available (setup done: %d)\n", fn, depth, (int)s.setup_done); fatal("API misuse", __FILE__, __LINE__); }
  return s.lev[depth];
}

void dda_set_interpolation(int depth, const float *P_lex) {
  Level &L = setup_level(depth, "dda_set_interpolation");
  DDA_ASSERT(!L.last);
  long V = L.geo.V; int nc = L.geo.nc, nv = L.nv;
  std::vector<float> one((size_t)2 * V * nc);
  for (int k = 0; k < nv; k++) {
    for (long i = 0; i < V * nc; i++) { one[2 * i] = P_lex[2 * (i * nv + k)]; one[2 * i + 1] = P_lex[2 * (i * nv + k) + 1]; }
    level_upload(L, L.P[k], one.data());
  }
  for (int d = depth; d + 1 < A->s.nlev; d++) mg_rebuild_coarse(A->s, d);
  dev_sync();
}

void dda_get_interpolation(int depth, float *P_lex) {
  Level &L = setup_level(depth, "dda_get_interpolation");
  DDA_ASSERT(!L.last);
  long V = L.geo.V; int nc = L.geo.nc, nv = L.nv;
  std::vector<float> one((size_t)2 * V * nc);
  for (int k = 0; k < nv; k++) {
    level_download(L, one.data(), L.P[k]);
    for (long i = 0; i < V * nc; i++) { P_lex[2 * (i * nv + k)] = one[2 * i]; P_lex[2 * (i * nv + k) + 1] = one[2 * i + 1]; }
  }
}

void dda_level_op(int op, int depth, float *out_lex, const float *in_lex, int iparam, int flag) {
  Level &L = setup_level(depth, "dda_level_op");
  Solver &s = A->s;
  switch (op) {
    case DDA_OP_APPLY: {
      level_upload(L, L.vb, in_lex);
      mg_apply_op(s, depth, L.vx, L.vb);
      level_download(L, out_lex, L.vx);
    } break;
    case DDA_OP_RESTRICT: {
      DDA_ASSERT(!L.last);
      Level &N = s.lev[depth + 1];
      level_upload(L, L.vb, in_lex);
      tr_restrict(L.tr, N.vb, N.geo.nc, 0, L.vb, L.tr_scratch);
      level_download(N, out_lex, N.vb);
    } break;
    case DDA_OP_INTERPOLATE: {
      DDA_ASSERT(!L.last);
      Level &N = s.lev[depth + 1];
      level_upload(N, N.vx, in_lex);
      tr_interpolate(L.tr, L.vx, N.vx, false);
      level_download(L, out_lex, L.vx);
    } break;
    case DDA_OP_SMOOTHER: {
      DDA_ASSERT(!L.last);
      level_upload(L, L.vb, in_lex);
      if (flag) level_upload(L, L.vx, out_lex);
      mg_smoother(s, depth, L.vx, L.vb, iparam, !flag);
      level_download(L, out_lex, L.vx);
    } break;
    case DDA_OP_VCYCLE: {
      DDA_ASSERT(!L.last);
      level_upload(L, L.vb, in_lex);
      cf *phi = L.w[9];
      mg_vcycle(s, depth, phi, L.vb, true);
      level_download(L, out_lex, phi);
    } break;
    case DDA_OP_COARSEST_SOLVE: {
      DDA_ASSERT(L.last);
      level_upload(L, L.vb, in_lex);
      s.coarse_iter_count = 0;
      mg_coarsest_solve(s);
      level_download(L, out_lex, L.vx);
    } break;
    default: fprintf(stderr, "dda_level_op: unknown op %d\n", op); fatal("API misuse", __FILE__, __LINE__);
  }
}

int dda_level_apply_mrhs(int depth, float *out_lex, const float *in_lex, int reps, double *ms_out) {
  Level &L = setup_level(depth, "dda_level_apply_mrhs");
  DDA_ASSERT(depth >= 1);
#ifdef DDA_HOST_EMU
  (void)out_lex; (void)in_lex; (void)reps; (void)ms_out; (void)L;
  return -1;
#else
  const int NR = 12;
  const long vs = L.geo.valloc(), zs = L.geo.V * 4 * L.geo.nc, nloc = L.geo.vlen();
  cf *vin = dev_alloc<cf>(NR * vs), *vout = dev_alloc<cf>(NR * vs), *Zs = dev_alloc<cf>(NR * zs);
  for (int j = 0; j < NR; j++) {
    level_upload(L, vin + j * vs, in_lex + 2 * j * nloc);
    lv_halo(L, vin + j * vs);
  }
  float *T = coarse_mrhs_tile(L.cop);                  // operator images: once per operator, outside the timed region
  int rc = coarse_apply_mrhs(L.cop, T, vout, vin, Zs, vs, zs) ? 0 : -1;
  dev_sync();
  if (rc == 0) {
    for (int j = 0; j < NR; j++) level_download(L, out_lex + 2 * j * nloc, vout + j * vs);
    if (reps > 0 && ms_out) {
      cudaEvent_t e0, e1;
      CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, g_stream));
      for (int r = 0; r < reps; r++) coarse_apply_mrhs(L.cop, T, vout, vin, Zs, vs, zs);
      CUDA_CHECK(cudaEventRecord(e1, g_stream));
      CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1));
      *ms_out = (double)ms / reps;
    }
  }
  dev_sync();
  dev_free(vin); dev_free(vout); dev_free(Zs);
  if (T) dev_free(T);
  return rc;
#endif
}

// device-resident timing of one hot-path operator: `reps` back-to-back applications, returns milliseconds per
// application measured on the launching stream (CUDA events; wall clock around a stream synchronisation in the
// emulation build)
double dda_bench_op(int op, int depth, int reps) {
  need_init("dda_bench_op");
  Solver &s = A->s;
  Level &L = s.lev[depth];
  const long n = L.geo.vlen();
  auto run = [&]() {
    switch (op) {
      case DDA_BENCH_DW_DOUBLE: solver_apply_dw<double>(s, s.xx, s.xb); break;
      case DDA_BENCH_DW_FLOAT: solver_apply_dw<float>(s, (cf *)s.xx, (cf *)s.xb); break;
      case DDA_BENCH_LEVEL_APPLY: mg_apply_op(s, depth, L.vx, L.vb); break;
      case DDA_BENCH_RESTRICT: tr_restrict(L.tr, s.lev[depth + 1].vb, s.lev[depth + 1].geo.nc, 0, L.vb, L.tr_scratch); break;
      case DDA_BENCH_INTERPOLATE: tr_interpolate(L.tr, L.vx, s.lev[depth + 1].vx, false); break;
      case DDA_BENCH_SMOOTHER: mg_smoother(s, depth, L.vx, L.vb, s.p.post_smooth_iter[depth], false); break;
      case DDA_BENCH_VCYCLE: mg_vcycle(s, depth, L.w[9], L.vb, true); break;
      case DDA_BENCH_COARSEST_SCHUR: mg_coarsest_schur(s, s.cst.x, s.cst.b, nullptr); break;
      default: fatal("dda_bench_op: unknown op", __FILE__, __LINE__);
    }
  };
  // deterministic non-trivial input
  if (op == DDA_BENCH_DW_DOUBLE) { cd *x = s.xb; launch_n(n, DLAMBDA(long i) { x[i] = cd(1.0 + 1e-3 * (double)(i % 97), 0.5 - 1e-3 * (double)(i % 89)); }); }
  else if (op == DDA_BENCH_DW_FLOAT) { cf *x = (cf *)s.xb; launch_n(n, DLAMBDA(long i) { x[i] = cf(1.f + 1e-3f * (float)(i % 97), 0.5f - 1e-3f * (float)(i % 89)); }); }
  else if (op == DDA_BENCH_COARSEST_SCHUR) {
    DDA_ASSERT(s.setup_done && s.cst.active && s.p.odd_even);
    cf *x = s.cst.b; const long ne = s.cst.op->n_even * s.cst.op->n;
    launch_n(ne, DLAMBDA(long i) { x[i] = cf(1.f + 1e-3f * (float)(i % 97), 0.5f - 1e-3f * (float)(i % 89)); });
  } else {
    DDA_ASSERT(s.setup_done);
    cf *x = L.vb; launch_n(n, DLAMBDA(long i) { x[i] = cf(1.f + 1e-3f * (float)(i % 97), 0.5f - 1e-3f * (float)(i % 89)); });
    if (op == DDA_BENCH_INTERPOLATE) { Level &N = s.lev[depth + 1]; cf *y = N.vx; launch_n(N.geo.vlen(), DLAMBDA(long i) { y[i] = cf(1.f + 1e-3f * (float)(i % 97), 0.5f - 1e-3f * (float)(i % 89)); }); }
    if (op == DDA_BENCH_SMOOTHER) vzero(L.vx, n);
  }
  run();
  dev_sync();
#ifndef DDA_HOST_EMU
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, g_stream));
  for (int r = 0; r < reps; r++) run();
  CUDA_CHECK(cudaEventRecord(e1, g_stream));
  CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1));
  return (double)ms / reps;
#else
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; r++) run();
  dev_sync();
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
#endif
}

// device-resident solve: source already uploaded by dda_upload_source, solution stays on the device
void dda_upload_source(const double *in_lex) {
  need_init("dda_upload_source");
  Solver &s = A->s; Level &L0 = s.lev[0];
  h2d(s.lexbuf, in_lex, sizeof(cd) * 12 * L0.geo.V);
  spinor_from_lex<double>(L0.geo, s.xb, s.lexbuf, 12);
  dev_sync();
}
double dda_solve_device(double tol, int *status, double *ms_out) {
  need_init("dda_solve_device");
  Solver &s = A->s;
  DDA_ASSERT(s.setup_done);
  dev_sync();
  auto t0 = std::chrono::steady_clock::now();
#ifndef DDA_HOST_EMU
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, g_stream));
#endif
  double res = mg_solve(s, s.xx, s.xb, tol, status);
#ifndef DDA_HOST_EMU
  CUDA_CHECK(cudaEventRecord(e1, g_stream));
  CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1));
  if (ms_out) *ms_out = ms;
  (void)t0;
#else
  if (ms_out) *ms_out = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
#endif
  return res;
}
void dda_download_solution(double *out_lex) {
  need_init("dda_download_solution");
  Solver &s = A->s; Level &L0 = s.lev[0];
  spinor_to_lex<double>(L0.geo, s.lexbuf, s.xx, 12);
  d2h(out_lex, s.lexbuf, sizeof(cd) * 12 * L0.geo.V);
}

}  // extern "C"

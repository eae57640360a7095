// solver.h -- state of the multigrid solver: levels, operators, work vectors.
#pragma once
#include "common.cuh"
#include "lattice.h"
#include "params.h"
#include "fine_op.h"
#include "coarse_op.h"
#include "transfer.h"
#include "krylov.h"
#include "dev_gmres.h"
#include <chrono>

namespace dda {

const int NWORK = 10;

struct Level {
  int depth = 0;
  Geometry geo;
  int nv = 0;             // test vectors of this level (columns of the interpolation to depth+1)
  bool last = false;
  // ---- fine level (depth 0) operator storage
  cd *Dd = nullptr; double *Cd = nullptr;                      // double operator (outer solver)
  cf *Df = nullptr; float *Cf = nullptr, *Cinvf = nullptr;     // float copy used inside the cycle
  cf *Dblk = nullptr;                    // 4^4 even-odd blocks: per block the 768 in-block links [mu][3*row+col][slot], the
                                         // shared-memory image of the fused SAP kernel (one TMA bulk copy per block visit)
  FineOp<double> opd; FineOp<float> opf;
  // ---- coarse levels (depth >= 1)
  CoarseOp cop;
  cf *copZ = nullptr;                    // scratch of the scatter-form coarse apply (4 n complex per site)
  // ---- transfer to depth+1
  std::vector<cf *> tv;   // test vectors
  std::vector<cf *> P;    // interpolation vectors = aggregate/chirality-orthonormalised test vectors
  Transfer tr;
  double *tr_scratch = nullptr;
  // ---- cycle work vectors (float, native layout, length geo.vlen())
  cf *vb = nullptr, *vx = nullptr;        // right-hand side / solution of this level inside the cycle
  cf *w[NWORK] = {};
  double *blockred = nullptr;             // per-block reduction scratch of the SAP block solver
  Fgmres<float> kc;                       // K-cycle wrapper (depth >= 1, not last) or coarsest-level solver (last)
};

// Coarsest-level solver state.  With several ranks (or forced ghost slabs) and a small coarsest lattice, every rank
// holds a copy of the WHOLE coarsest operator and solves the gathered system redundantly: one all-gather of the
// right-hand side per solve instead of two halo exchanges and two all-reduces per Arnoldi step on a few hundred sites
// per rank.  This is the limit case of the reference's coarse-level gathering onto fewer processes
// (gathering_generic.c:44-194, 285-346: vector_PRECISION_gather / distribute, idle ranks).
struct Coarsest {
  bool active = false, replicated = false, fast = false, host_driven = false;
  Geometry rgeo;                          // replicated: global coarsest lattice, no ghost slabs
  CoarseOp rop;                           // replicated: operator of the global lattice (own storage)
  Geometry *geo = nullptr;                // -> rgeo or the last level's geometry
  CoarseOp *op = nullptr;                 // -> rop or the last level's operator
  cf *b = nullptr, *x = nullptr;          // right-hand side / solution in `geo` order
  cf *t[4] = {};                          // work vectors (full lattice + ghosts)
  cf *dir = nullptr, *Z = nullptr;        // scratch of the scatter-form Schur kernels
  cf *gbuf = nullptr;                     // all-gather staging of a vector (replicated)
  unsigned *counter = nullptr;            // ticket of the fused orthogonalisation kernel (last CTA runs the Givens step)
  int *d_src = nullptr, *d_own = nullptr; // replicated: gather source of global site g; global site of local site k
  DevGmres dg;                            // device-resident GMRES (default)
  Fgmres<float> hostk;                    // host-driven GMRES (DDA_COARSEST_HOST=1: round-1 behaviour, for comparison)
};

struct Solver {
  Params p;
  int nlev = 0;
  Level lev[MAX_LEVELS];
  bool fine_alloc = false, conf_set = false, setup_done = false;
  double plaq = 0.0;
  double m0_op = 0.0;                     // mass currently folded into the clover diagonal
  Coarsest cst;
  Fgmres<double> outer;                   // outer double-precision FGMRES ("mixed precision: 0/1")
  FgmresMP outer_mp;                      // mixed-precision outer solver ("mixed precision: 2", linsolve.c:153)
  cd *xb = nullptr, *xx = nullptr;        // device source / solution of the outer solve (native layout)
  cd *lexbuf = nullptr;                   // device staging buffer, lexicographic (36 complex per site)
  long coarse_iter_count = 0, iter_count = 0;
  double norm_res = 0.0;
  int use_fast = 1;                       // 1: hand-tuned kernels where available, 0: generic kernels only
  unsigned long long seed = 0;
  // host mirrors for the raw-pointer API (reference dirac.c:171-176), lexicographic
  std::vector<double> h_gauge, h_clover;
  // profiling (seconds; device synchronisation only when profile != 0)
  int profile = 0;
  double t_smooth[MAX_LEVELS] = {}, t_coarse_solve = 0, t_restrict = 0, t_interp = 0, t_op[MAX_LEVELS] = {};
};

extern Solver *g_solver;

// device-synchronising wall-clock timer of one operator class (only when Solver::profile is set)
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct ProfScope {
  Solver &s; double *acc; double t0;
  ProfScope(Solver &s_, double *a) : s(s_), acc(a), t0(0) { if (s.profile) { dev_sync(); t0 = now_s(); } }
  ~ProfScope() { if (s.profile) { dev_sync(); *acc += now_s() - t0; } }
};

// ---- fine operator management (solver_fine.cu)
void solver_process_grid(Solver &s, int depth, Geometry &g);
void solver_alloc_fine(Solver &s);
void solver_free_fine(Solver &s);
void solver_upload_conf(Solver &s, const double *gauge_lex);   // [site lex][mu][3][3][2] doubles (U, not U/2)
void solver_sync_host_mirrors(Solver &s, bool to_device);      // gauge/clover pointer API
void solver_refresh_float_op(Solver &s);                        // double op -> float op + clover inverse
void solver_shift_mass(Solver &s, double new_m0);               // all levels (reference shift_update, dirac.c:669-691)
template <class T> void solver_apply_dw(Solver &s, cx<T> *out, const cx<T> *in);

// ---- multigrid (mg_setup.cu, mg_cycle.cu)
void mg_alloc(Solver &s);
void mg_setup(Solver &s, int setup_iters);
void mg_setup_update(Solver &s, int setup_iters);
void mg_rebuild_coarse(Solver &s, int depth);                  // Galerkin operator of level depth+1 (and below) from P
void mg_free(Solver &s);
void mg_resetup_from_test_vectors(Solver &s);                 // P and the coarse operators from the current test vectors
// test-vector files in the reference's vector format (io.cu)
void tv_write(Solver &s, const char *basename);
void tv_read(Solver &s, const char *basename);
void mg_preconditioner(Solver &s, cd *out, const cd *in);     // one V/K-cycle on the fine level, double in/out
void mg_vcycle(Solver &s, int depth, cf *phi, const cf *eta, bool zero_guess);
void mg_smoother(Solver &s, int depth, cf *phi, const cf *eta, int iters, bool zero_guess);
void mg_coarsest_solve(Solver &s);
// even-site Schur complement of the coarsest operator (vectors in the coarsest solver's geometry, Coarsest::geo)
void mg_coarsest_schur(Solver &s, cf *out, const cf *in, const int *skip = nullptr);
void coarsest_alloc(Solver &s);          // after the levels' geometries exist
void coarsest_free(Solver &s);
void coarsest_refresh(Solver &s);        // after the last level's operator changed: gather (if replicated) + Soo^-1
void mg_apply_op(Solver &s, int depth, cf *out, const cf *in);  // full operator of the level (float)
double mg_solve(Solver &s, cd *x, const cd *b, double tol, int *status);
// fills the ghost slabs of a level's float vector (no-op on an unpartitioned level)
void lv_halo(Level &L, const cf *v);
// generic operator dispatch of a level (float); hops that cross the rank boundary read the ghost slabs of `in`
void lv_apply(Level &L, cf *out, const cf *in, SiteSel sel, int hop, int dir, int self, int outmode,
              const cf *eta = nullptr, const cf *in_self = nullptr);

}  // namespace dda

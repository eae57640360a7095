// solver.h -- state of the multigrid solver: levels, operators, work vectors.
#pragma once
#include "common.cuh"
#include "lattice.h"
#include "params.h"
#include "fine_op.h"
#include "coarse_op.h"
#include "transfer.h"
#include "krylov.h"

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

struct Solver {
  Params p;
  int nlev = 0;
  Level lev[MAX_LEVELS];
  bool fine_alloc = false, conf_set = false, setup_done = false;
  double plaq = 0.0;
  double m0_op = 0.0;                     // mass currently folded into the clover diagonal
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
void mg_preconditioner(Solver &s, cd *out, const cd *in);     // one V/K-cycle on the fine level, double in/out
void mg_vcycle(Solver &s, int depth, cf *phi, const cf *eta, bool zero_guess);
void mg_smoother(Solver &s, int depth, cf *phi, const cf *eta, int iters, bool zero_guess);
void mg_coarsest_solve(Solver &s);
void mg_coarsest_schur(Solver &s, cf *out, const cf *in);   // even-site Schur complement of the coarsest operator
void mg_apply_op(Solver &s, int depth, cf *out, const cf *in);  // full operator of the level (float)
double mg_solve(Solver &s, cd *x, const cd *b, double tol, int *status);
// fills the ghost slabs of a level's float vector (no-op on an unpartitioned level)
void lv_halo(Level &L, const cf *v);
// generic operator dispatch of a level (float); hops that cross the rank boundary read the ghost slabs of `in`
void lv_apply(Level &L, cf *out, const cf *in, SiteSel sel, int hop, int dir, int self, int outmode,
              const cf *eta = nullptr, const cf *in_self = nullptr);

}  // namespace dda

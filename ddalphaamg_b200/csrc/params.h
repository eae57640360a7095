// params.h -- runtime parameters of the solver: the reference's .ini keys (init.c:592-962) with the same
// defaults, or the struct route of include/dd_alpha_amg_parameters.h (init.c:817-901).
#pragma once
#include <string>

namespace dda {

const int MAX_LEVELS = 4;

struct Params {
  int num_levels = 2;
  int global_lattice[MAX_LEVELS][4] = {};
  int local_lattice[MAX_LEVELS][4] = {};
  int block_lattice[MAX_LEVELS][4] = {};
  int post_smooth_iter[MAX_LEVELS] = {2, 2, 2, 2};
  int ncycle[MAX_LEVELS] = {1, 1, 1, 1};
  double relax_fac[MAX_LEVELS] = {1.0, 1.0, 1.0, 1.0};
  int block_iter[MAX_LEVELS] = {4, 4, 4, 4};
  int setup_iter[MAX_LEVELS] = {6, 3, 2, 2};
  int num_eig_vect[MAX_LEVELS] = {20, 30, 30, 30};
  int anti_pbc = 0;
  int mixed_precision = 2;
  int interpolation = 2;
  int randomize = 0;
  int coarse_iter = 25, coarse_restart = 40;
  double coarse_tol = 5e-2;
  int odd_even = 1;
  double m0 = 0.0, csw = 0.0, setup_m0 = 0.0;
  int method = 2;
  int restart = 10, max_restart = 100;
  double tol = 1e-10;
  int print = 0;
  int kcycle = 1, kcycle_restart = 5, kcycle_max_restart = 2;
  double kcycle_tol = 1e-1;
  int rhs = 1;
  std::string conf_path;
  std::string tv_file;       // "test vector io file name:" (interpolation: 4)
};

// parse the reference's "key: value" format; aborts (reference error0 convention) on a missing mandatory key
void params_from_ini(Params &p, const char *path);
// derive coarse lattices / block lattices the way read_geometry_data does (init.c:651-760) and validate
// (init.c:964-1046)
void params_finalize(Params &p);

}  // namespace dda

// params.cu -- .ini reader and geometry derivation (host only).
#include "params.h"
#include "common.cuh"
#include <fstream>
#include <sstream>

namespace dda {

namespace {
struct Ini {
  std::vector<std::string> lines;
  // reference semantics (init.c:448-531): first line that CONTAINS the key wins, the value follows the key
  bool find(const std::string &key, std::string &rest) const {
    for (const auto &l : lines) {
      size_t p = l.find(key);
      if (p != std::string::npos) { rest = l.substr(p + key.size()); return true; }
    }
    return false;
  }
  bool get_ints(const std::string &key, int *out, int n) const {
    std::string r; if (!find(key, r)) return false;
    std::istringstream is(r);
    for (int i = 0; i < n; i++) { if (!(is >> out[i])) { fprintf(stderr, "bad value for \"%s\"\n", key.c_str()); fatal("parameter file", __FILE__, __LINE__); } }
    return true;
  }
  bool get_double(const std::string &key, double *out) const {
    std::string r; if (!find(key, r)) return false;
    std::istringstream is(r);
    if (!(is >> *out)) { fprintf(stderr, "bad value for \"%s\"\n", key.c_str()); fatal("parameter file", __FILE__, __LINE__); }
    return true;
  }
  void need(bool ok, const std::string &key) const {
    if (!ok) { fprintf(stderr, "unable to find string \"%s\" --- fatal error\n", key.c_str()); fatal("parameter file", __FILE__, __LINE__); }
  }
};
}  // namespace

void params_from_ini(Params &p, const char *path) {
  Ini ini;
  std::ifstream f(path);
  if (!f.good()) { fprintf(stderr, "cannot open parameter file %s\n", path); fatal("parameter file", __FILE__, __LINE__); }
  std::string line;
  while (std::getline(f, line)) ini.lines.push_back(line);
  std::string s;
  if (ini.find("configuration:", s)) { size_t a = s.find_first_not_of(' '); p.conf_path = a == std::string::npos ? "" : s.substr(a); }
  ini.get_ints("right hand side:", &p.rhs, 1);
  ini.get_ints("number of levels:", &p.num_levels, 1);
  DDA_ASSERT(p.num_levels >= 1 && p.num_levels <= MAX_LEVELS);
  ini.get_ints("antiperiodic boundary conditions:", &p.anti_pbc, 1);
  // solver parameters first (odd_even / method influence the geometry derivation)
  ini.get_ints("mixed precision:", &p.mixed_precision, 1);
  if (p.num_levels == 1) p.interpolation = 0; else ini.get_ints("interpolation:", &p.interpolation, 1);
  if (p.interpolation == 4) {   // read_testvector_io_data_if_necessary (init.c:904-912); one file per vector, "<name>.NN"
    std::string tv;
    ini.need(ini.find("test vector io file name:", tv), "test vector io file name:");
    size_t a = tv.find_first_not_of(' ');
    p.tv_file = a == std::string::npos ? "" : tv.substr(a);
    while (!p.tv_file.empty() && (p.tv_file.back() == ' ' || p.tv_file.back() == '\r')) p.tv_file.pop_back();
  }
  ini.get_ints("randomize test vectors:", &p.randomize, 1);
  ini.get_ints("coarse grid iterations:", &p.coarse_iter, 1);
  ini.get_ints("coarse grid restarts:", &p.coarse_restart, 1);
  ini.get_double("coarse grid tolerance:", &p.coarse_tol);
  ini.get_ints("odd even preconditioning:", &p.odd_even, 1);
  ini.need(ini.get_double("m0:", &p.m0), "m0:");
  ini.get_double("solver m0:", &p.m0);
  ini.need(ini.get_double("csw:", &p.csw), "csw:");
  p.setup_m0 = p.m0;
  ini.get_double("setup m0:", &p.setup_m0);
  ini.get_ints("method:", &p.method, 1);
  ini.get_ints("iterations between restarts:", &p.restart, 1);
  ini.get_ints("maximum of restarts:", &p.max_restart, 1);
  ini.get_double("tolerance for relative residual:", &p.tol);
  ini.get_ints("print mode:", &p.print, 1);
  ini.get_ints("kcycle:", &p.kcycle, 1);
  ini.get_ints("kcycle length:", &p.kcycle_restart, 1);
  ini.get_ints("kcycle restarts:", &p.kcycle_max_restart, 1);
  ini.get_double("kcycle tolerance:", &p.kcycle_tol);

  int ls = p.num_levels < 2 ? 2 : p.num_levels;
  for (int i = 0; i < ls && i < MAX_LEVELS; i++) {
    char key[64];
    snprintf(key, sizeof key, "d%d global lattice:", i);
    bool ok = ini.get_ints(key, p.global_lattice[i], 4);
    if (i == 0) ini.need(ok, key); else if (!ok) p.global_lattice[i][0] = 0;   // derived in params_finalize
    snprintf(key, sizeof key, "d%d local lattice:", i);
    ok = ini.get_ints(key, p.local_lattice[i], 4);
    if (i == 0) ini.need(ok, key); else if (!ok) p.local_lattice[i][0] = 0;
    snprintf(key, sizeof key, "d%d block lattice:", i);
    ok = ini.get_ints(key, p.block_lattice[i], 4);
    if (i == 0 && p.num_levels > 1) ini.need(ok, key); else if (!ok) p.block_lattice[i][0] = 0;
    snprintf(key, sizeof key, "d%d post smooth iter:", i); ini.get_ints(key, &p.post_smooth_iter[i], 1);
    snprintf(key, sizeof key, "d%d preconditioner cycles:", i); ini.get_ints(key, &p.ncycle[i], 1);
    snprintf(key, sizeof key, "d%d relaxation factor:", i); ini.get_double(key, &p.relax_fac[i]);
    snprintf(key, sizeof key, "d%d block iter:", i); ini.get_ints(key, &p.block_iter[i], 1);
    snprintf(key, sizeof key, "d%d setup iter:", i); ini.get_ints(key, &p.setup_iter[i], 1);
    snprintf(key, sizeof key, "d%d test vectors:", i);
    if (i > 0) p.num_eig_vect[i] = (int)(1.5 * p.num_eig_vect[0]);
    ini.get_ints(key, &p.num_eig_vect[i], 1);
  }
}

void params_finalize(Params &p) {
  int ls = p.num_levels;
  for (int i = 0; i < ls; i++) {
    if (i > 0) {
      if (p.global_lattice[i][0] == 0) for (int m = 0; m < 4; m++) p.global_lattice[i][m] = p.global_lattice[i - 1][m] / p.block_lattice[i - 1][m];
      if (p.local_lattice[i][0] == 0) for (int m = 0; m < 4; m++) p.local_lattice[i][m] = p.local_lattice[i - 1][m] / p.block_lattice[i - 1][m];
    }
    if (i < ls - 1) {
      if (p.block_lattice[i][0] == 0) {
        for (int m = 0; m < 4; m++) {
          if (p.global_lattice[i][m] % 2 == 0) p.block_lattice[i][m] = 2;
          else if (p.global_lattice[i][m] % 3 == 0) p.block_lattice[i][m] = 3;
          else { fprintf(stderr, "lattice dimensions not valid for a %d-level method\n", p.num_levels); fatal("geometry", __FILE__, __LINE__); }
        }
      }
    } else {
      for (int m = 0; m < 4; m++) p.block_lattice[i][m] = 1;
    }
  }
  // validation (init.c:964-1046)
  for (int i = 0; i < ls; i++) for (int m = 0; m < 4; m++) {
    DDA_ASSERT(p.local_lattice[i][m] > 0 && p.global_lattice[i][m] % p.local_lattice[i][m] == 0);
  }
  for (int i = 0; i + 1 < ls; i++) for (int m = 0; m < 4; m++) {
    DDA_ASSERT(p.global_lattice[i][m] % p.global_lattice[i + 1][m] == 0);
    DDA_ASSERT(p.local_lattice[i][m] % p.block_lattice[i][m] == 0);
    int agg = p.global_lattice[i][m] / p.global_lattice[i + 1][m];
    DDA_ASSERT(p.local_lattice[i][m] % agg == 0);
    DDA_ASSERT(agg % p.block_lattice[i][m] == 0);
  }
  for (int i = 0; i + 2 < ls; i++) DDA_ASSERT(p.num_eig_vect[i] <= p.num_eig_vect[i + 1]);
  if (p.odd_even && ls > 1) {
    long cs = 1;
    for (int m = 0; m < 4; m++) { DDA_ASSERT(p.global_lattice[ls - 1][m] % 2 == 0); cs *= p.local_lattice[ls - 1][m]; }
    DDA_ASSERT(cs % 2 == 0);
  }
  if (p.method == 2) for (int i = 0; i + 1 < ls; i++) {
    long nb = 1; for (int m = 0; m < 4; m++) nb *= p.local_lattice[i][m] / p.block_lattice[i][m];
    DDA_ASSERT(nb >= 2);
  }
  DDA_ASSERT(p.max_restart > 0 && p.tol > 0 && p.tol < 1);
  if (ls > 1) DDA_ASSERT(p.coarse_iter > 0 && p.coarse_restart > 0 && p.coarse_tol > 0 && p.coarse_tol < 1);
}

}  // namespace dda

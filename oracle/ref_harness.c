/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline), never linked into the product.
 *
 * Thin C accessors around the UNMODIFIED reference (mrottmann/DDalphaAMG, /root/reference) compiled
 * from where it lies by oracle/build_ref.sh into oracle/_ref/libddref.so.  This translation unit
 * textually includes the reference's own library front end (src/dd_alpha_amg.c, generated copy in
 * the temporary build dir) so that the file-static level hierarchy `l` is reachable; everything
 * below only *calls* reference functions, it does not restate them.
 *
 * Caller-side layouts used by every entry point here:
 *   gauge   : double [t][z][y][x][mu=T,Z,Y,X][3][3][re,im]      (the reference's native conf order)
 *   spinors : complex [t][z][y][x][12]  (lexicographic, x fastest; 3*spin+colour)
 *   coarse  : complex [lexicographic coarse site][2*Nv]
 */
#include "dd_alpha_amg.c"

static int LT, LZ, LY, LX;
static int h_conf_index(int t, int z, int y, int x, int mu) { return 18*(4*(x + LX*(y + LY*(z + LZ*t))) + mu); }
static int h_vec_index(int t, int z, int y, int x) { return 24*(x + LX*(y + LY*(z + LZ*t))); }
static int h_global_time(int t) { return t; }
static struct Thread **mt_threading = NULL;
static int mt_n = 0;

static level_struct *level_at(int depth) {
  level_struct *lp = &l;
  while (lp && lp->depth < depth) lp = lp->next_level;
  return lp;
}

/* reference entry: dd_alpha_amg_init (src/dd_alpha_amg.c:95) */
void ref_init(const char *ini_path, double m0, double csw, int bc, int print) {
  dd_alpha_amg_par p;
  memset(&p, 0, sizeof(p));
  strncpy(p.param_file_path, ini_path, STRINGLENGTH-1);
  p.conf_index_fct = h_conf_index; p.vector_index_fct = h_vec_index; p.global_time = h_global_time;
  p.bc = bc; p.m0 = m0; p.csw = csw; p.setup_m0 = m0;
  dd_alpha_amg_init(p);
  g.print = print;
  LT = l.local_lattice[T]; LZ = l.local_lattice[Z]; LY = l.local_lattice[Y]; LX = l.local_lattice[X];
}

/* reference entry: dd_alpha_amg_set_conf (src/dd_alpha_amg.c:188); anti-periodic sign must be in the links */
double ref_set_conf(double *gauge) { return dd_alpha_amg_set_conf(gauge); }

/* reference entry: dd_alpha_amg_setup (src/dd_alpha_amg.c:258), single thread */
void ref_setup(int iters, int *status) { dd_alpha_amg_setup(iters, status); }

/* multi-threaded setup following the stand-alone driver (src/main.c:99-109) */
void ref_setup_mt(int iters, int nthreads, int *status) {
  g.coarse_iter_count = 0;
  if (g.setup_flag) method_free(&l);
  g.num_openmp_processes = nthreads;
  mt_n = nthreads;
  mt_threading = (struct Thread**)malloc(sizeof(struct Thread*)*nthreads);
  for (int i = 0; i < nthreads; i++) mt_threading[i] = (struct Thread*)malloc(sizeof(struct Thread));
#pragma omp parallel num_threads(nthreads)
  {
    struct Thread *th = mt_threading[omp_get_thread_num()];
    setup_threading(th, commonthreaddata, &l);
    method_setup(NULL, &l, th);
    START_LOCKED_MASTER(th)
    g.setup_flag = 1;
    END_LOCKED_MASTER(th)
    method_update(iters, &l, th);
  }
  status[0] = 1; status[1] = g.coarse_iter_count; g.conf_flag = 0;
}

/* reference entry: dd_alpha_amg_wilson_solve (src/dd_alpha_amg.c:324) */
double ref_solve(double *out, double *in, double tol, int *status) {
  return dd_alpha_amg_wilson_solve(out, in, tol, 1.0, 1.0, status);
}

/* same entry with the clover scaling by site parity (scale_clover, src/dirac.c:646-667) */
double ref_solve_scaled(double *out, double *in, double tol, double scale_even, double scale_odd, int *status) {
  return dd_alpha_amg_wilson_solve(out, in, tol, scale_even, scale_odd, status);
}

/* mass shift on every level: shift_update (src/dirac.c:669-691), as run_dd_alpha_amg_setup_if_necessary does
 * (src/dd_alpha_amg.c:91-92) */
void ref_shift_mass(double m0) {
#pragma omp parallel num_threads(threading[0]->n_core)
  {
    shift_update((complex_double)m0, &l, threading[omp_get_thread_num()]);
  }
}

/* multi-threaded solve: wilson_driver (src/top_level.c:64) inside a parallel region as in src/main.c:99 */
double ref_solve_mt(double *out, double *in, double tol, int *status, double *seconds) {
  int n = 2*l.inner_vector_size;
  g.coarse_iter_count = 0; g.iter_count = 0; g.p.tol = tol; g.p_MP.dp.tol = tol;
  double t0 = MPI_Wtime();
#pragma omp parallel num_threads(mt_n)
  {
    wilson_driver((vector_double)out, (vector_double)in, &l, mt_threading[omp_get_thread_num()]);
  }
  *seconds = MPI_Wtime() - t0;
  (void)n;
  status[0] = g.iter_count; status[1] = g.coarse_iter_count;
  if (g.norm_res > tol) status[0] = -1;
  return g.norm_res;
}

/* dd_alpha_amg_free (src/dd_alpha_amg.c:398) is only safe after a setup; without one the instance is abandoned
 * (test infrastructure: the leak is accepted) */
void ref_free(void) { if (g.setup_flag) dd_alpha_amg_free(); }

/* what: 0 num_levels, 1 num inner sites(depth), 2 site vars(depth), 3 num_eig_vect(depth),
 *       4 vector_size incl. ghost shell(depth), 5 schwarz_vector_size, 6 num aggregates(depth), 7.. lattice dims */
int ref_info(int what, int depth) {
  level_struct *lp = level_at(depth);
  if (what == 0) return g.num_levels;
  if (!lp) return -1;
  switch (what) {
    case 1: return lp->num_inner_lattice_sites;
    case 2: return lp->num_lattice_site_var;
    case 3: return lp->num_eig_vect;
    case 4: return lp->vector_size;
    case 5: return lp->schwarz_vector_size;
    case 6: return lp->is_float.num_agg;
    case 7: case 8: case 9: case 10: return lp->local_lattice[what-7];
    case 11: case 12: case 13: case 14: return lp->block_lattice ? lp->block_lattice[what-11] : -1;
    case 15: case 16: case 17: case 18: return lp->coarsening[what-15];
    default: return -1;
  }
}

/* fine operator arrays of the reference, g.op_double (src/dirac.c:80, :386-398) */
void ref_get_D(double *out) { memcpy(out, g.op_double.D, sizeof(complex_double)*36*(size_t)l.num_inner_lattice_sites); }
void ref_get_clover(double *out) { memcpy(out, g.op_double.clover, sizeof(complex_double)*42*(size_t)l.num_inner_lattice_sites); }

/* d_plus_clover_double (src/dirac_generic.c:159), lexicographic in/out */
void ref_dw_double(double *out, const double *in) {
  vector_double a = NULL, b = NULL;
  MALLOC(a, complex_double, l.vector_size); MALLOC(b, complex_double, l.vector_size);
  memcpy(a, in, sizeof(complex_double)*(size_t)l.inner_vector_size);
  d_plus_clover_double(b, a, &(g.op_double), &l, no_threading);
  memcpy(out, b, sizeof(complex_double)*(size_t)l.inner_vector_size);
  FREE(a, complex_double, l.vector_size); FREE(b, complex_double, l.vector_size);
}

/* repeated applies for CPU timing; returns best seconds per apply (threads via the reference's own OpenMP split) */
double ref_dw_double_time(int reps, int nthreads) {
  vector_double a = NULL, b = NULL; double best = 1e30;
  MALLOC(a, complex_double, l.vector_size); MALLOC(b, complex_double, l.vector_size);
  for (int i = 0; i < l.inner_vector_size; i++) a[i] = 1.0 + 0.5*I;
  if (nthreads <= 1) {
    for (int r = 0; r < reps; r++) { double t = MPI_Wtime(); d_plus_clover_double(b, a, &(g.op_double), &l, no_threading); t = MPI_Wtime()-t; if (t < best) best = t; }
  } else {
    struct Thread **th = (struct Thread**)malloc(sizeof(struct Thread*)*nthreads);
    for (int i = 0; i < nthreads; i++) th[i] = (struct Thread*)malloc(sizeof(struct Thread));
    double *ts = (double*)malloc(sizeof(double)*reps);
#pragma omp parallel num_threads(nthreads)
    {
      struct Thread *me = th[omp_get_thread_num()];
      setup_threading(me, commonthreaddata, &l);
      for (int r = 0; r < reps; r++) {
#pragma omp barrier
        double t = MPI_Wtime();
        d_plus_clover_double(b, a, &(g.op_double), &l, me);
#pragma omp barrier
        if (omp_get_thread_num() == 0) ts[r] = MPI_Wtime()-t;
      }
    }
    for (int r = 0; r < reps; r++) if (ts[r] < best) best = ts[r];
    free(ts);
  }
  FREE(a, complex_double, l.vector_size); FREE(b, complex_double, l.vector_size);
  return best;
}

/* d_plus_clover_float on the Schwarz-ordered float operator (l.s_float.op), lexicographic double in/out through
 * trans_float / trans_back_float (src/schwarz_generic.c:1807-1846) -- needs ref_setup first. */
void ref_dw_float(double *out, const double *in) {
  vector_float a = NULL, b = NULL; vector_double di = NULL, dout = NULL;
  MALLOC(a, complex_float, l.schwarz_vector_size); MALLOC(b, complex_float, l.schwarz_vector_size);
  MALLOC(di, complex_double, l.vector_size); MALLOC(dout, complex_double, l.vector_size);
  memcpy(di, in, sizeof(complex_double)*(size_t)l.inner_vector_size);
  trans_float(a, di, l.s_float.op.translation_table, &l, no_threading);
  d_plus_clover_float(b, a, &(l.s_float.op), &l, no_threading);
  trans_back_float(dout, b, l.s_float.op.translation_table, &l, no_threading);
  memcpy(out, dout, sizeof(complex_double)*(size_t)l.inner_vector_size);
  FREE(a, complex_float, l.schwarz_vector_size); FREE(b, complex_float, l.schwarz_vector_size);
  FREE(di, complex_double, l.vector_size); FREE(dout, complex_double, l.vector_size);
}

/* preconditioner = one float V/K-cycle (src/preconditioner.c:25), lexicographic double in/out */
void ref_preconditioner(double *out, const double *in) {
  vector_double a = NULL, b = NULL;
  MALLOC(a, complex_double, l.vector_size); MALLOC(b, complex_double, l.vector_size);
  memcpy(a, in, sizeof(complex_double)*(size_t)l.inner_vector_size);
  preconditioner(b, NULL, a, _NO_RES, &l, no_threading);
  memcpy(out, b, sizeof(complex_double)*(size_t)l.inner_vector_size);
  FREE(a, complex_double, l.vector_size); FREE(b, complex_double, l.vector_size);
}

/* translation table lexicographic -> native (Schwarz) site order of level `depth` (src/data_layout.c:241-250) */
void ref_get_translation(int depth, int *out) {
  level_struct *lp = level_at(depth);
  memcpy(out, lp->s_float.op.translation_table, sizeof(int)*(size_t)lp->num_inner_lattice_sites);
}

/* prolongator of level `depth` (is_float.operator, src/interpolation_generic.c:74-90): [native fine dof][Nv] */
void ref_get_interpolation(int depth, float *out) {
  level_struct *lp = level_at(depth);
  memcpy(out, lp->is_float.operator, sizeof(complex_float)*(size_t)lp->inner_vector_size*(size_t)lp->num_eig_vect);
}

/* coarse-level vectors at depth>=1 live in the level's native order; helpers to go from/to lexicographic.
 * intermediate levels: Schwarz order (s_float.op.translation_table); coarsest: lexicographic (identity). */
/* native index of lexicographic site i: Schwarz order on intermediate levels; on the coarsest level with odd-even
 * preconditioning the operator and all vectors are in even-then-odd order (src/gathering_generic.c:155-180) */
static int *coarsest_eo_table(level_struct *lp) {
  static int *tab = NULL; static int tabn = 0;
  int *le = lp->local_lattice, ns = lp->num_inner_lattice_sites;
  if (tab && tabn == ns) return tab;
  if (tab) free(tab);
  tab = (int*)malloc(sizeof(int)*ns); tabn = ns;
  int i = 0;
  for (int par = 0; par < 2; par++)
    for (int t = 0; t < le[T]; t++) for (int z = 0; z < le[Z]; z++) for (int y = 0; y < le[Y]; y++) for (int x = 0; x < le[X]; x++)
      if ((t+z+y+x)%2 == par) tab[x + le[X]*(y + le[Y]*(z + le[Z]*t))] = i++;
  return tab;
}
static int native_index(level_struct *lp, int i) {
  if (lp->level > 0) return lp->s_float.op.translation_table ? lp->s_float.op.translation_table[i] : i;
  if (g.odd_even) return coarsest_eo_table(lp)[i];
  return i;
}
static void lex_to_native(level_struct *lp, complex_float *dst, const float *src) {
  int nv = lp->num_lattice_site_var, ns = lp->num_inner_lattice_sites;
  for (int i = 0; i < ns; i++) { int k = native_index(lp, i);
    for (int c = 0; c < nv; c++) dst[(size_t)k*nv+c] = src[2*((size_t)i*nv+c)] + I*src[2*((size_t)i*nv+c)+1]; }
}
static void native_to_lex(level_struct *lp, float *dst, const complex_float *src) {
  int nv = lp->num_lattice_site_var, ns = lp->num_inner_lattice_sites;
  for (int i = 0; i < ns; i++) { int k = native_index(lp, i);
    for (int c = 0; c < nv; c++) { dst[2*((size_t)i*nv+c)] = crealf(src[(size_t)k*nv+c]); dst[2*((size_t)i*nv+c)+1] = cimagf(src[(size_t)k*nv+c]); } }
}

/* apply_coarse_operator_float (src/coarse_operator_generic.c:383) at depth>=1, lexicographic in/out */
void ref_coarse_apply(int depth, float *out, const float *in) {
  level_struct *lp = level_at(depth);
  vector_float a = NULL, b = NULL;
  MALLOC(a, complex_float, lp->schwarz_vector_size); MALLOC(b, complex_float, lp->schwarz_vector_size);
  lex_to_native(lp, a, in);
  /* coarsest level with odd-even: the full operator is only available through its even-odd factorisation
   * (coarse_odd_even_PRECISION_test, src/coarse_oddeven_generic.c:1271-1319) */
  if (lp->level == 0 && g.odd_even) coarse_odd_even_float_test(b, a, lp, no_threading);
  else apply_coarse_operator_float(b, a, &(lp->s_float.op), lp, no_threading);
  native_to_lex(lp, out, b);
  FREE(a, complex_float, lp->schwarz_vector_size); FREE(b, complex_float, lp->schwarz_vector_size);
}

/* restrict_float (src/interpolation_generic.c:169): fine level `depth` (lexicographic) -> level depth+1 (lexicographic) */
void ref_restrict(int depth, float *coarse_out, const float *fine_in) {
  level_struct *lp = level_at(depth), *lc = lp->next_level;
  vector_float a = NULL, b = NULL;
  MALLOC(a, complex_float, lp->schwarz_vector_size); MALLOC(b, complex_float, lc->schwarz_vector_size);
  if (depth == 0) {
    int *tt = l.s_float.op.translation_table;
    for (int i = 0; i < l.num_inner_lattice_sites; i++) for (int c = 0; c < 12; c++)
      a[12*(size_t)tt[i]+c] = fine_in[2*(12*(size_t)i+c)] + I*fine_in[2*(12*(size_t)i+c)+1];
  } else lex_to_native(lp, a, fine_in);
  restrict_float(b, a, lp, no_threading);
  native_to_lex(lc, coarse_out, b);
  FREE(a, complex_float, lp->schwarz_vector_size); FREE(b, complex_float, lc->schwarz_vector_size);
}

/* interpolate3_float (src/interpolation_generic.c:130): level depth+1 (lex) -> level depth (lex) */
void ref_interpolate(int depth, float *fine_out, const float *coarse_in) {
  level_struct *lp = level_at(depth), *lc = lp->next_level;
  vector_float a = NULL, b = NULL;
  MALLOC(a, complex_float, lp->schwarz_vector_size); MALLOC(b, complex_float, lc->schwarz_vector_size);
  lex_to_native(lc, b, coarse_in);
  interpolate3_float(a, b, lp, no_threading);
  if (depth == 0) {
    int *tt = l.s_float.op.translation_table;
    for (int i = 0; i < l.num_inner_lattice_sites; i++) for (int c = 0; c < 12; c++) {
      fine_out[2*(12*(size_t)i+c)] = crealf(a[12*(size_t)tt[i]+c]); fine_out[2*(12*(size_t)i+c)+1] = cimagf(a[12*(size_t)tt[i]+c]); }
  } else native_to_lex(lp, fine_out, a);
  FREE(a, complex_float, lp->schwarz_vector_size); FREE(b, complex_float, lc->schwarz_vector_size);
}

/* smoother_float (src/vcycle_generic.c:25) = red_black_schwarz_float at level `depth`, n iterations.
 * use_res=1: phi_io holds the initial guess (lexicographic) ; use_res=0: zero initial guess. */
void ref_smoother(int depth, float *phi_io, const float *eta, int n, int use_res) {
  level_struct *lp = level_at(depth);
  vector_float p = NULL, e = NULL;
  MALLOC(p, complex_float, lp->schwarz_vector_size); MALLOC(e, complex_float, lp->schwarz_vector_size);
  if (depth == 0) {
    int *tt = l.s_float.op.translation_table;
    for (int i = 0; i < l.num_inner_lattice_sites; i++) for (int c = 0; c < 12; c++) {
      e[12*(size_t)tt[i]+c] = eta[2*(12*(size_t)i+c)] + I*eta[2*(12*(size_t)i+c)+1];
      p[12*(size_t)tt[i]+c] = phi_io[2*(12*(size_t)i+c)] + I*phi_io[2*(12*(size_t)i+c)+1]; }
  } else { lex_to_native(lp, e, eta); lex_to_native(lp, p, phi_io); }
  smoother_float(p, NULL, e, n, use_res ? _RES : _NO_RES, _NO_SHIFT, lp, no_threading);
  if (depth == 0) {
    int *tt = l.s_float.op.translation_table;
    for (int i = 0; i < l.num_inner_lattice_sites; i++) for (int c = 0; c < 12; c++) {
      phi_io[2*(12*(size_t)i+c)] = crealf(p[12*(size_t)tt[i]+c]); phi_io[2*(12*(size_t)i+c)+1] = cimagf(p[12*(size_t)tt[i]+c]); }
  } else native_to_lex(lp, phi_io, p);
  FREE(p, complex_float, lp->schwarz_vector_size); FREE(e, complex_float, lp->schwarz_vector_size);
}

/* coarse_solve_odd_even_float (src/coarse_oddeven_generic.c:1139) on the coarsest level, lexicographic in/out;
 * returns the number of GMRES iterations */
int ref_coarsest_solve(float *x_out, const float *b_in) {
  level_struct *lp = &l; while (lp->next_level) lp = lp->next_level;
  int before = g.coarse_iter_count;
  lex_to_native(lp, lp->p_float.b, b_in);
  if (g.odd_even) coarse_solve_odd_even_float(&(lp->p_float), &(lp->oe_op_float), lp, no_threading);
  else fgmres_float(&(lp->p_float), lp, no_threading);
  native_to_lex(lp, x_out, lp->p_float.x);
  return g.coarse_iter_count - before;
}

/* vcycle_float (src/vcycle_generic.c:91) at level `depth`, zero initial guess, lexicographic in/out */
void ref_vcycle(int depth, float *phi_out, const float *eta) {
  level_struct *lp = level_at(depth);
  vector_float p = NULL, e = NULL;
  MALLOC(p, complex_float, lp->schwarz_vector_size); MALLOC(e, complex_float, lp->schwarz_vector_size);
  lex_to_native(lp, e, eta);
  vcycle_float(p, NULL, e, _NO_RES, lp, no_threading);
  native_to_lex(lp, phi_out, p);
  FREE(p, complex_float, lp->schwarz_vector_size); FREE(e, complex_float, lp->schwarz_vector_size);
}

double ref_plaquette(void) { return g.plaq; }
double ref_norm_res(void) { return g.norm_res; }

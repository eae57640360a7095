/* dd_alpha_amg.h -- C library interface of the B200-native DDalphaAMG solve path (libdd_alpha_amg.so).
 *
 * Drop-in for the reference's library API: every entry point below has the prototype and the semantics of the
 * reference function of the same name (reference: include/dd_alpha_amg.h:29-83, implemented in
 * src/dd_alpha_amg.c:95-404).  Behind it the solve phase (and the setup that feeds it) runs on one NVIDIA B200
 * per process as hand-written sm_100a CUDA kernels; there is no CPU fallback.
 *
 * Conventions kept from the reference:
 *  - the caller owns gauge / source / solution arrays; the library copies in and out (dd_alpha_amg.c:193-204,
 *    :345-352, :375-382) through the caller's index functions (offsets in doubles);
 *  - SU(3) links are row-major 3x3 complex (18 doubles), spinors 12 complex = 24 doubles, index 3*spin+colour;
 *  - errors print and abort (reference error0 -> MPI_Abort, main.h:424-439); non-convergence is signalled only
 *    by status[0] = -1 (dd_alpha_amg.c:391-392);
 *  - one instance per process (the reference keeps process-global state, dd_alpha_amg.c:28-33).
 */
#ifndef DD_ALPHA_AMG_B200_INTERFACE
#define DD_ALPHA_AMG_B200_INTERFACE

#include "dd_alpha_amg_parameters.h"

#ifdef __cplusplus
extern "C" {
#endif

#define STRINGLENGTH 500

typedef struct {
  char param_file_path[STRINGLENGTH];
  int (*conf_index_fct)(int t, int z, int y, int x, int mu);
  int (*vector_index_fct)(int t, int z, int y, int x);
  int (*global_time)(int t);
  int bc; /* 0 dirichlet, 1 periodic, 2 anti-periodic */
  double m0;
  double csw;
  double setup_m0;
  struct dd_alpha_amg_parameters amg_params;
} dd_alpha_amg_par;

/* reference: include/dd_alpha_amg.h:43, src/dd_alpha_amg.c:95 -- geometry/solver parameters from the .ini at
 * p.param_file_path, csw/m0/setup_m0 from p */
void dd_alpha_amg_init(dd_alpha_amg_par p);
/* reference: include/dd_alpha_amg.h:44, src/dd_alpha_amg.c:135 -- geometry from p.amg_params; the threading
 * arguments are accepted and ignored (one host thread drives the GPU) */
void dd_alpha_amg_init_external_threading(dd_alpha_amg_par p, int n_core, int n_thread);

/* reference: include/dd_alpha_amg.h:46-47, src/dirac.c:171-176 -- host mirrors of D = U/2 (36 complex/site) and
 * of the clover term (42 complex/site), lexicographic; call dd_alpha_amg_fields_updated() after writing them */
double *dd_alpha_amg_get_gauge_pointer(void);
double *dd_alpha_amg_get_clover_pointer(void);
/* reference: include/dd_alpha_amg.h:56, src/dd_alpha_amg.c:182-185 */
void dd_alpha_amg_fields_updated(void);

/* reference: include/dd_alpha_amg.h:58, src/dd_alpha_amg.c:188-250 -- returns the average plaquette in [0,3] */
double dd_alpha_amg_set_conf(double *gauge_field);

/* reference: include/dd_alpha_amg.h:62, src/init.c:1139-1148 */
void dd_alpha_amg_update_parameters(const struct dd_alpha_amg_parameters *amg_params);

/* reference: include/dd_alpha_amg.h:64-70, src/dd_alpha_amg.c:258-321 -- status[0] = 1, status[1] = coarsest-level
 * iteration count */
void dd_alpha_amg_setup(int iterations, int *status);
void dd_alpha_amg_setup_external_threading(int iterations, int *status, int core, int thread,
                                           void *thread_barrier_data, void (*thread_barrier)(void *, int));
void dd_alpha_amg_setup_update(int iterations, int *status);
void dd_alpha_amg_setup_update_external_threading(int iterations, int *status, int core, int thread,
                                                  void *thread_barrier_data, void (*thread_barrier)(void *, int));

/* reference: include/dd_alpha_amg.h:72, src/dd_alpha_amg.c:324-395 -- returns the relative residual;
 * status[0] = outer iterations (-1 if not converged), status[1] = coarsest-level iterations */
double dd_alpha_amg_wilson_solve(double *vector_out, double *vector_in, double tol, double scale_even,
                                 double scale_odd, int *status);

/* reference: include/dd_alpha_amg.h:77-81 (declared there, defined nowhere in the reference) -- one V/K-cycle */
void dd_alpha_amg_preconditioner(double *vector_out, double *vector_in, double scale_even, double scale_odd,
                                 int *status);
void dd_alpha_amg_preconditioner_external_threading(double *vector_out, double *vector_in, int *status, int core,
                                                    int thread, void *thread_barrier_data,
                                                    void (*thread_barrier)(void *, int));

/* reference: include/dd_alpha_amg.h:83, src/dd_alpha_amg.c:398-404 */
void dd_alpha_amg_free(void);

/* Aliases under the names BASELINE.json's north_star uses (later upstream API names); thin wrappers. */
void DDalphaAMG_initialize(dd_alpha_amg_par p);
void DDalphaAMG_update_parameters(const struct dd_alpha_amg_parameters *amg_params);
void DDalphaAMG_setup(int iterations, int *status);
double DDalphaAMG_solve(double *vector_out, double *vector_in, double tol, int *status);
void DDalphaAMG_finalize(void);

#ifdef __cplusplus
}
#endif
#endif

/* dd_alpha_amg_b200.h -- operator-level C entry points of libdd_alpha_amg.so (B200-native DDalphaAMG solve path).
 *
 * The reference exposes its hot-path operators only as internal C functions (no FFI); each entry point below is the
 * C-ABI handle on the device implementation of one of them, so that tests can compare operator by operator with the
 * reference and benchmarks can time one kernel family at a time.  Plain pointers and sizes only.
 *
 * Host-side vector conventions (identical to the reference's lexicographic order, data_layout.h:31-33):
 *   fine double vectors : complex double [t][z][y][x][12]           (index 3*spin+colour)
 *   level float vectors : complex float  [lexicographic site of the level][site vars]   (12 on the fine level,
 *                         2*Nv below; first Nv = upper chirality)
 * Everything is converted to the device-native order (aggregate -> Schwarz block -> even/odd, 32-site tiles)
 * inside the library.
 */
#ifndef DD_ALPHA_AMG_B200_OPS_H
#define DD_ALPHA_AMG_B200_OPS_H
#ifdef __cplusplus
extern "C" {
#endif

enum { DDA_DOUBLE = 0, DDA_FLOAT = 1 };
enum { DDA_INFO_NUM_LEVELS = 0, DDA_INFO_SITES = 1, DDA_INFO_SITE_VARS = 2, DDA_INFO_TEST_VECTORS = 3,
       DDA_INFO_BLOCK_SITES = 4, DDA_INFO_NUM_BLOCKS = 5, DDA_INFO_EMULATION = 6,
       DDA_INFO_COARSEST_REPLICATED = 7 /* 1: the coarsest lattice is gathered and solved on every rank */ };
enum { DDA_OPT_USE_FAST = 0, DDA_OPT_PROFILE = 1, DDA_OPT_SEED = 2, DDA_OPT_PRINT = 3 };
enum { DDA_STAT_LAUNCHES = 0, DDA_STAT_DEVICE_BYTES = 1, DDA_STAT_PLAQUETTE = 2, DDA_STAT_ITER = 3,
       DDA_STAT_COARSE_ITER = 4, DDA_STAT_T_COARSEST = 5, DDA_STAT_T_RESTRICT = 6, DDA_STAT_T_INTERPOLATE = 7,
       DDA_STAT_T_SMOOTH0 = 10, DDA_STAT_T_OP0 = 20 };
enum { DDA_OP_APPLY = 0, DDA_OP_RESTRICT = 1, DDA_OP_INTERPOLATE = 2, DDA_OP_SMOOTHER = 3, DDA_OP_VCYCLE = 4,
       DDA_OP_COARSEST_SOLVE = 5 };
enum { DDA_BENCH_DW_DOUBLE = 0, DDA_BENCH_DW_FLOAT = 1, DDA_BENCH_LEVEL_APPLY = 2, DDA_BENCH_RESTRICT = 3,
       DDA_BENCH_INTERPOLATE = 4, DDA_BENCH_SMOOTHER = 5, DDA_BENCH_VCYCLE = 6,
       DDA_BENCH_COARSEST_SCHUR = 7 /* even-odd Schur complement of the coarsest operator (coarse_oddeven_generic.c:1162) */ };

/* geometry of the hierarchy (reference: level_struct fields, main.h:263-341) */
int dda_info(int what, int depth);
int dda_is_emulation(void);   /* 1 only for the g++ host-emulation test build (tests/_emu), 0 for the CUDA library */
void dda_set_option(int what, double value);
double dda_get_stat(int what);
void dda_reset_stats(void);

/* eta = D_W phi on the fine level; reference: d_plus_clover_double / d_plus_clover_float (dirac_generic.c:159-277) */
void dda_apply_dw(int precision, double *out_lex, const double *in_lex);
/* operator arrays in the reference's formats: D[36*site+9*mu+3*r+c] = U/2 (dirac.c:80), clover[42*site+k]
 * (dirac.c:386-398); lexicographic sites, complex double */
void dda_get_operator(double *D_lex, double *clover_lex);

/* prolongator of level `depth`, [lexicographic site][site var][Nv] complex float; reference:
 * interpolation_PRECISION_struct.operator (interpolation_generic.c:74-90).  Setting it rebuilds the Galerkin coarse
 * operators below (coarse_operator_PRECISION_setup, coarse_operator_generic.c:53-205). */
void dda_set_interpolation(int depth, const float *P_lex);
void dda_get_interpolation(int depth, float *P_lex);

/* one hot-path operator of level `depth` on host vectors:
 *  DDA_OP_APPLY          out = D in            apply_coarse_operator_PRECISION (coarse_operator_generic.c:383) / d_plus_clover_float
 *  DDA_OP_RESTRICT       out(depth+1) = P^H in restrict_PRECISION (interpolation_generic.c:169)
 *  DDA_OP_INTERPOLATE    out = P in(depth+1)   interpolate3_PRECISION (interpolation_generic.c:130)
 *  DDA_OP_SMOOTHER       out = SAP(eta = in), iparam iterations, flag != 0: out holds the initial guess
 *                                              smoother_PRECISION -> red_black_schwarz_PRECISION (schwarz_generic.c:1260)
 *  DDA_OP_VCYCLE         out = one cycle applied to in   vcycle_PRECISION (vcycle_generic.c:91)
 *  DDA_OP_COARSEST_SOLVE out = approximate D_c^-1 in     coarse_solve_odd_even_PRECISION (coarse_oddeven_generic.c:1139) */
void dda_level_op(int op, int depth, float *out_lex, const float *in_lex, int iparam, int flag);

/* 12 right-hand sides at once: out[j] = D_c in[j], j < 12, on a coarse level (depth >= 1) through the tensor-core kernel
 * (tcgen05.mma, TF32 x 3 split = fp32 accuracy).  Host arrays [12][sites * site vars] complex float, lexicographic.
 * No reference counterpart; per column it equals DDA_OP_APPLY.  Returns 0, or -1 when the level's
 * shape is not supported.  reps > 0: returns in *ms_out the device time per 12-RHS application over `reps` repetitions. */
int dda_level_apply_mrhs(int depth, float *out_lex, const float *in_lex, int reps, double *ms_out);

/* fine-level test vectors as files "<basename>.NN" in the reference's vector format (vector_io, io.c:704-845: global
 * lexicographic sites, 12 complex doubles each; an optional <header> block is skipped on reading).  The reference reads
 * the same files with "interpolation: 4" + "test vector io file name:" (setup_generic.c:131-162), and so does this
 * library's parameter-file route.  dda_read_test_vectors also rebuilds P and the coarse operators (re_setup). */
void dda_write_test_vectors(const char *basename);
void dda_read_test_vectors(const char *basename);
/* setup policy of the struct route (run_dd_alpha_amg_setup_if_necessary, dd_alpha_amg.c:85-93): full setup when
 * discard_setup_after gauge updates have passed since the last setup, setup update after update_setup_after, then the mass
 * of dd_alpha_amg_update_parameters.  The reference defines this routine but never calls it; it is offered here for
 * HMC callers and does nothing unless called.  Returns 2 (setup), 1 (setup update) or 0. */
int dda_setup_if_necessary(void);

/* device-resident timing (CUDA events on the launching stream), milliseconds per application */
double dda_bench_op(int op, int depth, int reps);
/* device-resident solve: upload once, solve (timed on the device), download */
void dda_upload_source(const double *in_lex);
double dda_solve_device(double tol, int *status, double *ms_out);
void dda_download_solution(double *out_lex);

/* ---- multi-GPU: one process per GPU, process grid = "d0 global lattice" / "d0 local lattice" (T, Z, ... partitions).
 * Replaces the reference's MPI_COMM_WORLD / MPI_Cart_create plumbing (ghost.c:42-66); call before dd_alpha_amg_init.
 * rank 0 creates the NCCL unique id (dda_comm_unique_id, <= 128 bytes), the launcher broadcasts it (MPI_Bcast,
 * torch.distributed, a file ...) and every rank calls dda_comm_init(rank, size, id, cuda device or -1). */
int dda_comm_unique_id(char *out, int len);
void dda_comm_init(int rank, int size, const char *unique_id, int device);
void dda_comm_finalize(void);
int dda_comm_rank(void);
int dda_comm_size(void);
/* host-emulation (test) build only: exchanges are delegated to the test harness (torch.distributed / gloo) */
typedef void (*dda_sendrecv_fn)(const void *send, void *recv, long bytes, int to, int from);
typedef void (*dda_allreduce_fn)(double *buf, int n);
void dda_comm_init_callbacks(int rank, int size, dda_sendrecv_fn sendrecv, dda_allreduce_fn allreduce);

#ifdef __cplusplus
}
#endif
#endif

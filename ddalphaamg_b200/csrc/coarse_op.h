// coarse_op.h -- coarse-grid operator (levels >= 1): storage and kernels.
//
//   (D_c phi)(x) = S(x) phi(x) + sum_mu [ F_mu(x) phi(x+mu) + G5 F_mu(x-mu)^H G5 phi(x-mu) ],   G5 = diag(1_Nv, -1_Nv)
// S(x) and the four forward hops F_mu(x) are dense n x n complex matrices (n = 2 Nv), column-major; the backward hop
// is the gamma5-conjugate transpose of the neighbour's forward hop, exactly the block structure
// [A^H -C^H; -B^H D^H] the reference uses (coarse_operator_generic.h:119-172).  Reference counterparts:
// apply_coarse_operator_PRECISION (coarse_operator_generic.c:383-395), coarse_self_couplings (:288-315),
// coarse_hopping_term / coarse_n_hopping_term (coarse_oddeven_generic.c:447-728), coarse_block_operator
// (coarse_operator_generic.c:208-236), coarse_diag_ee / coarse_diag_oo_inv (coarse_oddeven_generic.c:123-198).
// The hop matrices here carry the operator's sign (F = P^H H P with H = -(1-gamma_mu) D_mu on the fine level).
#pragma once
#include "common.cuh"
#include "lattice.h"
#include "fine_op.h"

namespace dda {

struct CoarseOp {
  int n = 0;                 // complex dofs per site
  long V = 0;
  cf *F = nullptr;           // [site][mu][n*n] column-major forward hops
  cf *S = nullptr;           // [site][n*n] column-major self coupling
  cf *Sinv = nullptr;        // coarsest level: inverse self coupling of the odd sites, [site - n_even][n*n] column-major
  long n_even = 0;
  const int *nb = nullptr;
  const unsigned char *blkflag = nullptr, *aggflag = nullptr;
};

// generic masked apply, same selectors as fine_apply.  self: SELF_C -> S, SELF_CINV -> Sinv (odd sites only).
void coarse_apply(const CoarseOp &op, cf *out, const cf *in, SiteSel sel, int hop, int dir, int self, int outmode,
                  const cf *eta = nullptr, const cf *in_self = nullptr);
// Sinv = S^{-1} on the odd sites of the coarsest level (dense inversion in double, no pivoting; the reference
// factorises LU without pivoting, coarse_oddeven_generic.c:24-73)
void coarse_invert_odd_self(CoarseOp &op);
// optimised full-lattice apply (sm_100a only; coarse_kernel.cu); Z: scratch of 4*n complex per site.  Returns false
// when the shape is not supported (caller falls back to coarse_apply).
bool coarse_apply_fast(const CoarseOp &op, cf *out, const cf *in, cf *Z);
// fused SAP block solve of an intermediate level: for every listed block, biter minimal-residual steps on the block
// operator starting from the block residual r, then x += e (sm_100a only; coarse_kernel.cu).  d_jobs / njobs: the
// level's block-operator job list (Geometry::d_sapjobs).
bool coarse_sap_mr_fast(const CoarseOp &op, cf *x, const cf *r, const int *d_blocklist, int nblk, int bs, int biter,
                        const int *d_jobs, int njobs);

// 12 right-hand sides at once on the tensor cores (sm_100a only; mrhs_kernel.cu): out_j = D_c in_j, vectors j at
// in + j * vstride; Z: scratch of 12 x zstride complex (zstride >= 4 n V); T: the operator's MMA-ready images from
// coarse_mrhs_tile (72 n^2 bytes per site, dev_free() it; rebuild when S / F change).  Ghost slabs of the inputs must be current.
float *coarse_mrhs_tile(const CoarseOp &op);
bool coarse_apply_mrhs(const CoarseOp &op, const float *T, cf *out, const cf *in, cf *Z, long vstride, long zstride);
// even-odd Schur complement of the coarsest operator as streaming kernels (sm_100a only; schur_kernel.cu); vectors are
// full-lattice arrays in global even-odd order, Z: 4*n complex per site, `skip`: device flag (kernels return if set)
bool schur_fast_supported(const CoarseOp &op);
void schur_hop(const CoarseOp &op, int phase, const cf *in, const cf *self, cf *dir, cf *Z, const int *skip);
void schur_mid(const CoarseOp &op, const cf *eta, const cf *dir, const cf *Z, cf *out, float a, float b, float cS, const int *skip);
void schur_fin(const CoarseOp &op, const cf *eta, const cf *dir, const cf *Z, cf *out, float a, float b, const int *skip);
// fused Arnoldi-step kernels of the coarsest-level GMRES (see dev_gmres.h): last Schur stage + inner products with the
// basis; orthogonalisation + norm + Givens step by the last CTA
struct GmresOff;
void schur_fin_dots(const CoarseOp &op, const cf *dir, const cf *Z, cf *w, const cf *V, long stride, int j, double *hb, const int *skip);
void gmres_axpy_givens(cf *w, const cf *V, long stride, int j, long nelem, double *S, int *ct, const GmresOff &o, double tol, unsigned *counter);

}  // namespace dda

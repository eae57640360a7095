// dev_gmres.h -- device-resident restarted GMRES for the coarsest level: the Hessenberg matrix, the Givens rotations,
// the convergence test and the back substitution live on the DEVICE, so an Arnoldi step is a fixed sequence of
// launches without any host synchronisation.  The host only decides how many steps to enqueue: it polls a 4-int control
// block after the number of steps the previous solve needed (then every few steps), and the kernels of steps enqueued
// past convergence return at once (the `skip` flag handed to the operator).
//
// Arithmetic and control flow are those of Fgmres<float> (krylov.h), i.e. the reference's fgmres_PRECISION /
// arnoldi_step / qr_update / compute_solution (linsolve_generic.c:219-413, 809-982): classical Gram-Schmidt with all
// inner products of a step in one reduction, a second reduction for the norm, double-precision scalars.  With several
// ranks the reductions are NCCL all-reduces on the device buffer (every rank sees identical values, hence identical
// control flow).
#pragma once
#include "common.cuh"
#include "blas.h"

namespace dda {

struct DevGmres {
  long n = 0, stride = 0;
  int m = 0, max_restart = 0;
  double tol = 0;
  bool allocated = false;
  cf *V = nullptr;        // m+1 basis vectors, stride `stride` (>= n: room for ghost slabs)
  cf *w = nullptr;
  double *st = nullptr;   // device scalars, see offsets in dev_gmres.cu
  int *ctrl = nullptr;    // device: [0] current restart cycle finished, [1] iterations, [2] valid Arnoldi columns, [3] solve finished
  // out = A in; `skip` (device pointer, may be null): the kernels may return immediately when *skip != 0
  std::function<void(cf *out, const cf *in, const int *skip)> op;
  int last_iter = 0, predicted = 8;
  long polls = 0;
  double last_relres = 0;

  void alloc(long n_, int m_, int max_restart_, double tol_, long nalloc_);
  void release();
  int solve(cf *x, const cf *b);   // zero initial guess; returns the number of iterations
  cf *basis(int k) const { return V + (long)k * stride; }
};

}  // namespace dda

// transfer.h -- aggregation-based interpolation P, restriction P^H, aggregate Gram-Schmidt and the Galerkin
// coarse operator.  Reference counterparts: interpolation_generic.c:93-207 (interpolate, interpolate3, restrict),
// linalg_generic.c:400-454 (gram_schmidt_on_aggregates), coarse_operator_generic.c:53-205 (coarse operator setup).
// P is kept as Nv vectors in the level's own vector layout (aggregates are contiguous site ranges), chirality
// preserving: dofs of the first half of a site feed coarse components [0,Nv), the second half [Nv,2Nv).
#pragma once
#include "common.cuh"
#include "lattice.h"
#include "coarse_op.h"

namespace dda {

const int MAX_NV = 64;

struct Transfer {
  Lay lay;              // layout of the finer level
  long V = 0;           // sites of the finer level
  int nc = 0;           // dofs per site of the finer level
  int nv = 0;           // test vectors -> coarse site has 2*nv dofs
  int nagg = 0, as = 0;
  const int *agg2coarse = nullptr;
  const cf *P[MAX_NV];
};

// phi_c = P^H phi.  Output element (aggregate a, component k) goes to out[agg2coarse[a]*site_stride + offset + k].
void tr_restrict(const Transfer &t, cf *out, long site_stride, long offset, const cf *phi, double *scratch);
// phi = P phi_c (add = false) or phi += P phi_c (add = true)
void tr_interpolate(const Transfer &t, cf *phi, const cf *phi_c, bool add);
// orthonormalise the nv vectors per aggregate and chirality (classical Gram-Schmidt with one re-orthogonalisation)
void tr_gram_schmidt_aggregates(const Transfer &t, cf *const *vecs, double *scratch);
// v = chirality part `ch` of src (other chirality zeroed)
void tr_chirality_part(const Transfer &t, cf *v, const cf *src, int ch);
long tr_scratch_doubles(const Transfer &t);
extern int g_transfer_fast;   // 1: hand-tuned kernels where available (follows Solver::use_fast)

}  // namespace dda

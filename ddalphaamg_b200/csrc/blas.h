// blas.h -- BLAS-1 and reduction kernels on device vectors (layout-agnostic: flat complex arrays).
// Reference counterparts: linalg_generic.c:29-353 (global_inner_product, global_norm, process_multi_inner_product,
// vector_PRECISION_{plus,minus,scale,real_scale,copy,saxpy,multi_saxpy}) and linalg.c:25-111 (mixed precision).
// All reductions accumulate in double, also for float vectors.
#pragma once
#include "common.cuh"

namespace dda {

template <class T> void vzero(cx<T> *x, long n);
template <class T> void vcopy(cx<T> *y, const cx<T> *x, long n);
template <class T> void vscale(cx<T> *y, const cx<T> *x, double a, long n);                 // y = a x (a real)
template <class T> void vaxpy(cx<T> *y, cd a, const cx<T> *x, long n);                      // y += a x
template <class T> void vxpay(cx<T> *z, const cx<T> *x, cd a, const cx<T> *y, long n);      // z = x + a y
template <class T> void vsub(cx<T> *z, const cx<T> *x, const cx<T> *y, long n);             // z = x - y
template <class T> void vadd(cx<T> *z, const cx<T> *x, const cx<T> *y, long n);             // z = x + y
template <class T, class S> void vcast(cx<T> *y, const cx<S> *x, long n);                   // precision cast
// y -= sum_k coef[k] V[k]   (coef on the host; k < m <= 64); reference vector_PRECISION_multi_saxpy
template <class T> void vmulti_axpy(cx<T> *y, cx<T> *const *V, const cd *coef, int m, int sign, long n);
// same update, returns sum |y_new|^2 (fused pass; synchronises)
template <class T> double vmulti_axpy_norm2(cx<T> *y, cx<T> *const *V, const cd *coef, int m, int sign, long n);

// host-returning reductions (synchronise the stream); the sum over ranks (comm_allreduce_sum, NCCL) is applied to the
// device buffer before the single device->host copy
template <class T> cd vdot(const cx<T> *x, const cx<T> *y, long n);     // <x,y> = sum conj(x) y
template <class T> double vnorm2(const cx<T> *x, long n);                 // sum |x|^2
template <class T> void vmulti_dot(cd *out, cx<T> *const *V, int m, const cx<T> *w, long n);   // out[k] = <V[k], w>
// fused: out[k] = <V[k], w> for k<m and out[m] = <w,w> in ONE launch + ONE device->host copy
template <class T> void vmulti_dot_norm(cd *out, cx<T> *const *V, int m, const cx<T> *w, long n);


}  // namespace dda

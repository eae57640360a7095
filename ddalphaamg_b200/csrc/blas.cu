// blas.cu -- BLAS-1 + reductions (see blas.h)
#include "blas.h"
#include "comm.h"

namespace dda {


static const int MAXV = 64;
template <class T> struct PtrArr { const cx<T> *p[MAXV]; };
struct CoefArr { double re[MAXV], im[MAXV]; };

static double *red_buf() {   // device scratch for reduction results
  static double *b = nullptr;
  if (!b) b = dev_alloc<double>(4 * (MAXV + 2));
  return b;
}

template <class T> void vzero(cx<T> *x, long n) { dev_zero(x, sizeof(cx<T>) * n); }
template <class T> void vcopy(cx<T> *y, const cx<T> *x, long n) { if (y != x) d2d(y, x, sizeof(cx<T>) * n); }
template <class T> void vscale(cx<T> *y, const cx<T> *x, double a, long n) {
  T aa = (T)a;
  launch_n(n, DLAMBDA(long i) { y[i] = aa * x[i]; });
}
template <class T> void vaxpy(cx<T> *y, cd a, const cx<T> *x, long n) {
  cx<T> aa((T)a.re, (T)a.im);
  launch_n(n, DLAMBDA(long i) { cx<T> v = y[i]; fma_(v, aa, x[i]); y[i] = v; });
}
template <class T> void vxpay(cx<T> *z, const cx<T> *x, cd a, const cx<T> *y, long n) {
  cx<T> aa((T)a.re, (T)a.im);
  launch_n(n, DLAMBDA(long i) { cx<T> v = x[i]; fma_(v, aa, y[i]); z[i] = v; });
}
template <class T> void vsub(cx<T> *z, const cx<T> *x, const cx<T> *y, long n) { launch_n(n, DLAMBDA(long i) { z[i] = x[i] - y[i]; }); }
template <class T> void vadd(cx<T> *z, const cx<T> *x, const cx<T> *y, long n) { launch_n(n, DLAMBDA(long i) { z[i] = x[i] + y[i]; }); }
template <class T, class S> void vcast(cx<T> *y, const cx<S> *x, long n) { launch_n(n, DLAMBDA(long i) { y[i] = cx<T>((T)x[i].re, (T)x[i].im); }); }

template <class T> static void vmulti_axpy_chunk(cx<T> *y, cx<T> *const *V, const cd *coef, int m, int sign, long n) {
  DDA_ASSERT(m <= MAXV);
  if (m <= 0) return;
  PtrArr<T> pa; CoefArr ca;
  for (int k = 0; k < m; k++) { pa.p[k] = V[k]; ca.re[k] = sign * coef[k].re; ca.im[k] = sign * coef[k].im; }
  launch_n(n, DLAMBDA(long i) {
    cx<T> v = y[i];
    for (int k = 0; k < m; k++) fma_(v, cx<T>((T)ca.re[k], (T)ca.im[k]), pa.p[k][i]);
    y[i] = v;
  });
}

// out[k] = <V[k], w>, k < m.  One pass per chunk of 8 basis vectors: every thread reads w[i] once and the 8 V[k][i],
// so the traffic is (m + ceil(m/8)) vectors instead of 2 m (one segmented reduction per vector re-reads w each time).
// Reference: process_multi_inner_product_PRECISION (linalg_generic.c:107-154).
#ifndef DDA_HOST_EMU
template <class T, int KB> __global__ void __launch_bounds__(256) k_multi_dot(PtrArr<T> pa, int k0, int kn, const cx<T> *__restrict__ w, long n, double *out) {
  double ar[KB], ai[KB];
#pragma unroll
  for (int k = 0; k < KB; k++) { ar[k] = 0.0; ai[k] = 0.0; }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const cx<T> b = w[i];
#pragma unroll
    for (int k = 0; k < KB; k++) if (k < kn) {
      const cx<T> a = pa.p[k0 + k][i];
      ar[k] += (double)a.re * b.re + (double)a.im * b.im;
      ai[k] += (double)a.re * b.im - (double)a.im * b.re;
    }
  }
  __shared__ double sm[2 * KB][8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KB; k++) {
    double x = ar[k], y = ai[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, o); y += __shfl_xor_sync(0xffffffffu, y, o); }
    if (lane == 0) { sm[2 * k][wp] = x; sm[2 * k + 1][wp] = y; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * KB) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) v += sm[threadIdx.x][q];
    const int k = threadIdx.x >> 1;
    if (k < kn) atomicAdd(&out[2 * (k0 + k) + (threadIdx.x & 1)], v);
  }
}
#endif

template <class T> static void vmulti_dot_chunk(cd *out, cx<T> *const *V, int m, const cx<T> *w, long n) {
  DDA_ASSERT(m <= MAXV);
  if (m <= 0) return;
  PtrArr<T> pa;
  for (int k = 0; k < m; k++) pa.p[k] = V[k];
  double *buf = red_buf();
#ifndef DDA_HOST_EMU
  if (n >= (1L << 16)) {
    dev_zero(buf, sizeof(double) * 2 * m);
    const int KB = 8;
    long blocks = std::min<long>((n + 255) / 256, (long)dev_sm_count() * 8);
    for (int k0 = 0; k0 < m; k0 += KB) {
      k_multi_dot<T, KB><<<(unsigned)blocks, 256, 0, g_stream>>>(pa, k0, std::min(KB, m - k0), w, n, buf);
      g_launch_count++;
    }
  } else
#endif
  launch_reduce<2>(m, n, DLAMBDA(long seg, long i, double *acc) {
    cx<T> a = pa.p[seg][i], b = w[i];
    acc[0] += (double)a.re * b.re + (double)a.im * b.im;
    acc[1] += (double)a.re * b.im - (double)a.im * b.re;
  }, buf);
  double h[2 * MAXV];
  comm_allreduce_sum(buf, 2 * m);
  d2h(h, buf, sizeof(double) * 2 * m);
  for (int k = 0; k < m; k++) out[k] = cd(h[2 * k], h[2 * k + 1]);
}

// y += sign * sum_k coef[k] V[k] and ||y_new||^2 in the same pass (the Arnoldi orthogonalisation followed by the norm of
// the new direction, linsolve_generic.c:859-880: one read of y less than saxpy + norm)
template <class T> static double vmulti_axpy_norm2_chunk(cx<T> *y, cx<T> *const *V, const cd *coef, int m, int sign, long n) {
  DDA_ASSERT(m <= MAXV);
  PtrArr<T> pa; CoefArr ca;
  for (int k = 0; k < m; k++) { pa.p[k] = V[k]; ca.re[k] = sign * coef[k].re; ca.im[k] = sign * coef[k].im; }
  double *buf = red_buf();
  launch_reduce<1>(1, n, DLAMBDA(long seg, long i, double *acc) {
    (void)seg;
    cx<T> v = y[i];
    for (int k = 0; k < m; k++) fma_(v, cx<T>((T)ca.re[k], (T)ca.im[k]), pa.p[k][i]);
    y[i] = v;
    acc[0] += (double)v.re * v.re + (double)v.im * v.im;
  }, buf);
  comm_allreduce_sum(buf, 1);
  double h; d2h(&h, buf, sizeof(double));
  return h;
}

// public entry points: any number of basis vectors (the coarsest-level GMRES runs up to `coarse grid iterations`, 100 in
// sample.ini, Arnoldi steps per restart), processed in chunks of MAXV kernel-argument slots
template <class T> void vmulti_axpy(cx<T> *y, cx<T> *const *V, const cd *coef, int m, int sign, long n) {
  for (int k0 = 0; k0 < m; k0 += MAXV) vmulti_axpy_chunk(y, V + k0, coef + k0, std::min(MAXV, m - k0), sign, n);
}
template <class T> void vmulti_dot(cd *out, cx<T> *const *V, int m, const cx<T> *w, long n) {
  for (int k0 = 0; k0 < m; k0 += MAXV) vmulti_dot_chunk(out + k0, V + k0, std::min(MAXV, m - k0), w, n);
}
template <class T> double vmulti_axpy_norm2(cx<T> *y, cx<T> *const *V, const cd *coef, int m, int sign, long n) {
  int k0 = 0;
  for (; m - k0 > MAXV; k0 += MAXV) vmulti_axpy_chunk(y, V + k0, coef + k0, MAXV, sign, n);
  return vmulti_axpy_norm2_chunk(y, V + k0, coef + k0, m - k0, sign, n);
}


template <class T> void vmulti_dot_norm(cd *out, cx<T> *const *V, int m, const cx<T> *w, long n) {
  DDA_ASSERT(m + 1 <= MAXV);
  PtrArr<T> pa;
  for (int k = 0; k < m; k++) pa.p[k] = V[k];
  pa.p[m] = w;
  double *buf = red_buf();
  launch_reduce<2>(m + 1, n, DLAMBDA(long seg, long i, double *acc) {
    cx<T> a = pa.p[seg][i], b = w[i];
    acc[0] += (double)a.re * b.re + (double)a.im * b.im;
    acc[1] += (double)a.re * b.im - (double)a.im * b.re;
  }, buf);
  double h[2 * MAXV];
  comm_allreduce_sum(buf, 2 * (m + 1));
  d2h(h, buf, sizeof(double) * 2 * (m + 1));
  for (int k = 0; k <= m; k++) out[k] = cd(h[2 * k], h[2 * k + 1]);
}

template <class T> cd vdot(const cx<T> *x, const cx<T> *y, long n) {
  cd r; cx<T> *v[1] = {const_cast<cx<T> *>(x)};
  vmulti_dot(&r, v, 1, y, n);
  return r;
}
template <class T> double vnorm2(const cx<T> *x, long n) {
  double *buf = red_buf();
  launch_reduce<1>(1, n, DLAMBDA(long seg, long i, double *acc) { (void)seg; cx<T> a = x[i]; acc[0] += (double)a.re * a.re + (double)a.im * a.im; }, buf);
  comm_allreduce_sum(buf, 1);
  double h; d2h(&h, buf, sizeof(double));
  return h;
}

#define INST(T) \
  template void vzero<T>(cx<T> *, long); \
  template void vcopy<T>(cx<T> *, const cx<T> *, long); \
  template void vscale<T>(cx<T> *, const cx<T> *, double, long); \
  template void vaxpy<T>(cx<T> *, cd, const cx<T> *, long); \
  template void vxpay<T>(cx<T> *, const cx<T> *, cd, const cx<T> *, long); \
  template void vsub<T>(cx<T> *, const cx<T> *, const cx<T> *, long); \
  template void vadd<T>(cx<T> *, const cx<T> *, const cx<T> *, long); \
  template void vmulti_axpy<T>(cx<T> *, cx<T> *const *, const cd *, int, int, long); \
  template double vmulti_axpy_norm2<T>(cx<T> *, cx<T> *const *, const cd *, int, int, long); \
  template cd vdot<T>(const cx<T> *, const cx<T> *, long); \
  template double vnorm2<T>(const cx<T> *, long); \
  template void vmulti_dot<T>(cd *, cx<T> *const *, int, const cx<T> *, long); \
  template void vmulti_dot_norm<T>(cd *, cx<T> *const *, int, const cx<T> *, long);
INST(float)
INST(double)
template void vcast<float, double>(cf *, const cd *, long);
template void vcast<double, float>(cd *, const cf *, long);
template void vcast<float, float>(cf *, const cf *, long);
template void vcast<double, double>(cd *, const cd *, long);

}  // namespace dda

// common.cuh -- core types and launch helpers of the B200 DDalphaAMG solve path.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a (all kernels run on the GPU, no CPU fallback).
// Test-only build: g++ -x c++ -DDDA_HOST_EMU compiles the *same* host logic with kernels executed as host
// loops, so the control flow (cycles, Krylov, setup) can be unit-tested in the GPU-less container.  The
// emulation library is never loaded by the product package (built by ddalphaamg_b200/build.py --emu into tests/_emu).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <functional>
#include <algorithm>

#ifdef DDA_HOST_EMU
  #define __host__
  #define __device__
  #define __global__
  #define __forceinline__ inline
  #define __restrict__
  typedef void *cudaStream_t;
  #define DLAMBDA [=]
#else
  #include <cuda_runtime.h>
  #define DLAMBDA [=] __device__
#endif
#define HD __host__ __device__ __forceinline__

namespace dda {

[[noreturn]] inline void fatal(const char *msg, const char *file, int line) {
  fprintf(stderr, "dd_alpha_amg_b200 fatal: %s (%s:%d)\n", msg, file, line);
  fflush(NULL);
  abort();   // reference convention: errors print and abort (main.h:424-439 error0 -> MPI_Abort)
}
#define DDA_ASSERT(c) do { if (!(c)) ::dda::fatal("assertion failed: " #c, __FILE__, __LINE__); } while (0)

#ifndef DDA_HOST_EMU
#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s: %s\n", #x, cudaGetErrorString(e_)); ::dda::fatal("cuda", __FILE__, __LINE__); } } while (0)
#endif

extern cudaStream_t g_stream;       // compute stream of this process (one process per GPU)
extern long g_launch_count;          // kernels launched (reported by bench.py as gpu_launches)

// ---------------------------------------------------------------------------------------------
// complex numbers
template <class T> struct alignas(2 * sizeof(T)) cx {
  T re, im;
  HD cx() {}
  HD cx(T r, T i = T(0)) : re(r), im(i) {}
  template <class S> HD explicit cx(const cx<S> &o) : re(T(o.re)), im(T(o.im)) {}
};
template <class T> HD cx<T> operator+(cx<T> a, cx<T> b) { return cx<T>(a.re + b.re, a.im + b.im); }
template <class T> HD cx<T> operator-(cx<T> a, cx<T> b) { return cx<T>(a.re - b.re, a.im - b.im); }
template <class T> HD cx<T> operator-(cx<T> a) { return cx<T>(-a.re, -a.im); }
template <class T> HD cx<T> operator*(cx<T> a, cx<T> b) { return cx<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
template <class T> HD cx<T> operator*(T a, cx<T> b) { return cx<T>(a * b.re, a * b.im); }
template <class T> HD cx<T> operator*(cx<T> b, T a) { return cx<T>(a * b.re, a * b.im); }
template <class T> HD cx<T> &operator+=(cx<T> &a, cx<T> b) { a.re += b.re; a.im += b.im; return a; }
template <class T> HD cx<T> &operator-=(cx<T> &a, cx<T> b) { a.re -= b.re; a.im -= b.im; return a; }
template <class T> HD cx<T> conj(cx<T> a) { return cx<T>(a.re, -a.im); }
template <class T> HD T norm2(cx<T> a) { return a.re * a.re + a.im * a.im; }
// fused multiply-add on reals: fma() on the device (one FFMA/DFMA; without it `a += b*c - d*e` compiles to
// FMUL + FFMA + FADD because the compiler may not re-associate), plain expression in the host-emulation build
#if defined(__CUDA_ARCH__) && !defined(DDA_NO_FUSE)
HD float fma_r(float a, float b, float c) { return __fmaf_rn(a, b, c); }
HD double fma_r(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
HD float fma_r(float a, float b, float c) { return a * b + c; }
HD double fma_r(double a, double b, double c) { return a * b + c; }
#endif
// a += b*c ; a += conj(b)*c ; a -= b*c ; a -= conj(b)*c      (4 fused multiply-adds each)
#ifdef DDA_NO_FUSE   // measurement build only (build.py --nofuse): the round-1 expressions, FMUL + FFMA + FADD per component
template <class T> HD void fma_(cx<T> &a, cx<T> b, cx<T> c) { a.re += b.re * c.re - b.im * c.im; a.im += b.re * c.im + b.im * c.re; }
template <class T> HD void fmac_(cx<T> &a, cx<T> b, cx<T> c) { a.re += b.re * c.re + b.im * c.im; a.im += b.re * c.im - b.im * c.re; }
template <class T> HD void fms_(cx<T> &a, cx<T> b, cx<T> c) { a.re -= b.re * c.re - b.im * c.im; a.im -= b.re * c.im + b.im * c.re; }
template <class T> HD void fmsc_(cx<T> &a, cx<T> b, cx<T> c) { a.re -= b.re * c.re + b.im * c.im; a.im -= b.re * c.im - b.im * c.re; }
#else
template <class T> HD void fma_(cx<T> &a, cx<T> b, cx<T> c) { a.re = fma_r(-b.im, c.im, fma_r(b.re, c.re, a.re)); a.im = fma_r(b.im, c.re, fma_r(b.re, c.im, a.im)); }
template <class T> HD void fmac_(cx<T> &a, cx<T> b, cx<T> c) { a.re = fma_r(b.im, c.im, fma_r(b.re, c.re, a.re)); a.im = fma_r(-b.im, c.re, fma_r(b.re, c.im, a.im)); }
template <class T> HD void fms_(cx<T> &a, cx<T> b, cx<T> c) { a.re = fma_r(b.im, c.im, fma_r(-b.re, c.re, a.re)); a.im = fma_r(-b.im, c.re, fma_r(-b.re, c.im, a.im)); }
template <class T> HD void fmsc_(cx<T> &a, cx<T> b, cx<T> c) { a.re = fma_r(-b.im, c.im, fma_r(-b.re, c.re, a.re)); a.im = fma_r(b.im, c.re, fma_r(-b.re, c.im, a.im)); }
#endif
// multiply by a unit: code 0:+1 1:-1 2:+i 3:-i
template <int CODE, class T> HD cx<T> mul_unit(cx<T> z) {
  if (CODE == 0) return z;
  if (CODE == 1) return cx<T>(-z.re, -z.im);
  if (CODE == 2) return cx<T>(-z.im, z.re);
  return cx<T>(z.im, -z.re);
}
typedef cx<float> cf;
typedef cx<double> cd;

// ---------------------------------------------------------------------------------------------
// vector layout: site s, component c of nc  ->  (((s>>sh)*nc + c) << sh) + (s & mask)
// fine level: sh = 5 (tiles of 32 sites, component-major inside a tile => coalesced per-site threads)
// coarse levels: sh = 0 (site-major)
struct Lay {
  int nc, sh;
  HD long idx(long s, int c) const { return ((((s >> sh) * nc + c)) << sh) + (s & ((1L << sh) - 1)); }
  // inverse of idx: flat element q -> (site, component); consecutive q are consecutive in memory
  HD void decode(long q, long &s, int &c) const {
    long t = q >> sh; long tile = t / nc; c = (int)(t - tile * nc); s = (tile << sh) + (q & ((1L << sh) - 1));
  }
};

// ---------------------------------------------------------------------------------------------
// device memory
void *dev_alloc_bytes(size_t bytes);
void dev_free(void *p);
void dev_zero(void *p, size_t bytes);
void h2d(void *dst, const void *src, size_t bytes);
void d2h(void *dst, const void *src, size_t bytes);
void d2d(void *dst, const void *src, size_t bytes);
void dev_sync();
size_t dev_bytes_in_use();
int dev_sm_count();            // multiprocessors of the current device (queried once; 1 in the emulation build)
template <class T> T *dev_alloc(size_t n) { return (T *)dev_alloc_bytes(n * sizeof(T)); }
template <class T> T *dev_upload(const std::vector<T> &v) { T *p = dev_alloc<T>(v.size() ? v.size() : 1); if (v.size()) h2d(p, v.data(), v.size() * sizeof(T)); return p; }

// ---------------------------------------------------------------------------------------------
// launch helpers
#ifndef DDA_HOST_EMU
template <class F> __global__ void k_launch_n(long n, F f) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) f(i);
}
template <class F> void launch_n(long n, F f, int block = 256) {
  if (n <= 0) return;
  k_launch_n<<<(unsigned)((n + block - 1) / block), block, 0, g_stream>>>(n, f);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

// segmented reduction: out[seg*NV + k] += sum_i f(seg,i)[k]; out must be zeroed by the caller (reduce_zero).
// one CTA row per segment, `chunks` CTAs per segment, warp shuffles + one atomicAdd(double) per CTA and value.
template <int NV, class F> __global__ void k_reduce(long seglen, F f, double *out) {
  long seg = blockIdx.x;
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) acc[k] = 0.0;
  for (long i = blockIdx.y * (long)blockDim.x + threadIdx.x; i < seglen; i += (long)gridDim.y * blockDim.x) f(seg, i, acc);
  __shared__ double sm[NV][32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[k][w] = v;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double v = lane < nw ? sm[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(&out[seg * NV + k], v);
    }
  }
}
template <int NV, class F> void launch_reduce(long nseg, long seglen, F f, double *out) {
  if (nseg <= 0) return;
  dev_zero(out, sizeof(double) * NV * nseg);
  if (seglen <= 0) return;
  int block = seglen >= 256 ? 256 : (seglen >= 128 ? 128 : 64);
  long want = ((long)dev_sm_count() * 8 + nseg - 1) / nseg;   // aim for >= 8 CTAs per SM overall
  long maxc = (seglen + block * 4L - 1) / (block * 4L);   // at least 4 elements per thread
  long chunks = std::max(1L, std::min(want, std::min(maxc, 65535L)));
  dim3 grid((unsigned)nseg, (unsigned)chunks);
  k_reduce<NV><<<grid, block, 0, g_stream>>>(seglen, f, out);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}
#else
template <class F> void launch_n(long n, F f, int block = 256) {
  (void)block;
  g_launch_count++;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) f(i);
}
template <int NV, class F> void launch_reduce(long nseg, long seglen, F f, double *out) {
  g_launch_count++;
  if (nseg >= 8) {
#pragma omp parallel for schedule(static)
    for (long seg = 0; seg < nseg; seg++) {
      double acc[NV];
      for (int k = 0; k < NV; k++) acc[k] = 0.0;
      for (long i = 0; i < seglen; i++) f(seg, i, acc);
      for (int k = 0; k < NV; k++) out[seg * NV + k] = acc[k];
    }
  } else {
    for (long seg = 0; seg < nseg; seg++) {
      double tot[NV];
      for (int k = 0; k < NV; k++) tot[k] = 0.0;
#pragma omp parallel
      {
        double acc[NV];
        for (int k = 0; k < NV; k++) acc[k] = 0.0;
#pragma omp for schedule(static) nowait
        for (long i = 0; i < seglen; i++) f(seg, i, acc);
#pragma omp critical
        for (int k = 0; k < NV; k++) tot[k] += acc[k];
      }
      for (int k = 0; k < NV; k++) out[seg * NV + k] = tot[k];
    }
  }
}
#endif

}  // namespace dda

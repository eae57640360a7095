// transfer.cu -- restriction / interpolation / aggregate Gram-Schmidt (generic kernels); see transfer.h.
#include "transfer.h"

namespace dda {

struct VecPtrs { cf *p[MAX_NV]; };

long tr_scratch_doubles(const Transfer &t) { return (long)t.nagg * 2 * t.nv * 2 + 16; }

// phi_c(a)[ch*nv + k] = sum over the sites of aggregate a and the dofs of chirality ch of conj(P_k) * phi
// reference: restrict_PRECISION, interpolation_generic.c:169-207
bool tr_restrict_fast(const Transfer &t, cf *out, long site_stride, long offset, const cf *phi);
int g_transfer_fast = 1;

void tr_restrict(const Transfer &t, cf *out, long site_stride, long offset, const cf *phi, double *scratch) {
#ifndef DDA_HOST_EMU
  if (g_transfer_fast && tr_restrict_fast(t, out, site_stride, offset, phi)) return;
#endif
  const int nv = t.nv, h = t.nc / 2, as = t.as;
  const long nseg = (long)t.nagg * 2 * nv, seglen = (long)as * h;
  const Transfer tt = t;
  launch_reduce<2>(nseg, seglen, DLAMBDA(long seg, long i, double *acc) {
    long a = seg / (2 * nv); int j = (int)(seg - a * 2 * nv); int ch = j / nv, k = j - ch * nv;
    long sl = i / h; int c = ch * h + (int)(i - sl * h);
    long q = tt.lay.idx(a * as + sl, c);
    cf p = tt.P[k][q], v = phi[q];
    acc[0] += (double)p.re * v.re + (double)p.im * v.im;
    acc[1] += (double)p.re * v.im - (double)p.im * v.re;
  }, scratch);
  const int *a2c = t.agg2coarse;
  launch_n(nseg, DLAMBDA(long seg) {
    long a = seg / (2 * nv); int j = (int)(seg - a * 2 * nv);
    out[(long)a2c[a] * site_stride + offset + j] = cf((float)scratch[2 * seg], (float)scratch[2 * seg + 1]);
  });
}

// phi (+)= P phi_c ; reference: interpolate_PRECISION / interpolate3_PRECISION, interpolation_generic.c:93-166
void tr_interpolate(const Transfer &t, cf *phi, const cf *phi_c, bool add) {
  const int nv = t.nv, h = t.nc / 2, as = t.as;
  const Transfer tt = t;
  const int *a2c = t.agg2coarse;
  launch_n(t.V * t.nc, DLAMBDA(long q) {
    long s; int c; tt.lay.decode(q, s, c);
    long a = s / as; int ch = c / h;
    const cf *pc = phi_c + (long)a2c[a] * 2 * nv + ch * nv;
    cf acc = add ? phi[q] : cf(0.f, 0.f);
    for (int k = 0; k < nv; k++) fma_(acc, tt.P[k][q], pc[k]);
    phi[q] = acc;
  });
}

void tr_chirality_part(const Transfer &t, cf *v, const cf *src, int ch) {
  const Lay lay = t.lay; const int h = t.nc / 2;
  launch_n(t.V * t.nc, DLAMBDA(long q) {
    long s; int c; lay.decode(q, s, c);
    v[q] = (c / h == ch) ? src[q] : cf(0.f, 0.f);
  });
}

// Orthonormalise vecs[0..nv) per aggregate and per chirality.  Classical Gram-Schmidt applied twice per vector
// (numerically equivalent to the reference's modified Gram-Schmidt, linalg_generic.c:400-454); all coefficients stay
// on the device, no host synchronisation.
bool tr_gram_schmidt_fast(const Transfer &t, cf *const *vecs);   // transfer_kernel.cu: CholeskyQR2, one CTA per (aggregate, chirality)

void tr_gram_schmidt_aggregates(const Transfer &t, cf *const *vecs, double *scratch) {
#ifndef DDA_HOST_EMU
  {
    const char *e = getenv("DDA_GS_FAST");                         // A/B switch, read per call (tests compare both paths)
    if (g_transfer_fast && !(e && atoi(e) == 0) && tr_gram_schmidt_fast(t, vecs)) return;
  }
#endif
  const int nv = t.nv, h = t.nc / 2, as = t.as;
  const Lay lay = t.lay;
  VecPtrs vp;
  for (int k = 0; k < nv; k++) vp.p[k] = vecs[k];
  const long nagg = t.nagg, seglen = (long)as * h;
  for (int k1 = 0; k1 < nv; k1++) {
    for (int pass = 0; pass < 2 && k1 > 0; pass++) {
      // coef[(a*2+ch)*k1 + k2] = <v_k2, v_k1> on (a,ch)
      launch_reduce<2>(nagg * 2 * k1, seglen, DLAMBDA(long seg, long i, double *acc) {
        long ac = seg / k1; int k2 = (int)(seg - ac * k1); long a = ac >> 1; int ch = (int)(ac & 1);
        long sl = i / h; int c = ch * h + (int)(i - sl * h);
        long q = lay.idx(a * as + sl, c);
        cf p = vp.p[k2][q], v = vp.p[k1][q];
        acc[0] += (double)p.re * v.re + (double)p.im * v.im;
        acc[1] += (double)p.re * v.im - (double)p.im * v.re;
      }, scratch);
      launch_n(t.V * t.nc, DLAMBDA(long q) {
        long s; int c; lay.decode(q, s, c);
        long a = s / as; int ch = c / h;
        const double *co = scratch + 2 * ((a * 2 + ch) * k1);
        cf v = vp.p[k1][q];
        for (int k2 = 0; k2 < k1; k2++) fms_(v, cf((float)co[2 * k2], (float)co[2 * k2 + 1]), vp.p[k2][q]);
        vp.p[k1][q] = v;
      });
    }
    launch_reduce<1>(nagg * 2, seglen, DLAMBDA(long seg, long i, double *acc) {
      long a = seg >> 1; int ch = (int)(seg & 1);
      long sl = i / h; int c = ch * h + (int)(i - sl * h);
      cf v = vp.p[k1][lay.idx(a * as + sl, c)];
      acc[0] += (double)v.re * v.re + (double)v.im * v.im;
    }, scratch);
    launch_n(t.V * t.nc, DLAMBDA(long q) {
      long s; int c; lay.decode(q, s, c);
      long a = s / as; int ch = c / h;
      double nn = scratch[a * 2 + ch];
      float f = nn > 0.0 ? (float)(1.0 / sqrt(nn)) : 0.f;
      vp.p[k1][q] = f * vp.p[k1][q];
    });
  }
}

}  // namespace dda

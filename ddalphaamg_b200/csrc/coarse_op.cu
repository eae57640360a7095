// coarse_op.cu -- generic coarse operator kernels (one thread per (site,row)); see coarse_op.h.
#include "coarse_op.h"

namespace dda {

void coarse_apply(const CoarseOp &op, cf *out, const cf *in, SiteSel sel, int hop, int dir, int self, int outmode,
                  const cf *eta, const cf *in_self) {
  const int n = op.n, nh = op.n / 2;
  if (!in_self) in_self = in;
  const long nn = (long)n * n;
  launch_n(sel.n * n, DLAMBDA(long i) {
    long q = i / n; int r = (int)(i - q * n);
    long s = sel_site(sel, q);
    unsigned mask = 0;
    if (hop == HOP_ALL) mask = 0xFFu;
    else if (hop == HOP_INBLOCK) mask = (~(unsigned)op.blkflag[s]) & 0xFFu;
    else if (hop == HOP_INAGG) mask = (~(unsigned)op.aggflag[s]) & 0xFFu;
    else if (hop == HOP_CROSSAGG) mask = ((unsigned)op.aggflag[s]) & (1u << dir);
    else if (hop == HOP_CROSSBLOCK) mask = ((unsigned)op.blkflag[s]) & 0xFFu;
    cf acc(0.f, 0.f);
    if (self == SELF_C) {
      const cf *M = op.S + s * nn, *v = in_self + s * n;
      for (int c = 0; c < n; c++) fma_(acc, M[(long)c * n + r], v[c]);
    } else if (self == SELF_CINV) {
      const cf *M = op.Sinv + (s - op.n_even) * nn, *v = in_self + s * n;
      for (int c = 0; c < n; c++) fma_(acc, M[(long)c * n + r], v[c]);
    }
    for (int mu = 0; mu < 4; mu++) {
      if (mask & (1u << mu)) {
        long nbr = op.nb[(long)mu * op.V + s];
        const cf *M = op.F + (s * 4 + mu) * nn, *v = in + nbr * n;
        for (int c = 0; c < n; c++) fma_(acc, M[(long)c * n + r], v[c]);
      }
      if (mask & (1u << (4 + mu))) {
        long nbr = op.nb[(long)(4 + mu) * op.V + s];
        const cf *M = op.F + (nbr * 4 + mu) * nn + (long)r * n, *v = in + nbr * n;   // row r of F^H = conj of column r of F
        cf a1(0.f, 0.f), a2(0.f, 0.f);
        for (int c = 0; c < nh; c++) fmac_(a1, M[c], v[c]);
        for (int c = nh; c < n; c++) fmac_(a2, M[c], v[c]);
        acc += (r < nh) ? (a1 - a2) : (a2 - a1);
      }
    }
    long k = s * n + r;
    if (outmode == OUT_SET) out[k] = acc;
    else if (outmode == OUT_ADD) out[k] += acc;
    else if (outmode == OUT_SUB) out[k] -= acc;
    else if (outmode == OUT_NEG) out[k] = -acc;
    else out[k] = eta[k] - acc;
  }, 128);
}

bool coarse_invert_odd_self_fast(CoarseOp &op);   // coarse_kernel.cu (sm_100a): one CTA per site, block in shared memory

void coarse_invert_odd_self(CoarseOp &op) {
  const int n = op.n; const long nn = (long)n * n;
  long nodd = op.V - op.n_even;
  if (nodd <= 0) return;
#ifndef DDA_HOST_EMU
  if (coarse_invert_odd_self_fast(op)) return;
#endif
  cd *scratch = dev_alloc<cd>(2 * nn * nodd);
  const cf *S = op.S; cf *Sinv = op.Sinv; long ne = op.n_even;
  launch_n(nodd, DLAMBDA(long o) {
    cd *A = scratch + 2 * nn * o, *R = A + nn;      // row-major work copies
    const cf *M = S + (ne + o) * nn;
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) { cf v = M[(long)c * n + r]; A[(long)r * n + c] = cd(v.re, v.im); R[(long)r * n + c] = cd(r == c ? 1.0 : 0.0, 0.0); }
    for (int p = 0; p < n; p++) {
      cd piv = A[(long)p * n + p]; double d = 1.0 / norm2(piv); cd ip(piv.re * d, -piv.im * d);
      for (int c = 0; c < n; c++) { A[(long)p * n + c] = A[(long)p * n + c] * ip; R[(long)p * n + c] = R[(long)p * n + c] * ip; }
      for (int r = 0; r < n; r++) if (r != p) {
        cd f = A[(long)r * n + p];
        if (f.re == 0.0 && f.im == 0.0) continue;
        for (int c = 0; c < n; c++) { fms_(A[(long)r * n + c], f, A[(long)p * n + c]); fms_(R[(long)r * n + c], f, R[(long)p * n + c]); }
      }
    }
    cf *O = Sinv + o * nn;
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) { cd v = R[(long)r * n + c]; O[(long)c * n + r] = cf((float)v.re, (float)v.im); }
  }, 32);
  dev_sync();
  dev_free(scratch);
}

}  // namespace dda

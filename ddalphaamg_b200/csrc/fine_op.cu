// fine_op.cu -- fine-level Wilson-Clover kernels: generic masked apply, clover/plaquette construction,
// precision casts, clover-block inversion for the SAP Schur complement.
#include "fine_op.h"
#include "fine_op.cuh"
#include "comm.h"

namespace dda {

// ---------------------------------------------------------------------------------------------------
// generic (masked) operator apply: res = [self term] + [selected hops], then combine into `out`.
// Reference counterparts: d_plus_clover (dirac_generic.c:159), block_d_plus_clover (:83), block_diag_ee/oo_inv and
// block_(n_)hopping_term (oddeven_generic.c:975-1314), aggregate self/neighbour couplings (dirac_generic.c:308-462).
template <class T> void fine_apply(const FineOp<T> &op, cx<T> *out, const cx<T> *in, SiteSel sel, int hop, int dir,
                                   int self, int outmode, const cx<T> *eta, const cx<T> *in_self) {
  const Lay ls = {12, op.sh};
  if (!in_self) in_self = in;
  launch_n(sel.n, DLAMBDA(long i) {
    long s = sel_site(sel, i);
    unsigned mask = 0;
    if (hop == HOP_ALL) mask = 0xFFu;
    else if (hop == HOP_INBLOCK) mask = (~(unsigned)op.blkflag[s]) & 0xFFu;
    else if (hop == HOP_INAGG) mask = (~(unsigned)op.aggflag[s]) & 0xFFu;
    else if (hop == HOP_CROSSAGG) mask = ((unsigned)op.aggflag[s]) & (1u << dir);
    else if (hop == HOP_CROSSBLOCK) mask = ((unsigned)op.blkflag[s]) & 0xFFu;
    cx<T> r[12];
    if (self != SELF_NONE) {
      cx<T> x[12];
#pragma unroll
      for (int c = 0; c < 12; c++) x[c] = in_self[ls.idx(s, c)];
      clover_mul(self == SELF_C ? op.C : op.Cinv, op.sh, s, x, r);
    } else {
#pragma unroll
      for (int c = 0; c < 12; c++) r[c] = cx<T>(T(0), T(0));
    }
    if (mask) {
      hop_pair<0>(op, s, mask, in, r);
      hop_pair<1>(op, s, mask, in, r);
      hop_pair<2>(op, s, mask, in, r);
      hop_pair<3>(op, s, mask, in, r);
    }
#pragma unroll
    for (int c = 0; c < 12; c++) {
      long k = ls.idx(s, c);
      if (outmode == OUT_SET) out[k] = r[c];
      else if (outmode == OUT_ADD) out[k] += r[c];
      else if (outmode == OUT_SUB) out[k] -= r[c];
      else if (outmode == OUT_NEG) out[k] = -r[c];
      else out[k] = eta[k] - r[c];
    }
  }, 128);
}
template void fine_apply<float>(const FineOp<float> &, cf *, const cf *, SiteSel, int, int, int, int, const cf *, const cf *);
template void fine_apply<double>(const FineOp<double> &, cd *, const cd *, SiteSel, int, int, int, int, const cd *, const cd *);

// ---------------------------------------------------------------------------------------------------
// clover term + plaquette from the links (double).  Restates dirac.c:24-58 (compute_clover_term), :304-402
// (Q, Qdiff, set_clover) and :568-622 (calc_plaq):
//   C(x) = (4+m0) - csw * sum_{mu<nu} (gamma_mu gamma_nu) (x) (Q_munu - Q_munu^dagger),
//   Q_munu = (sum of the four plaquette leaves in the mu-nu plane touching x)/16 computed from U (here from D=U/2).
struct M3 { cd a[9]; };
static HD M3 m3_mul(const M3 &A, const M3 &B) { M3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cd s(0.0, 0.0); for (int k = 0; k < 3; k++) fma_(s, A.a[3 * i + k], B.a[3 * k + j]); C.a[3 * i + j] = s; } return C; }
static HD M3 m3_mul_bd(const M3 &A, const M3 &B) { M3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cd s(0.0, 0.0); for (int k = 0; k < 3; k++) fma_(s, A.a[3 * i + k], conj(B.a[3 * j + k])); C.a[3 * i + j] = s; } return C; }   // A B^dag
static HD M3 m3_mul_ad(const M3 &A, const M3 &B) { M3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cd s(0.0, 0.0); for (int k = 0; k < 3; k++) fmac_(s, A.a[3 * k + i], B.a[3 * k + j]); C.a[3 * i + j] = s; } return C; }  // A^dag B
static HD M3 m3_mul_adbd(const M3 &A, const M3 &B) { M3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cd s(0.0, 0.0); for (int k = 0; k < 3; k++) fma_(s, conj(A.a[3 * k + i]), conj(B.a[3 * j + k])); C.a[3 * i + j] = s; } return C; }  // A^dag B^dag

// neighbour-table walk from site s: shift s1 along d1, then s0 along d0 (d < 0: no shift).  Intermediate sites may be
// ghost sites of a partitioned direction: they have their own neighbour table (Geometry::d_nbg), and the ghost slabs are
// built so that the corner sites exist (lattice.cu).
struct CloverGeom { const int *nb, *nbg; long V, Vg; int sh; };
static HD long cg_step(const CloverGeom &g, long s, int d) {
  return s < g.V ? g.nb[(long)d * g.V + s] : g.nbg[(long)d * g.Vg + (s - g.V)];
}
static HD long cg_site(const CloverGeom &g, long s, int d0, int s0, int d1, int s1) {
  if (d1 >= 0) s = cg_step(g, s, s1 > 0 ? d1 : 4 + d1);
  if (d0 >= 0) s = cg_step(g, s, s0 > 0 ? d0 : 4 + d0);
  return s;
}
static HD M3 cg_link(const CloverGeom &g, const cd *D, long s, int mu) {
  const Lay lu = {36, g.sh};
  M3 m; for (int k = 0; k < 9; k++) m.a[k] = D[lu.idx(s, 9 * mu + k)];
  return m;
}

struct GammaTab { int co[4][4]; double vr[4][4], vi[4][4]; };
static GammaTab gamma_tab() {
  GammaTab g;
  const int co[4][4] = {{2, 3, 0, 1}, {3, 2, 1, 0}, {3, 2, 1, 0}, {2, 3, 0, 1}};
  const int code[4][4] = {{1, 1, 1, 1}, {3, 3, 2, 2}, {1, 0, 0, 1}, {3, 2, 2, 3}};
  const double re[4] = {1, -1, 0, 0}, im[4] = {0, 0, 1, -1};
  for (int m = 0; m < 4; m++) for (int r = 0; r < 4; r++) { g.co[m][r] = co[m][r]; g.vr[m][r] = re[code[m][r]]; g.vi[m][r] = im[code[m][r]]; }
  return g;
}

void fine_build_clover(const Geometry &geo, const cd *D, double *C, double m0, double csw, double *plaq_out) {
  CloverGeom cg; cg.nb = geo.d_nb; cg.nbg = geo.d_nbg; cg.V = geo.V; cg.Vg = geo.Vg; cg.sh = geo.sh;
  const GammaTab gt = gamma_tab();
  const Lay lc = {72, geo.sh};
  double *d_plaq = dev_alloc<double>(1);
  long V = geo.V;
  // plaquette: sum_x sum_{mu<nu} Re tr U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag / (6 V)   [0,3]
  launch_reduce<1>(1, V, DLAMBDA(long seg, long s, double *acc) {
    (void)seg;
    double tr = 0;
    for (int mu = 0; mu < 4; mu++) for (int nu = mu + 1; nu < 4; nu++) {
      M3 a = m3_mul(cg_link(cg, D, s, mu), cg_link(cg, D, cg_site(cg, s, mu, 1, -1, 0), nu));
      M3 b = m3_mul_bd(a, cg_link(cg, D, cg_site(cg, s, nu, 1, -1, 0), mu));
      M3 p = m3_mul_bd(b, cg_link(cg, D, s, nu));
      tr += p.a[0].re + p.a[4].re + p.a[8].re;
    }
    acc[0] += 16.0 * tr;     // D = U/2
  }, d_plaq);
  comm_allreduce_sum(d_plaq, 1);
  double h; d2h(&h, d_plaq, sizeof(double)); dev_free(d_plaq);
  if (plaq_out) *plaq_out = h / (6.0 * (double)V * (double)g_comm.size);

  launch_n(V, DLAMBDA(long s) {
    cd blk[2][36];
    for (int b = 0; b < 2; b++) for (int k = 0; k < 36; k++) blk[b][k] = cd(0.0, 0.0);
    if (csw != 0.0) for (int mu = 0; mu < 4; mu++) for (int nu = mu + 1; nu < 4; nu++) {
      long xpm = cg_site(cg, s, mu, 1, -1, 0), xpn = cg_site(cg, s, nu, 1, -1, 0);
      long xmm = cg_site(cg, s, mu, -1, -1, 0), xmn = cg_site(cg, s, nu, -1, -1, 0);
      long xpn_mm = cg_site(cg, s, mu, -1, nu, 1), xmm_mn = cg_site(cg, s, mu, -1, nu, -1), xmn_pm = cg_site(cg, s, mu, 1, nu, -1);
      // leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag
      M3 q = m3_mul_bd(m3_mul_bd(m3_mul(cg_link(cg, D, s, mu), cg_link(cg, D, xpm, nu)), cg_link(cg, D, xpn, mu)), cg_link(cg, D, s, nu));
      // leaf 2: U_nu(x) U_mu(x+nu-mu)^dag U_nu(x-mu)^dag U_mu(x-mu)
      M3 l2 = m3_mul(m3_mul_bd(m3_mul_bd(cg_link(cg, D, s, nu), cg_link(cg, D, xpn_mm, mu)), cg_link(cg, D, xmm, nu)), cg_link(cg, D, xmm, mu));
      // leaf 3: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu)
      M3 l3 = m3_mul(m3_mul(m3_mul_adbd(cg_link(cg, D, xmm, mu), cg_link(cg, D, xmm_mn, nu)), cg_link(cg, D, xmm_mn, mu)), cg_link(cg, D, xmn, nu));
      // leaf 4: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu) U_mu(x)^dag
      M3 l4 = m3_mul_bd(m3_mul(m3_mul_ad(cg_link(cg, D, xmn, nu), cg_link(cg, D, xmn, mu)), cg_link(cg, D, xmn_pm, nu)), cg_link(cg, D, s, mu));
      M3 Qd;
      for (int k = 0; k < 9; k++) q.a[k] = q.a[k] + l2.a[k] + l3.a[k] + l4.a[k];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Qd.a[3 * i + j] = q.a[3 * i + j] - conj(q.a[3 * j + i]);
      // -csw * (gamma_mu gamma_nu) (x) Qd, only the two diagonal 2x2 spin blocks are populated
      for (int r = 0; r < 4; r++) {
        int a = gt.co[mu][r], cc = gt.co[nu][a];
        cd v = cd(gt.vr[mu][r], gt.vi[mu][r]) * cd(gt.vr[nu][a], gt.vi[nu][a]);
        int b = r >> 1;
        if ((cc >> 1) != b) continue;
        cd f = cd(-csw, 0.0) * v;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
          fma_(blk[b][6 * (3 * (r & 1) + i) + 3 * (cc & 1) + j], f, Qd.a[3 * i + j]);
      }
    }
    for (int b = 0; b < 2; b++) {
      for (int i = 0; i < 6; i++) C[lc.idx(s, 6 * b + i)] = 4.0 + m0 + blk[b][7 * i].re;
      int m = 0;
      for (int i = 0; i < 6; i++) for (int j = i + 1; j < 6; j++, m++) {
        C[lc.idx(s, 12 + 2 * (15 * b + m))] = blk[b][6 * i + j].re;
        C[lc.idx(s, 12 + 2 * (15 * b + m) + 1)] = blk[b][6 * i + j].im;
      }
    }
  }, 64);
}

// add a real shift to the clover diagonal (mass update; reference shift_update_PRECISION, dirac_generic.c:504-551)
void fine_shift_clover(const Geometry &geo, double *C, double delta) {
  const Lay lc = {72, geo.sh};
  launch_n(geo.V * 12, DLAMBDA(long i) { long s = i / 12; int k = (int)(i - 12 * s); C[lc.idx(s, k)] += delta; });
}

// even/odd scaling of the clover term (reference scale_clover, dirac.c:646-667); parity from lexicographic coords
void fine_scale_clover(const Geometry &geo, double *C, double se, double so) {
  const Lay lc = {72, geo.sh};
  const int *n2l = geo.d_nat2lex; int L1 = geo.L[1], L2 = geo.L[2], L3 = geo.L[3];
  const int poff = (geo.pc[0] * geo.L[0] + geo.pc[1] * geo.L[1] + geo.pc[2] * geo.L[2] + geo.pc[3] * geo.L[3]) & 1;
  launch_n(geo.V * 72, DLAMBDA(long i) {
    long s = i / 72; int k = (int)(i - 72 * s);
    long lx = n2l[s]; int x = (int)(lx % L3); lx /= L3; int y = (int)(lx % L2); lx /= L2; int z = (int)(lx % L1); int t = (int)(lx / L1);
    C[lc.idx(s, k)] *= ((t + z + y + x + poff) & 1) ? so : se;
  });
}

// precision casts of operator arrays
void cast_links(const cd *src, cf *dst, long n) { launch_n(n, DLAMBDA(long i) { dst[i] = cf((float)src[i].re, (float)src[i].im); }); }
void cast_reals(const double *src, float *dst, long n) { launch_n(n, DLAMBDA(long i) { dst[i] = (float)src[i]; }); }

// Cinv(s) = inverse of the two Hermitian 6x6 clover blocks, same packing (computed in double by Gauss-Jordan on the
// Hermitian positive definite blocks).  Replaces the reference's per-site Cholesky factor + forward/backward
// substitution (oddeven_generic.c:24-114): one multiply with the explicit inverse is the GPU-friendly form.
void fine_invert_clover(const Geometry &geo, const double *C, double *Cinv) {
  const Lay lc = {72, geo.sh};
  launch_n(geo.V * 2, DLAMBDA(long i) {
    long s = i >> 1; int b = (int)(i & 1);
    cd A[36], R[36];
    for (int k = 0; k < 36; k++) { A[k] = cd(0.0, 0.0); R[k] = cd(0.0, 0.0); }
    for (int k = 0; k < 6; k++) { A[7 * k] = cd(C[lc.idx(s, 6 * b + k)], 0.0); R[7 * k] = cd(1.0, 0.0); }
    int m = 0;
    for (int r = 0; r < 6; r++) for (int c = r + 1; c < 6; c++, m++) {
      cd v(C[lc.idx(s, 12 + 2 * (15 * b + m))], C[lc.idx(s, 12 + 2 * (15 * b + m) + 1)]);
      A[6 * r + c] = v; A[6 * c + r] = conj(v);
    }
    for (int p = 0; p < 6; p++) {          // Gauss-Jordan, no pivoting (HPD)
      cd piv = A[7 * p]; double d = 1.0 / norm2(piv); cd ip(piv.re * d, -piv.im * d);
      for (int c = 0; c < 6; c++) { A[6 * p + c] = A[6 * p + c] * ip; R[6 * p + c] = R[6 * p + c] * ip; }
      for (int r = 0; r < 6; r++) if (r != p) {
        cd f = A[6 * r + p];
        for (int c = 0; c < 6; c++) { fms_(A[6 * r + c], f, A[6 * p + c]); fms_(R[6 * r + c], f, R[6 * p + c]); }
      }
    }
    for (int k = 0; k < 6; k++) Cinv[lc.idx(s, 6 * b + k)] = R[7 * k].re;
    m = 0;
    for (int r = 0; r < 6; r++) for (int c = r + 1; c < 6; c++, m++) {
      cd v = cd(0.5, 0.0) * (R[6 * r + c] + conj(R[6 * c + r]));
      Cinv[lc.idx(s, 12 + 2 * (15 * b + m))] = v.re; Cinv[lc.idx(s, 12 + 2 * (15 * b + m) + 1)] = v.im;
    }
  }, 64);
}

// ---------------------------------------------------------------------------------------------------
// layout conversion between the reference's lexicographic site-major arrays and the native tiled layout
// user spinor (lexicographic, 12 complex/site, double) -> native vector of precision T, and back
template <class T> void spinor_from_lex(const Geometry &geo, cx<T> *dst, const cd *src_lex, int ncomp) {
  const Lay l = {ncomp, geo.sh}; const int *n2l = geo.d_nat2lex;
  launch_n(geo.V * ncomp, DLAMBDA(long i) { long s = i / ncomp; int c = (int)(i - s * ncomp); cd v = src_lex[(long)n2l[s] * ncomp + c]; dst[l.idx(s, c)] = cx<T>((T)v.re, (T)v.im); });
}
template <class T> void spinor_to_lex(const Geometry &geo, cd *dst_lex, const cx<T> *src, int ncomp) {
  const Lay l = {ncomp, geo.sh}; const int *n2l = geo.d_nat2lex;
  launch_n(geo.V * ncomp, DLAMBDA(long i) { long s = i / ncomp; int c = (int)(i - s * ncomp); cx<T> v = src[l.idx(s, c)]; dst_lex[(long)n2l[s] * ncomp + c] = cd((double)v.re, (double)v.im); });
}
template void spinor_from_lex<float>(const Geometry &, cf *, const cd *, int);
template void spinor_from_lex<double>(const Geometry &, cd *, const cd *, int);
template void spinor_to_lex<float>(const Geometry &, cd *, const cf *, int);
template void spinor_to_lex<double>(const Geometry &, cd *, const cd *, int);

// reals: lexicographic [site][nk] <-> native Lay{nk,sh}
void reals_from_lex(const Geometry &geo, double *dst, const double *src_lex, int nk) {
  const Lay l = {nk, geo.sh}; const int *n2l = geo.d_nat2lex;
  launch_n(geo.V * nk, DLAMBDA(long i) { long s = i / nk; int c = (int)(i - s * nk); dst[l.idx(s, c)] = src_lex[(long)n2l[s] * nk + c]; });
}
void reals_to_lex(const Geometry &geo, double *dst_lex, const double *src, int nk) {
  const Lay l = {nk, geo.sh}; const int *n2l = geo.d_nat2lex;
  launch_n(geo.V * nk, DLAMBDA(long i) { long s = i / nk; int c = (int)(i - s * nk); dst_lex[(long)n2l[s] * nk + c] = src[l.idx(s, c)]; });
}

}  // namespace dda

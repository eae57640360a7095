// fine_op.h -- host-side interface of the fine Wilson-Clover kernels
#pragma once
#include "common.cuh"
#include "lattice.h"

namespace dda {

// operator data of the fine level in precision T (device pointers, tiled layouts)
template <class T> struct FineOp {
  const cx<T> *D;        // links U/2: Lay{36,sh}: component 9*mu + 3*row + col
  const T *C;            // clover: Lay{72,sh} reals: 12 diagonal, then 30 complex (re,im) = upper triangles of the two 6x6 blocks
  const T *Cinv;         // inverse of the clover blocks (same packing); used on block-odd sites by the SAP Schur complement
  const int *nb;         // [8][V]
  const unsigned char *blkflag, *aggflag;
  long V;
  int sh;
};


// selection of output sites of a kernel: all sites [off, off+n) or, with a block list, the sub-range
// [off, off+cnt) of every listed Schwarz block (cnt = bs: whole block, bs_even: even sites, ...)
struct SiteSel {
  long n;
  const int *blocklist;
  int bs, off, cnt;
};
// thread index i of a launch over (selected sites) x (nc components) -> position in the selection and component;
// groups of 32 consecutive selected sites are kept together so that tiled vectors are accessed coalesced
HD bool sel_decode(long n, int nc, long i, long &si, int &c) {
  long grp = i / (32L * nc); int rem = (int)(i - grp * 32L * nc);
  c = rem >> 5; si = grp * 32 + (rem & 31);
  return si < n;
}
inline long sel_threads(long n, int nc) { return ((n + 31) / 32) * 32 * nc; }
HD long sel_site(const SiteSel &s, long i) {
  if (!s.blocklist) return s.off + i;
  long b = i / s.cnt;
  return (long)s.blocklist[b] * s.bs + s.off + (i - b * s.cnt);
}
inline SiteSel sel_all(long V) { SiteSel s; s.n = V; s.blocklist = nullptr; s.bs = 0; s.off = 0; s.cnt = 0; return s; }
inline SiteSel sel_range(long off, long n) { SiteSel s; s.n = n; s.blocklist = nullptr; s.bs = 0; s.off = (int)off; s.cnt = 0; return s; }
inline SiteSel sel_blocks(const int *list, int nblk, int bs, int off, int cnt) { SiteSel s; s.n = (long)nblk * cnt; s.blocklist = list; s.bs = bs; s.off = off; s.cnt = cnt; return s; }

enum { HOP_NONE = 0, HOP_ALL = 1, HOP_INBLOCK = 2, HOP_INAGG = 3, HOP_CROSSAGG = 4, HOP_CROSSBLOCK = 5 };
enum { OUT_SET = 0, OUT_ADD = 1, OUT_SUB = 2, OUT_ETA_MINUS = 3, OUT_NEG = 4 };
enum { SELF_NONE = 0, SELF_C = 1, SELF_CINV = 2 };

// res(s) = [self term applied to in_self (default: in)] + [selected hops applied to in]; combined into out per outmode
template <class T> void fine_apply(const FineOp<T> &op, cx<T> *out, const cx<T> *in, SiteSel sel, int hop, int dir,
                                   int self, int outmode, const cx<T> *eta = nullptr, const cx<T> *in_self = nullptr);
void fine_build_clover(const Geometry &geo, const cd *D, double *C, double m0, double csw, double *plaq_out);
void fine_shift_clover(const Geometry &geo, double *C, double delta);
void fine_scale_clover(const Geometry &geo, double *C, double se, double so);
void fine_invert_clover(const Geometry &geo, const double *C, double *Cinv);
void cast_links(const cd *src, cf *dst, long n);
void cast_reals(const double *src, float *dst, long n);
template <class T> void spinor_from_lex(const Geometry &geo, cx<T> *dst, const cd *src_lex, int ncomp);
template <class T> void spinor_to_lex(const Geometry &geo, cd *dst_lex, const cx<T> *src, int ncomp);
void reals_from_lex(const Geometry &geo, double *dst, const double *src_lex, int nk);
void reals_to_lex(const Geometry &geo, double *dst_lex, const double *src, int nk);

// optimised full-lattice D_W apply (sm_100a; dw_kernel.cu).  Not available in the emulation build.
template <class T> void dw_apply_fast(const FineOp<T> &op, cx<T> *out, const cx<T> *in, int mode = 0, const int *list = nullptr, long nlist = 0);

// out = eta - (hops that leave the Schwarz block) in, on the sites of the listed blocks (sm_100a; dw_kernel.cu)
void dw_outer_fast(const FineOp<float> &op, cx<float> *out, const cx<float> *in, const cx<float> *eta, const int *blocklist, int nblk, int bs);

}  // namespace dda

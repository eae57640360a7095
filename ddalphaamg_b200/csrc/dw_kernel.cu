// dw_kernel.cu -- hand-tuned full-lattice Wilson-Clover apply for sm_100a (double and float).
//
//   eta(x) = C(x) phi(x) - sum_mu [ (1-gamma_mu) D_mu(x) phi(x+mu) + (1+gamma_mu) D_mu(x-mu)^dagger phi(x-mu) ]
// Reference counterpart: d_plus_clover_PRECISION (dirac_generic.c:159-277), which makes five passes over the volume
// and materialises eight half-spinor fields; here it is ONE gather-form pass: every thread owns one site, reads its
// 8 neighbours, 8 links and the clover block and writes the result once.
//
// Data layout (see common.cuh Lay): 32-site tiles, component-major inside a tile, so lane l of a warp reads
// component c of site (tile*32 + l): every load instruction of a warp is one contiguous 256 B (float) / 512 B
// (double) segment for the site's own data, and a permutation of at most a few such segments for neighbour data
// (sites are ordered Schwarz-block-wise, even sites first, so neighbours of a tile live in few tiles).
// Algorithmic traffic per site: 24 (phi) + 24 (eta) + 72 (4 links) + 72 (clover: 12 real diagonal + 30 complex) reals.
#include "fine_op.h"
#include "fine_op.cuh"

namespace dda {

#ifndef DDA_HOST_EMU

template <class T> struct LdTraits;
template <> struct LdTraits<float> { typedef float2 V; };
template <> struct LdTraits<double> { typedef double2 V; };

template <class T> __device__ __forceinline__ cx<T> ldc(const cx<T> *p) {
  typedef typename LdTraits<T>::V V;
  V v = __ldg(reinterpret_cast<const V *>(p));
  return cx<T>(v.x, v.y);
}

template <int MU, class T>
__device__ __forceinline__ void dw_hop_fwd(const cx<T> *__restrict__ D, const cx<T> *__restrict__ in, long tile_u, int lane, long n, cx<T> *out) {
  // (1-gamma_mu) D_mu(x) phi(x+mu): link of the own site (coalesced), spinor of the +mu neighbour
  const long nt = (n >> 5) * (12L << 5) + (n & 31);
  cx<T> p[12], h[6], g[6], M[9];
#pragma unroll
  for (int c = 0; c < 12; c++) p[c] = ldc(in + nt + ((long)c << 5));
#pragma unroll
  for (int k = 0; k < 9; k++) M[k] = ldc(D + tile_u + ((long)(9 * MU + k) << 5) + lane);
  project<MU, +1>(p, h);
  su3_mul(M, h, g);
  reconstruct_sub<MU, +1>(g, out);
}

template <int MU, class T>
__device__ __forceinline__ void dw_hop_bwd(const cx<T> *__restrict__ D, const cx<T> *__restrict__ in, long n, cx<T> *out) {
  // (1+gamma_mu) D_mu(x-mu)^dagger phi(x-mu): link and spinor of the -mu neighbour
  const long nt = (n >> 5) * (12L << 5) + (n & 31);
  const long nu = (n >> 5) * (36L << 5) + (n & 31);
  cx<T> p[12], h[6], g[6], M[9];
#pragma unroll
  for (int c = 0; c < 12; c++) p[c] = ldc(in + nt + ((long)c << 5));
#pragma unroll
  for (int k = 0; k < 9; k++) M[k] = ldc(D + nu + ((long)(9 * MU + k) << 5));
  project<MU, -1>(p, h);
  su3_mul_dag(M, h, g);
  reconstruct_sub<MU, -1>(g, out);
}

// MODE 0: all sites.  MODE 1: only sites whose neighbours are all local (runs while the halo exchange is in flight).
// MODE 2: the sites of `list` (the rank-boundary sites, after the exchange).
template <class T, int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB)
k_dw_full(const cx<T> *__restrict__ D, const T *__restrict__ C, const int *__restrict__ nb, const cx<T> *__restrict__ in,
          cx<T> *__restrict__ out, long V, const int *__restrict__ list, long nlist) {
  const long i0 = blockIdx.x * (long)BLOCK + threadIdx.x;
  if (MODE == 2 ? (i0 >= nlist) : (i0 >= V)) return;
  const long s = (MODE == 2) ? (long)list[i0] : i0;
  const int lane = (int)(s & 31);
  const long tile = s >> 5;
  const long tile_s = tile * (12L << 5), tile_u = tile * (36L << 5), tile_c = tile * (72L << 5);
  int n[8];
#pragma unroll
  for (int d = 0; d < 8; d++) n[d] = __ldg(nb + (long)d * V + s);
  if (MODE == 1) {
    bool ghost = false;
#pragma unroll
    for (int d = 0; d < 8; d++) ghost = ghost || (n[d] >= V);
    if (ghost) return;
  }

  cx<T> r[12];
  {
    cx<T> x[12];
#pragma unroll
    for (int c = 0; c < 12; c++) x[c] = ldc(in + tile_s + ((long)c << 5) + lane);
    // clover: two Hermitian 6x6 blocks, packed (site_clover_PRECISION, dirac_generic.h:723-799)
    const T *Cs = C + tile_c + lane;
#pragma unroll
    for (int b = 0; b < 2; b++) {
#pragma unroll
      for (int i = 0; i < 6; i++) r[6 * b + i] = __ldg(Cs + ((long)(6 * b + i) << 5)) * x[6 * b + i];
      int m = 0;
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i + 1; j < 6; j++, m++) {
          cx<T> cij(__ldg(Cs + ((long)(12 + 2 * (15 * b + m)) << 5)), __ldg(Cs + ((long)(12 + 2 * (15 * b + m) + 1) << 5)));
          fma_(r[6 * b + i], cij, x[6 * b + j]);
          fmac_(r[6 * b + j], cij, x[6 * b + i]);
        }
    }
  }
  dw_hop_fwd<0>(D, in, tile_u, lane, n[0], r);
  dw_hop_bwd<0>(D, in, n[4], r);
  dw_hop_fwd<1>(D, in, tile_u, lane, n[1], r);
  dw_hop_bwd<1>(D, in, n[5], r);
  dw_hop_fwd<2>(D, in, tile_u, lane, n[2], r);
  dw_hop_bwd<2>(D, in, n[6], r);
  dw_hop_fwd<3>(D, in, tile_u, lane, n[3], r);
  dw_hop_bwd<3>(D, in, n[7], r);
#pragma unroll
  for (int c = 0; c < 12; c++) out[tile_s + ((long)c << 5) + lane] = r[c];
}

// out(s) = eta(s) - [hops of s that LEAVE its Schwarz block] in, for the sites of the listed blocks: the couplings to the
// neighbouring blocks of the SAP block residual (block_PRECISION_boundary_op, schwarz_generic.c:743-856) as a kernel of its
// own with full occupancy -- inside the fused block-solve kernel (8 warps per block, 2 blocks per SM) the dependent global
// loads of these hops were a fifth of the visit (profiles/r2_ncu_full_k_sap_fine2_a.txt).
template <class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 4)
k_dw_outer(const cx<T> *__restrict__ D, const int *__restrict__ nb, const unsigned char *__restrict__ blkflag,
           const cx<T> *__restrict__ in, const cx<T> *__restrict__ eta, cx<T> *__restrict__ out, long V,
           const int *__restrict__ blocklist, long nsel, int bs) {
  const long i0 = blockIdx.x * (long)BLOCK + threadIdx.x;
  if (i0 >= nsel) return;
  const long b = i0 / bs;
  const long s = (long)blocklist[b] * bs + (i0 - b * bs);
  const int lane = (int)(s & 31);
  const long tile = s >> 5;
  const long tile_s = tile * (12L << 5), tile_u = tile * (36L << 5);
  const unsigned f = blkflag[s];
  // the hop functions SUBTRACT (1 -+ gamma) U phi, i.e. they add the operator's off-diagonal part N phi; the residual needs
  // eta - N phi, so the accumulation runs on -eta and the sign is flipped at the end
  cx<T> r[12];
#pragma unroll
  for (int c = 0; c < 12; c++) r[c] = -ldc(eta + tile_s + ((long)c << 5) + lane);
  if (f & 0x01u) dw_hop_fwd<0>(D, in, tile_u, lane, (long)__ldg(nb + 0 * V + s), r);
  if (f & 0x10u) dw_hop_bwd<0>(D, in, (long)__ldg(nb + 4 * V + s), r);
  if (f & 0x02u) dw_hop_fwd<1>(D, in, tile_u, lane, (long)__ldg(nb + 1 * V + s), r);
  if (f & 0x20u) dw_hop_bwd<1>(D, in, (long)__ldg(nb + 5 * V + s), r);
  if (f & 0x04u) dw_hop_fwd<2>(D, in, tile_u, lane, (long)__ldg(nb + 2 * V + s), r);
  if (f & 0x40u) dw_hop_bwd<2>(D, in, (long)__ldg(nb + 6 * V + s), r);
  if (f & 0x08u) dw_hop_fwd<3>(D, in, tile_u, lane, (long)__ldg(nb + 3 * V + s), r);
  if (f & 0x80u) dw_hop_bwd<3>(D, in, (long)__ldg(nb + 7 * V + s), r);
#pragma unroll
  for (int c = 0; c < 12; c++) out[tile_s + ((long)c << 5) + lane] = -r[c];
}

void dw_outer_fast(const FineOp<float> &op, cf *out, const cf *in, const cf *eta, const int *blocklist, int nblk, int bs) {
  DDA_ASSERT(op.sh == 5);
  const long nsel = (long)nblk * bs;
  if (nsel <= 0) return;
  const int BLOCK = 128;
  k_dw_outer<float, BLOCK><<<(unsigned)((nsel + BLOCK - 1) / BLOCK), BLOCK, 0, g_stream>>>(op.D, op.nb, op.blkflag, in, eta, out, op.V, blocklist, nsel, bs);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

template <class T, int MODE> static void dw_launch(const FineOp<T> &op, cx<T> *out, const cx<T> *in, const int *list, long nlist) {
  const int BLOCK = 128;
  const long nthreads = (MODE == 2) ? nlist : op.V;
  if (nthreads <= 0) return;
  unsigned grid = (unsigned)((nthreads + BLOCK - 1) / BLOCK);
  if constexpr (sizeof(T) == 8) k_dw_full<T, BLOCK, 2, MODE><<<grid, BLOCK, 0, g_stream>>>(op.D, op.C, op.nb, in, out, op.V, list, nlist);
  else k_dw_full<T, BLOCK, 4, MODE><<<grid, BLOCK, 0, g_stream>>>(op.D, op.C, op.nb, in, out, op.V, list, nlist);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
}

// mode 0: all sites; 1: interior sites only; 2: the listed (boundary) sites
template <class T> void dw_apply_fast(const FineOp<T> &op, cx<T> *out, const cx<T> *in, int mode, const int *list, long nlist) {
  DDA_ASSERT(op.sh == 5);
  if (mode == 0) dw_launch<T, 0>(op, out, in, nullptr, 0);
  else if (mode == 1) dw_launch<T, 1>(op, out, in, nullptr, 0);
  else dw_launch<T, 2>(op, out, in, list, nlist);
}
template void dw_apply_fast<float>(const FineOp<float> &, cf *, const cf *, int, const int *, long);
template void dw_apply_fast<double>(const FineOp<double> &, cd *, const cd *, int, const int *, long);

#endif

}  // namespace dda

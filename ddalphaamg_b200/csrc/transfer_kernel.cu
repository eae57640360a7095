// transfer_kernel.cu -- hand-tuned restriction for the fine level (sm_100a): phi_c = P^H phi.
//
// Reference counterpart: restrict_PRECISION (interpolation_generic.c:169-207).  One CTA per aggregate; every thread
// owns sites of the aggregate, keeps phi(site) in registers and streams the Nv prolongator vectors ONCE (coalesced
// 256 B rows of the tiled layout), accumulating the 2*Nv partial inner products in registers, 8 test vectors at a
// time; warp-shuffle reduction, one shared-memory atomic per warp and value.  Algorithmic traffic per fine site:
// (12*Nv + 12) * 8 B.
#include "transfer.h"

namespace dda {

#ifndef DDA_HOST_EMU

template <int KC>
__global__ void __launch_bounds__(256)
k_restrict_fine(Transfer t, cf *__restrict__ out, long site_stride, long offset, const cf *__restrict__ phi) {
  __shared__ float acc_s[2 * 2 * MAX_NV];
  const int a = blockIdx.x, as = t.as, nv = t.nv;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * nv; i += blockDim.x) acc_s[i] = 0.f;
  __syncthreads();
  for (int k0 = 0; k0 < nv; k0 += KC) {
    float ar[2][KC], ai[2][KC];
#pragma unroll
    for (int k = 0; k < KC; k++) { ar[0][k] = ai[0][k] = ar[1][k] = ai[1][k] = 0.f; }
    for (int sl = threadIdx.x; sl < as; sl += blockDim.x) {
      const long s = (long)a * as + sl;
      const long base = (s >> 5) * (12L << 5) + (s & 31);
      cf v[12];
#pragma unroll
      for (int c = 0; c < 12; c++) v[c] = phi[base + ((long)c << 5)];
#pragma unroll
      for (int k = 0; k < KC; k++) {
        if (k0 + k < nv) {
          const cf *__restrict__ P = t.P[k0 + k];
#pragma unroll
          for (int c = 0; c < 12; c++) {
            const float2 p = __ldg(reinterpret_cast<const float2 *>(P + base + ((long)c << 5)));
            const int ch = c / 6;
            ar[ch][k] = __fmaf_rn(p.y, v[c].im, __fmaf_rn(p.x, v[c].re, ar[ch][k]));       // conj(P) * phi
            ai[ch][k] = __fmaf_rn(-p.y, v[c].re, __fmaf_rn(p.x, v[c].im, ai[ch][k]));
          }
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 2; ch++)
#pragma unroll
      for (int k = 0; k < KC; k++) {
        float x = ar[ch][k], y = ai[ch][k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, o); y += __shfl_xor_sync(0xffffffffu, y, o); }
        if (lane == 0 && k0 + k < nv) { atomicAdd(&acc_s[2 * (ch * nv + k0 + k)], x); atomicAdd(&acc_s[2 * (ch * nv + k0 + k) + 1], y); }
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * nv; i += blockDim.x)
    out[(long)t.agg2coarse[a] * site_stride + offset + i] = cf(acc_s[2 * i], acc_s[2 * i + 1]);
}

// coarse levels (site-major vectors, nc = 2 Nv' dofs per site): one CTA per aggregate, phi of the aggregate in shared memory,
// every warp takes output components (chirality, k) = w, w + 8, ... and streams P_k over the aggregate with coalesced loads
// (the generic path launches one CTA row per (aggregate, component): 0.15 of the HBM roofline at level 1).
__global__ void __launch_bounds__(256)
k_restrict_coarse(Transfer t, cf *__restrict__ out, long site_stride, long offset, const cf *__restrict__ phi) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *ps = reinterpret_cast<cf *>(smem_raw);
  const int a = blockIdx.x, as = t.as, nv = t.nv, nc = t.nc, h = nc / 2;
  const int len = as * nc, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long base = (long)a * len;
  for (int q = threadIdx.x; q < len; q += 256) ps[q] = phi[base + q];
  __syncthreads();
  const int per = as * h;                                     // elements of one chirality in the aggregate
  for (int j = w; j < 2 * nv; j += 8) {
    const int ch = j / nv, k = j - ch * nv;
    const cf *__restrict__ P = t.P[k] + base;
    float ar = 0.f, ai = 0.f;
    for (int e = lane; e < per; e += 32) {
      const int sl = e / h, q = sl * nc + ch * h + (e - sl * h);
      const float2 p = __ldg(reinterpret_cast<const float2 *>(P + q));
      const cf v = ps[q];
      ar = __fmaf_rn(p.y, v.im, __fmaf_rn(p.x, v.re, ar));    // conj(P) * phi
      ai = __fmaf_rn(-p.y, v.re, __fmaf_rn(p.x, v.im, ai));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
    if (lane == 0) out[(long)t.agg2coarse[a] * site_stride + offset + j] = cf(ar, ai);
  }
}

bool tr_restrict_fast(const Transfer &t, cf *out, long site_stride, long offset, const cf *phi) {
  if (t.lay.sh == 0 && (t.nc & 1) == 0 && (size_t)t.as * t.nc * sizeof(cf) <= 48 * 1024) {
    k_restrict_coarse<<<t.nagg, 256, (size_t)t.as * t.nc * sizeof(cf), g_stream>>>(t, out, site_stride, offset, phi);
    g_launch_count++;
    return true;
  }
  if (!(t.lay.sh == 5 && t.nc == 12 && t.as % 32 == 0)) return false;
  int block = t.as >= 256 ? 256 : t.as;
  k_restrict_fine<4><<<t.nagg, block, 0, g_stream>>>(t, out, site_stride, offset, phi);
  g_launch_count++;
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
}

#endif

}  // namespace dda

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

// -------------------------------------------------------------------------------------------------------------------
// Aggregate orthonormalisation as CholeskyQR in double precision: per (aggregate, chirality) the nv vectors restricted to
// the aggregate form a tall matrix V (as * nc/2 rows); G = V^H V (double accumulation of exact float products), G = R^H R,
// V <- V R^-1 (double accumulation, rounded to float on the store).  Applied twice (CholeskyQR2) the result is the Q factor
// of the unique QR factorisation with positive diagonal, i.e. what Gram-Schmidt in the same vector order computes
// (gram_schmidt_on_aggregates_PRECISION, linalg_generic.c:400-454), with orthogonality at the float rounding level.
// One CTA per (aggregate, chirality): two passes over the aggregate's data (the second one from L2) instead of the
// 2 * (3 k + 2) full-vector passes per vector k of the generic kernels (1 220 passes of 1 GB at 48^3 x 96, Nv = 20).
const int GS_CE = 64;          // elements per staged chunk
const int GS_NVMAX = 32;       // vectors held in registers by the apply phase
const int GS_KG = 4;           // outputs accumulated at a time by the apply phase

struct GsVecs { cf *p[MAX_NV]; };

__global__ void __launch_bounds__(256, 2) k_agg_cholqr(Transfer t, GsVecs vp) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  const int nv = t.nv, h = t.nc / 2, as = t.as, E = as * h;
  const long a = blockIdx.x >> 1;
  const int ch = blockIdx.x & 1, tid = threadIdx.x;
  double2 *chunk = reinterpret_cast<double2 *>(gs_smem);             // [nv][GS_CE + 1]
  double2 *A = chunk + (size_t)nv * (GS_CE + 1);                     // [nv][nv]: G, then R (upper triangle)
  double2 *T = A + (size_t)nv * nv;                                  // [nv][nv]: R^-1 (upper triangle)
  double *inv = reinterpret_cast<double *>(T + (size_t)nv * nv);     // [nv] 1 / R_jj (0 for a vanishing vector)
  const Lay lay = t.lay;

  // ---- G[k1][k2] = <v_k1, v_k2>, k2 <= k1: up to two entries per thread
  const int nent = nv * (nv + 1) / 2;
  int k1s[2], k2s[2];
  double2 acc[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int ent = tid + 256 * q;
    k1s[q] = -1; k2s[q] = 0; acc[q] = make_double2(0.0, 0.0);
    if (ent < nent) {
      int k1 = (int)((sqrtf(8.f * ent + 1.f) - 1.f) * 0.5f);
      while (k1 * (k1 + 1) / 2 > ent) k1--;
      while ((k1 + 1) * (k1 + 2) / 2 <= ent) k1++;
      k1s[q] = k1; k2s[q] = ent - k1 * (k1 + 1) / 2;
    }
  }
  for (int e0 = 0; e0 < E; e0 += GS_CE) {
    for (int i = tid; i < nv * GS_CE; i += 256) {
      const int k = i / GS_CE, e = i - k * GS_CE, ee = e0 + e;
      double2 val = make_double2(0.0, 0.0);
      if (ee < E) {
        const int c = ee / as, sl = ee - c * as;
        const cf v = vp.p[k][lay.idx(a * as + sl, ch * h + c)];
        val = make_double2((double)v.re, (double)v.im);
      }
      chunk[k * (GS_CE + 1) + e] = val;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (k1s[q] >= 0) {
        const double2 *x = chunk + k1s[q] * (GS_CE + 1), *y = chunk + k2s[q] * (GS_CE + 1);
        double sr = acc[q].x, si = acc[q].y;
#pragma unroll 4
        for (int e = 0; e < GS_CE; e++) {
          const double2 xv = x[e], yv = y[e];
          sr = fma(xv.x, yv.x, fma(xv.y, yv.y, sr));
          si = fma(xv.x, yv.y, fma(-xv.y, yv.x, si));
        }
        acc[q] = make_double2(sr, si);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 2; q++) {
    if (k1s[q] >= 0) {
      A[k1s[q] * nv + k2s[q]] = acc[q];                                         // G[k1][k2]
      if (k2s[q] != k1s[q]) A[k2s[q] * nv + k1s[q]] = make_double2(acc[q].x, -acc[q].y);   // G[k2][k1] = conj
    }
  }
  __syncthreads();

  // ---- Cholesky G = R^H R, R upper triangular with positive diagonal (right-looking, row j of R per step)
  for (int j = 0; j < nv; j++) {
    if (tid == 0) {
      const double d = A[j * nv + j].x;
      const double r = d > 0.0 ? sqrt(d) : 0.0;
      A[j * nv + j] = make_double2(r, 0.0);
      inv[j] = r > 0.0 ? 1.0 / r : 0.0;
    }
    __syncthreads();
    const double ij = inv[j];
    for (int i = j + 1 + tid; i < nv; i += 256) { double2 v = A[j * nv + i]; A[j * nv + i] = make_double2(v.x * ij, v.y * ij); }
    __syncthreads();
    const int m = nv - 1 - j;                                                   // trailing block (p, q), j < p <= q
    for (int i = tid; i < m * m; i += 256) {
      const int p = j + 1 + i / m, q = j + 1 + i % m;
      if (q >= p) {
        const double2 rp = A[j * nv + p], rq = A[j * nv + q];                   // A[p][q] -= conj(R[j][p]) R[j][q]
        double2 v = A[p * nv + q];
        v.x -= rp.x * rq.x + rp.y * rq.y;
        v.y -= rp.x * rq.y - rp.y * rq.x;
        A[p * nv + q] = v;
      }
    }
    __syncthreads();
  }
  // ---- T = R^-1: column k by back substitution (one thread per column)
  for (int k = tid; k < nv; k += 256) {
    for (int i = 0; i < nv; i++) T[i * nv + k] = make_double2(0.0, 0.0);
    T[k * nv + k] = make_double2(inv[k], 0.0);
    for (int i = k - 1; i >= 0; i--) {
      double sr = 0.0, si = 0.0;
      for (int m = i + 1; m <= k; m++) {
        const double2 r = A[i * nv + m], tv = T[m * nv + k];
        sr += r.x * tv.x - r.y * tv.y; si += r.x * tv.y + r.y * tv.x;
      }
      T[i * nv + k] = make_double2(-sr * inv[i], -si * inv[i]);
    }
  }
  __syncthreads();

  // ---- V <- V T: one element per thread at a time, the nv input values in registers, outputs in groups of GS_KG (every input is
  // converted to double once per group, not once per product)
  for (int ee = tid; ee < E; ee += 256) {
    const int c = ee / as, sl = ee - c * as;
    const long q = lay.idx(a * as + sl, ch * h + c);
    cf x[GS_NVMAX];
#pragma unroll
    for (int i = 0; i < GS_NVMAX; i++) x[i] = (i < nv) ? vp.p[i][q] : cf(0.f, 0.f);
#pragma unroll
    for (int k0 = 0; k0 < GS_NVMAX; k0 += GS_KG) {
      if (k0 < nv) {
        double sr[GS_KG], si[GS_KG];
#pragma unroll
        for (int j = 0; j < GS_KG; j++) { sr[j] = 0.0; si[j] = 0.0; }
#pragma unroll
        for (int i = 0; i < k0 + GS_KG; i++) {
          if (i < nv) {
            const double xr = (double)x[i].re, xi = (double)x[i].im;
#pragma unroll
            for (int j = 0; j < GS_KG; j++) {
              if (k0 + j >= i && k0 + j < nv) {
                const double2 tv = T[i * nv + k0 + j];
                sr[j] = fma(xr, tv.x, fma(-xi, tv.y, sr[j]));
                si[j] = fma(xr, tv.y, fma(xi, tv.x, si[j]));
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < GS_KG; j++)
          if (k0 + j < nv) vp.p[k0 + j][q] = cf((float)sr[j], (float)si[j]);
      }
    }
  }
}

// CholeskyQR2 over all aggregates; returns false when the shape is not supported (caller: generic Gram-Schmidt kernels)
bool tr_gram_schmidt_fast(const Transfer &t, cf *const *vecs) {
  if (t.nv > GS_NVMAX || t.nv < 1 || (t.nc & 1) || t.nagg <= 0) return false;
  const size_t smem = ((size_t)t.nv * (GS_CE + 1) + 2 * (size_t)t.nv * t.nv) * sizeof(double2) + t.nv * sizeof(double) + 16;
  static size_t attr = 0;
  if (smem > attr) { CUDA_CHECK(cudaFuncSetAttribute(k_agg_cholqr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  GsVecs vp;
  for (int k = 0; k < MAX_NV; k++) vp.p[k] = k < t.nv ? vecs[k] : nullptr;
  for (int pass = 0; pass < 2; pass++) {
    k_agg_cholqr<<<(unsigned)(2 * t.nagg), 256, smem, g_stream>>>(t, vp);
    g_launch_count++;
  }
#ifdef DDA_DEBUG_SYNC
  CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError());
#endif
  return true;
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

// halo.cu -- see halo.h.  Pack kernel (gather the boundary slice into a contiguous face in the slab's own layout),
// grouped send/recv with both neighbours of every partitioned direction, receive directly into the ghost slab.
#include "halo.h"

namespace dda {

long g_halo_bytes = 0;

template <class E> void halo_exchange(const Geometry &g, E *v, int nc, int sh) {
  if (!g.partitioned()) return;
  const Lay lay = {nc, sh};
  for (int m = 0; m < 4; m++) {
    if (!g.split(m)) continue;
    const long ns = g.slab[m];
    const size_t bytes = sizeof(E) * (size_t)ns * nc;
    E *b0 = (E *)comm_buffer(0, bytes), *b1 = (E *)comm_buffer(1, bytes);
    const int *s0 = g.d_slice[m], *s1 = g.d_slice[4 + m];
    launch_n(2 * ns * nc, DLAMBDA(long q) {
      const long half = ns * nc;
      const int side = q >= half; const long r = side ? q - half : q;
      long i; int c; lay.decode(r, i, c);
      const int *sl = side ? s1 : s0;
      E *b = side ? b1 : b0;
      b[r] = v[lay.idx(sl[i], c)];
    });
    // my x_m = 0 slice -> the -m neighbour's +m slab ; my x_m = L-1 slice -> the +m neighbour's -m slab
    comm_group_begin();
    comm_sendrecv(b0, v + g.gh_off[m] * nc, bytes, g.nbr_rank[4 + m], g.nbr_rank[m]);
    comm_sendrecv(b1, v + g.gh_off[4 + m] * nc, bytes, g.nbr_rank[m], g.nbr_rank[4 + m]);
    comm_group_end();
    g_halo_bytes += 2 * (long)bytes;
  }
}
#ifndef DDA_HOST_EMU
static cudaStream_t g_halo_stream = nullptr;
static cudaEvent_t g_ev_ready = nullptr, g_ev_done = nullptr;
template <class E> void halo_begin(const Geometry &g, E *v, int nc, int sh) {
  if (!g.partitioned()) return;
  if (!g_halo_stream) {
    int lo = 0, hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // the copy/NCCL kernels must get SM slots ahead of the
    CUDA_CHECK(cudaStreamCreateWithPriority(&g_halo_stream, cudaStreamNonBlocking, hi));   // interior kernel's CTAs
    CUDA_CHECK(cudaEventCreateWithFlags(&g_ev_ready, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_ev_done, cudaEventDisableTiming));
  }
  long smax = 0;
  for (int m = 0; m < 4; m++) smax = std::max(smax, g.slab[m]);
  comm_buffer(0, sizeof(E) * (size_t)smax * nc); comm_buffer(1, sizeof(E) * (size_t)smax * nc);   // grow (may sync) before forking
  CUDA_CHECK(cudaEventRecord(g_ev_ready, g_stream));
  CUDA_CHECK(cudaStreamWaitEvent(g_halo_stream, g_ev_ready, 0));
  {
    // pack kernel + NCCL calls of halo_exchange go to the second stream: the library's current stream is switched for the
    // scope of this block (single host thread; halo_end always precedes the next collective on the compute stream)
    struct StreamScope { cudaStream_t saved; StreamScope(cudaStream_t s) : saved(g_stream) { g_stream = s; } ~StreamScope() { g_stream = saved; } } scope(g_halo_stream);
    halo_exchange<E>(g, v, nc, sh);
  }
  CUDA_CHECK(cudaEventRecord(g_ev_done, g_halo_stream));
}
void halo_end(const Geometry &g) {
  if (!g.partitioned()) return;
  CUDA_CHECK(cudaStreamWaitEvent(g_stream, g_ev_done, 0));
}
void halo_finalize() {
  if (g_halo_stream) {
    cudaStreamSynchronize(g_halo_stream);
    cudaEventDestroy(g_ev_ready); cudaEventDestroy(g_ev_done); cudaStreamDestroy(g_halo_stream);
    g_halo_stream = nullptr; g_ev_ready = g_ev_done = nullptr;
  }
}
#else
template <class E> void halo_begin(const Geometry &g, E *v, int nc, int sh) { halo_exchange<E>(g, v, nc, sh); }
void halo_end(const Geometry &) {}
void halo_finalize() {}
#endif
template void halo_begin<cf>(const Geometry &, cf *, int, int);
template void halo_begin<cd>(const Geometry &, cd *, int, int);
template void halo_exchange<cf>(const Geometry &, cf *, int, int);
template void halo_exchange<cd>(const Geometry &, cd *, int, int);

}  // namespace dda

// comm_layer.cu -- see comm.h.  NCCL in the product build, harness callbacks in the host-emulation build.
#include "comm.h"
#include "../../include/dd_alpha_amg_b200.h"
#ifndef DDA_HOST_EMU
#include <nccl.h>
#endif

namespace dda {

Comm g_comm;
static void *g_buf[2] = {nullptr, nullptr};
static size_t g_buf_bytes[2] = {0, 0};

void *comm_buffer(int which, size_t bytes) {
  if (bytes > g_buf_bytes[which]) {
    if (g_buf[which]) { dev_sync(); dev_free(g_buf[which]); }
    g_buf_bytes[which] = bytes + bytes / 4;
    g_buf[which] = dev_alloc_bytes(g_buf_bytes[which]);
  }
  return g_buf[which];
}

#ifndef DDA_HOST_EMU
static ncclComm_t g_nccl = nullptr;
#define NCCL_CHECK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { \
  fprintf(stderr, "NCCL error %s: %s\n", #x, ncclGetErrorString(r_)); ::dda::fatal("nccl", __FILE__, __LINE__); } } while (0)

void comm_sendrecv(const void *send, void *recv, size_t bytes, int to, int from) {
  if (to == g_comm.rank && from == g_comm.rank) { d2d(recv, send, bytes); return; }   // own periodic neighbour
  DDA_ASSERT(g_nccl);
  NCCL_CHECK(ncclGroupStart());
  NCCL_CHECK(ncclSend(send, bytes, ncclChar, to, g_nccl, g_stream));
  NCCL_CHECK(ncclRecv(recv, bytes, ncclChar, from, g_nccl, g_stream));
  NCCL_CHECK(ncclGroupEnd());
}
void comm_group_begin() { if (g_nccl) NCCL_CHECK(ncclGroupStart()); }
void comm_group_end() { if (g_nccl) NCCL_CHECK(ncclGroupEnd()); }
void comm_allreduce_sum(double *buf, int n) {
  if (!g_comm.active()) return;
  NCCL_CHECK(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, g_nccl, g_stream));
}
void comm_allgather(const void *send, void *recv, size_t bytes) {
  if (!g_comm.active()) { d2d(recv, send, bytes); return; }
  NCCL_CHECK(ncclAllGather(send, recv, bytes, ncclChar, g_nccl, g_stream));
}
void comm_finalize() {
  for (int i = 0; i < 2; i++) { if (g_buf[i]) dev_free(g_buf[i]); g_buf[i] = nullptr; g_buf_bytes[i] = 0; }
  if (g_nccl) { ncclCommDestroy(g_nccl); g_nccl = nullptr; }
  g_comm = Comm();
}
#else
static dda_sendrecv_fn g_cb_sendrecv = nullptr;
static dda_allreduce_fn g_cb_allreduce = nullptr;
void comm_sendrecv(const void *send, void *recv, size_t bytes, int to, int from) {
  if (to == g_comm.rank && from == g_comm.rank) { d2d(recv, send, bytes); return; }   // own periodic neighbour
  DDA_ASSERT(g_cb_sendrecv);
  g_cb_sendrecv(send, recv, (long)bytes, to, from);
}
void comm_group_begin() {}
void comm_group_end() {}
void comm_allgather(const void *send, void *recv, size_t bytes) {   // ring over the harness' send/recv callback
  const int r = g_comm.rank, p = g_comm.size;
  char *out = (char *)recv;
  d2d(out + (size_t)r * bytes, send, bytes);
  for (int s = 1; s < p; s++) {
    const int sb = (r - s + 1 + p) % p, rb = (r - s + p) % p;
    comm_sendrecv(out + (size_t)sb * bytes, out + (size_t)rb * bytes, bytes, (r + 1) % p, (r - 1 + p) % p);
  }
}
void comm_allreduce_sum(double *buf, int n) {
  if (!g_comm.active()) return;
  DDA_ASSERT(g_cb_allreduce);
  g_cb_allreduce(buf, n);
}
void comm_finalize() {
  for (int i = 0; i < 2; i++) { if (g_buf[i]) dev_free(g_buf[i]); g_buf[i] = nullptr; g_buf_bytes[i] = 0; }
  g_comm = Comm(); g_cb_sendrecv = nullptr; g_cb_allreduce = nullptr;
}
#endif

}  // namespace dda

using namespace dda;

extern "C" {

int dda_comm_unique_id(char *out, int len) {
#ifndef DDA_HOST_EMU
  if (len < (int)sizeof(ncclUniqueId)) return -1;
  ncclUniqueId id;
  NCCL_CHECK(ncclGetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return (int)sizeof(id);
#else
  (void)out; (void)len;
  return 0;
#endif
}

void dda_comm_init(int rank, int size, const char *unique_id, int device) {
#ifndef DDA_HOST_EMU
  DDA_ASSERT(size >= 1 && rank >= 0 && rank < size);
  if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
  if (!g_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  if (size > 1) {
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    NCCL_CHECK(ncclCommInitRank(&g_nccl, size, id, rank));
  }
  g_comm.rank = rank; g_comm.size = size;
#else
  (void)rank; (void)size; (void)unique_id; (void)device;
  fatal("dda_comm_init: the host-emulation build uses dda_comm_init_callbacks", __FILE__, __LINE__);
#endif
}

void dda_comm_init_callbacks(int rank, int size, dda_sendrecv_fn sr, dda_allreduce_fn ar) {
#ifdef DDA_HOST_EMU
  g_comm.rank = rank; g_comm.size = size; g_cb_sendrecv = sr; g_cb_allreduce = ar;
#else
  (void)rank; (void)size; (void)sr; (void)ar;
  fatal("dda_comm_init_callbacks exists only in the host-emulation (test) build; use dda_comm_init", __FILE__, __LINE__);
#endif
}

void dda_comm_finalize(void) { comm_finalize(); }
int dda_comm_rank(void) { return g_comm.rank; }
int dda_comm_size(void) { return g_comm.size; }

}  // extern "C"

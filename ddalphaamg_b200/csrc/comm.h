// comm.h -- process grid + communication primitives (one process per GPU).
//
// Reference counterparts: cart_define / neighbor_define (ghost.c:24-66), ghost_sendrecv / ghost_update
// (ghost_generic.c:171-414: MPI_Isend/Irecv halos), MPI_Allreduce in global_inner_product / global_norm
// (linalg_generic.c:57,201).  Here: ncclSend/ncclRecv pairs and ncclAllReduce on the library's compute stream
// (NVLink 5 / NVSwitch).  The host-emulation build (tests only)
// routes the same calls through callbacks that the test harness implements with torch.distributed/gloo.
#pragma once
#include "common.cuh"

namespace dda {

struct Comm {
  int rank = 0, size = 1;
  bool active() const { return size > 1; }
};
extern Comm g_comm;

// send `bytes` from `send` to rank `to` and receive `bytes` into `recv` from rank `from` (device buffers in the
// CUDA build, ordered on g_stream; several calls may be bracketed by comm_group_begin/end)
void comm_sendrecv(const void *send, void *recv, size_t bytes, int to, int from);
void comm_group_begin();
void comm_group_end();
// in-place sum over all ranks of n doubles (device buffer in the CUDA build)
void comm_allreduce_sum(double *buf, int n);
// recv[r * bytes ...] = send of rank r, for all ranks (device buffers in the CUDA build; ncclAllGather)
void comm_allgather(const void *send, void *recv, size_t bytes);
// scratch buffers for packed faces (two, grown on demand)
void *comm_buffer(int which, size_t bytes);
void comm_finalize();

}  // namespace dda

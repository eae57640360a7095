// halo.h -- ghost-slab exchange of a level's vectors / operator arrays (see lattice.h for the slab layout).
// Reference counterparts: ghost_sendrecv_PRECISION / ghost_update_PRECISION (ghost_generic.c:171-414).
#pragma once
#include "common.cuh"
#include "lattice.h"
#include "comm.h"

namespace dda {

// fills the ghost slabs of `v` (E = element type, nc elements per site, layout Lay{nc, sh}) from the neighbour ranks.
// No-op when the level is not partitioned.
template <class E> void halo_exchange(const Geometry &g, E *v, int nc, int sh);
// split form for overlap: halo_begin starts the exchange of `v` on a second stream (after everything already queued on
// the compute stream), the caller queues work that does not touch the ghost slabs, halo_end makes the compute stream
// wait for the exchange.  Without a GPU build (host emulation) halo_begin does the whole exchange and halo_end nothing.
template <class E> void halo_begin(const Geometry &g, E *v, int nc, int sh);
void halo_end(const Geometry &g);
void halo_finalize();    // destroys the second stream and its events (recreated on demand); called by dd_alpha_amg_free / comm_finalize
extern long g_halo_bytes;   // bytes sent by this rank (statistics)

}  // namespace dda

// lattice.h -- per-level lattice geometry, site ordering and index tables.
//
// Native site order of a level (all device vectors of the level use it):
//   aggregates (lexicographic, T slowest) -> Schwarz blocks inside the aggregate (lexicographic)
//   -> sites of the block: fine level = even sites then odd sites (parity relative to the block origin),
//      coarse intermediate levels = lexicographic;
//   coarsest level: all even sites (lexicographic) then all odd sites (global even-odd order).
// This is the same ordering idea the reference uses for its Schwarz-ordered operator
// (schwarz_generic.c:368-474) and its coarsest even-odd operator (data_layout.c:43-100), but here it is the
// ONLY order of the level: the outer double-precision solver runs in it too, so no per-iteration
// lexicographic<->Schwarz permutation (reference trans_PRECISION, schwarz_generic.c:1807-1846) is needed.
#pragma once
#include "common.cuh"

namespace dda {

enum { T_ = 0, Z_ = 1, Y_ = 2, X_ = 3 };

struct Geometry {
  int L[4] = {0, 0, 0, 0};   // local lattice, order T Z Y X
  int B[4] = {0, 0, 0, 0};   // Schwarz block
  int A[4] = {0, 0, 0, 0};   // aggregate (= coarsening towards the next level); A[0]==0 on the coarsest level
  long V = 0;                // sites
  int nc = 12;               // complex dofs per site
  int sh = 0;                // tile shift of the vector layout (5 on the fine level)
  bool block_eo = false;     // even-odd order inside blocks
  bool global_eo = false;    // coarsest level: global even-odd order
  int bs = 0, nblocks = 0, nblk_color[2] = {0, 0};
  int as = 0, nagg = 0;
  long n_even = 0;           // global_eo: number of even sites; block_eo: even sites per block (bs_even)
  int bs_even = 0;
  // ---- domain decomposition (one process per GPU): process grid, this rank's coordinates, ghost slabs.
  // Vector arrays of the level hold V local sites followed by Vg ghost sites: for every partitioned direction mu a
  // +mu slab (copy of the +mu neighbour rank's x_mu = 0 slice) and a -mu slab (the -mu neighbour's x_mu = L-1 slice);
  // the neighbour table points into the slabs.  Directions are processed in the order T, Z, Y, X and the slab of a
  // direction spans the lattice extended by the slabs of the earlier ones (corner sites for the clover term).
  // Reference: ghost shell of data_layout.c:24-40, ghost_generic.c.
  int P[4] = {1, 1, 1, 1}, pc[4] = {0, 0, 0, 0};
  // fg[mu] != 0: direction mu carries ghost slabs although the process grid has extent 1 there -- the rank is its own
  // periodic neighbour (slabs filled by a local copy).  Runs every partitioned code path (interior / boundary kernels,
  // pack kernel, ghost branches) on ONE GPU; set by DDA_FORCE_SPLIT for the parity tests of those paths.
  int fg[4] = {0, 0, 0, 0};
  bool split(int m) const { return P[m] > 1 || fg[m] != 0; }
  long Vg = 0;
  long gh_off[8] = {-1, -1, -1, -1, -1, -1, -1, -1};   // first site index (>= V) of slab d (d<4: +mu, d>=4: -mu)
  long slab[4] = {0, 0, 0, 0};
  int nbr_rank[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int *d_slice[8] = {};                                 // d<4: local sites with x_mu = 0, d>=4: x_mu = L-1 (native order)
  int *d_nbg = nullptr;                                 // [8][Vg] neighbour table of the ghost sites (-1: not present)
  int *d_bnd = nullptr; long nbnd = 0;                  // local sites with at least one ghost neighbour (sorted)
  int *d_blocklist_int[2] = {nullptr, nullptr}, *d_blocklist_bnd[2] = {nullptr, nullptr};   // per colour: blocks without /
  int nblk_int[2] = {0, 0}, nblk_bnd[2] = {0, 0};       //   with sites on the rank boundary (halo overlap)
  std::vector<int> lex2nat, nat2lex, block_color;
  std::vector<int> h_nb;     // [8][V]
  std::vector<unsigned char> h_blkflag;   // [V]
  int *d_sapslotsite = nullptr;                 // [4][192]: block-local site of link slot (mu, slot)
  unsigned *d_saptab = nullptr;                 // fine level, 4^4 even-odd blocks: per block-local site {in-block mask, +mu / -mu
                                                // neighbour index inside its parity, link slots} (fused SAP kernel v2)
  int *d_sapjobs = nullptr; int nsapjobs = 0;   // block operator as a job list {type, i, j, 0}: type 0 = self coupling of
                                                // block-local site i, 1+mu = in-block link i -> j (same for every block)
  // device tables
  int *d_nb = nullptr;                 // [8][V] dir 0..3 = +T,+Z,+Y,+X ; 4..7 = -T,-Z,-Y,-X
  unsigned char *d_blkflag = nullptr;  // bit d: neighbour d is outside the Schwarz block
  unsigned char *d_aggflag = nullptr;  // bit d: neighbour d is outside the aggregate
  int *d_blocklist[2] = {nullptr, nullptr};
  int *d_lex2nat = nullptr, *d_nat2lex = nullptr;
  int *d_agg2coarse = nullptr;         // aggregate index -> native site index on the next coarser level

  Lay lay() const { Lay l; l.nc = nc; l.sh = sh; return l; }
  long vlen() const { return V * nc; }   // complex elements of the local part of a vector
  long valloc() const { return (V + Vg) * nc; }   // complex elements to allocate (local + ghost slabs)
  bool partitioned() const { return Vg > 0; }
  bool coarsest() const { return A[0] == 0; }
  void build();
  void destroy();
  long lex(int t, int z, int y, int x) const { return x + (long)L[3] * (y + (long)L[2] * (z + (long)L[1] * t)); }
};

}  // namespace dda

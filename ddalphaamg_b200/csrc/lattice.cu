// lattice.cu -- builds the site ordering and neighbour / boundary-flag tables of one level (host, once).
#include "lattice.h"

namespace dda {

void Geometry::build() {
  V = (long)L[0] * L[1] * L[2] * L[3];
  DDA_ASSERT(V > 0 && V < (1L << 31));
  const bool last = coarsest();
  int Bq[4], Aq[4];
  for (int m = 0; m < 4; m++) {
    Aq[m] = last ? L[m] : A[m];
    Bq[m] = last ? L[m] : B[m];
    DDA_ASSERT(L[m] % Aq[m] == 0 && Aq[m] % Bq[m] == 0);
  }
  bs = Bq[0] * Bq[1] * Bq[2] * Bq[3];
  as = Aq[0] * Aq[1] * Aq[2] * Aq[3];
  nblocks = (int)(V / bs);
  nagg = (int)(V / as);
  if (sh > 0) DDA_ASSERT(V % (1L << sh) == 0);

  // in-block position table
  std::vector<int> inblock(bs);
  {
    int ne = 0, no = 0, k = 0;
    for (int t = 0; t < Bq[0]; t++) for (int z = 0; z < Bq[1]; z++) for (int y = 0; y < Bq[2]; y++) for (int x = 0; x < Bq[3]; x++, k++)
      if (((t + z + y + x) & 1) == 0) ne++;
    bs_even = ne;
    int ce = 0; no = 0; k = 0;
    for (int t = 0; t < Bq[0]; t++) for (int z = 0; z < Bq[1]; z++) for (int y = 0; y < Bq[2]; y++) for (int x = 0; x < Bq[3]; x++, k++) {
      if (block_eo && !last) { if (((t + z + y + x) & 1) == 0) inblock[k] = ce++; else inblock[k] = ne + no++; }
      else inblock[k] = k;
    }
  }
  lex2nat.assign(V, 0); nat2lex.assign(V, 0);
  block_color.assign(nblocks, 0);
  std::vector<int> coord(4 * V);
  if (last && global_eo) {
    const int poff = (pc[0] * L[0] + pc[1] * L[1] + pc[2] * L[2] + pc[3] * L[3]) & 1;   // parity of the rank's origin
    long ne = 0;
    for (int t = 0; t < L[0]; t++) for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++) for (int x = 0; x < L[3]; x++)
      if (((t + z + y + x + poff) & 1) == 0) ne++;
    n_even = ne;
    DDA_ASSERT(2 * ne == V);   // every rank holds the same number of even and odd sites
    long ce = 0, co = 0;
    for (int t = 0; t < L[0]; t++) for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++) for (int x = 0; x < L[3]; x++) {
      long i = lex(t, z, y, x);
      long k = (((t + z + y + x + poff) & 1) == 0) ? ce++ : ne + co++;
      lex2nat[i] = (int)k;
    }
  } else {
    int na[4], nbpa[4];
    for (int m = 0; m < 4; m++) { na[m] = L[m] / Aq[m]; nbpa[m] = Aq[m] / Bq[m]; }
    int bpa = nbpa[0] * nbpa[1] * nbpa[2] * nbpa[3];
    for (int t = 0; t < L[0]; t++) for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++) for (int x = 0; x < L[3]; x++) {
      int c[4] = {t, z, y, x};
      long ai = 0, bi = 0, li = 0; int bsum = 0;
      for (int m = 0; m < 4; m++) {
        ai = ai * na[m] + c[m] / Aq[m];
        bi = bi * nbpa[m] + (c[m] % Aq[m]) / Bq[m];
        li = li * Bq[m] + c[m] % Bq[m];
        bsum += (pc[m] * L[m] + c[m]) / Bq[m];
      }
      long blk = ai * bpa + bi;
      lex2nat[lex(t, z, y, x)] = (int)(blk * bs + inblock[li]);
      block_color[blk] = bsum & 1;
    }
    n_even = bs_even;
  }
  for (int t = 0; t < L[0]; t++) for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++) for (int x = 0; x < L[3]; x++) {
    long i = lex(t, z, y, x); int k = lex2nat[i];
    nat2lex[k] = (int)i;
    coord[4 * (long)k + 0] = t; coord[4 * (long)k + 1] = z; coord[4 * (long)k + 2] = y; coord[4 * (long)k + 3] = x;
  }
  // Ghost slabs of the partitioned directions, processed in the order T, Z, Y, X.  The slab of direction m is taken
  // over the lattice already EXTENDED by the ghosts of the directions processed before it, so after the exchanges (same
  // order) the corner sites x +- T +- Z that the clover term needs are present as well.  Slab-local order = lexicographic
  // in the (extended) coordinates: identical on every rank (the native order of the coarsest level depends on the parity
  // of the rank's origin).  Slices may therefore contain ghost sites of earlier directions.
  Vg = 0;
  std::vector<int> slices[8];
  const int E[4] = {L[0] + 2, L[1] + 2, L[2] + 2, L[3] + 2};
  auto eidx = [&](const int *c) { return (((long)(c[0] + 1) * E[1] + (c[1] + 1)) * E[2] + (c[2] + 1)) * E[3] + (c[3] + 1); };
  std::vector<int> ext((size_t)E[0] * E[1] * E[2] * E[3], -1);
  for (long k = 0; k < V; k++) ext[eidx(&coord[4 * k])] = (int)k;
  std::vector<int> gcoord;                      // coordinates of the ghost sites, 4 per site
  int lo[4] = {0, 0, 0, 0}, hi[4] = {L[0] - 1, L[1] - 1, L[2] - 1, L[3] - 1};
  for (int m = 0; m < 4; m++) {
    slab[m] = 0; gh_off[m] = gh_off[4 + m] = -1;
    if (!split(m)) continue;
    for (int side = 0; side < 2; side++) {
      const int d = side == 0 ? m : 4 + m;       // +m ghost <- neighbour's x_m = 0 slice ; -m ghost <- its x_m = L-1 slice
      const int src = side == 0 ? 0 : L[m] - 1, dst = side == 0 ? L[m] : -1;
      gh_off[d] = V + Vg;
      int c[4];
      long cnt = 0;
      for (c[0] = lo[0]; c[0] <= hi[0]; c[0]++) for (c[1] = lo[1]; c[1] <= hi[1]; c[1]++)
        for (c[2] = lo[2]; c[2] <= hi[2]; c[2]++) for (c[3] = lo[3]; c[3] <= hi[3]; c[3]++) {
          if (c[m] != lo[m]) continue;           // the cross-section: every other coordinate once
          int cs[4] = {c[0], c[1], c[2], c[3]}, cg[4] = {c[0], c[1], c[2], c[3]};
          cs[m] = src; cg[m] = dst;
          const int si = ext[eidx(cs)];
          DDA_ASSERT(si >= 0);
          slices[d].push_back(si);
          ext[eidx(cg)] = (int)(V + Vg + cnt);
          for (int q = 0; q < 4; q++) gcoord.push_back(cg[q]);
          cnt++;
        }
      if (side == 0) slab[m] = cnt; else DDA_ASSERT(slab[m] == cnt);
      Vg += cnt;
    }
    if (sh > 0) DDA_ASSERT(slab[m] % (1L << sh) == 0);
    lo[m] = -1; hi[m] = L[m];
    int cp[4] = {pc[0], pc[1], pc[2], pc[3]}, cm[4] = {pc[0], pc[1], pc[2], pc[3]};
    cp[m] = (pc[m] + 1) % P[m]; cm[m] = (pc[m] + P[m] - 1) % P[m];
    nbr_rank[m] = ((cp[0] * P[1] + cp[1]) * P[2] + cp[2]) * P[3] + cp[3];
    nbr_rank[4 + m] = ((cm[0] * P[1] + cm[1]) * P[2] + cm[2]) * P[3] + cm[3];
  }
  DDA_ASSERT(V + Vg < (1L << 31));
  // neighbour of the (possibly ghost) site with coordinates c in direction d; -1 if it is not part of the extended lattice
  auto neighbour = [&](const int *c, int d) {
    const int m = d & 3, sgn = d < 4 ? +1 : -1;
    int q[4] = {c[0], c[1], c[2], c[3]};
    q[m] += sgn;
    if (!split(m)) q[m] = (q[m] + L[m]) % L[m];
    else if (q[m] < -1 || q[m] > L[m]) return -1;
    return ext[eidx(q)];
  };
  h_nb.assign(8 * V, 0);
  std::vector<unsigned char> bf(V, 0), af(V, 0);
  for (long k = 0; k < V; k++) {
    int *c = &coord[4 * k];
    for (int d = 0; d < 8; d++) { h_nb[(long)d * V + k] = neighbour(c, d); DDA_ASSERT(h_nb[(long)d * V + k] >= 0); }
    for (int m = 0; m < 4; m++) {
      if (c[m] % Bq[m] == Bq[m] - 1) bf[k] |= (unsigned char)(1u << m);
      if (c[m] % Bq[m] == 0) bf[k] |= (unsigned char)(1u << (4 + m));
      if (c[m] % Aq[m] == Aq[m] - 1) af[k] |= (unsigned char)(1u << m);
      if (c[m] % Aq[m] == 0) af[k] |= (unsigned char)(1u << (4 + m));
    }
  }
  // neighbour table of the ghost sites (only the clover construction walks from ghost sites)
  {
    std::vector<int> nbg((size_t)8 * (Vg > 0 ? Vg : 1), -1);
    for (long gi = 0; gi < Vg; gi++) for (int d = 0; d < 8; d++) nbg[(long)d * Vg + gi] = neighbour(&gcoord[4 * gi], d);
    d_nbg = dev_upload(nbg);
  }
  d_nb = dev_upload(h_nb);
  h_blkflag = bf;
  d_blkflag = dev_upload(bf);
  d_aggflag = dev_upload(af);
  for (int d = 0; d < 8; d++) d_slice[d] = slices[d].empty() ? nullptr : dev_upload(slices[d]);
  d_lex2nat = dev_upload(lex2nat);
  d_nat2lex = dev_upload(nat2lex);
  std::vector<int> lists[2];
  for (int b = 0; b < nblocks; b++) lists[block_color[b]].push_back(b);
  for (int c = 0; c < 2; c++) { nblk_color[c] = (int)lists[c].size(); d_blocklist[c] = dev_upload(lists[c]); }
  // job list of the block operator (fused coarse SAP kernel): blocks are contiguous site ranges with identical
  // internal structure, block 0 (sites 0..bs-1) defines the list
  if (!block_eo && !last && nblocks > 0) {
    std::vector<int> jobs;
    for (int i = 0; i < bs; i++) { jobs.push_back(0); jobs.push_back(i); jobs.push_back(i); jobs.push_back(0); }
    for (int i = 0; i < bs; i++) for (int m = 0; m < 4; m++)
      if (!((bf[i] >> m) & 1)) { jobs.push_back(1 + m); jobs.push_back(i); jobs.push_back(h_nb[(long)m * V + i]); jobs.push_back(0); }
    nsapjobs = (int)(jobs.size() / 4);
    d_sapjobs = dev_upload(jobs);
  }
  // fused fine SAP kernel (sap_kernel.cu, v2): structure of one 4^4 even-odd block, identical for every block
  if (block_eo && !last && bs == 256 && bs_even == 128) {
    std::vector<unsigned> tab(5 * 256, 0u);
    int slot[4][256], cnt[4] = {0, 0, 0, 0};
    for (int l = 0; l < 256; l++) for (int m = 0; m < 4; m++) slot[m][l] = ((bf[l] >> m) & 1) ? 255 : cnt[m]++;
    for (int m = 0; m < 4; m++) DDA_ASSERT(cnt[m] == 192);
    for (int l = 0; l < 256; l++) {
      unsigned in = (~(unsigned)bf[l]) & 0xFFu, nf = 0, nbk = 0, lf = 0, lb = 0;
      for (int m = 0; m < 4; m++) {
        if (in & (1u << m)) {
          const int n = h_nb[(long)m * V + l];
          DDA_ASSERT(n >= 0 && n < 256 && ((n < 128) != (l < 128)));
          nf |= (unsigned)(n & 127) << (8 * m); lf |= (unsigned)slot[m][l] << (8 * m);
        }
        if (in & (1u << (4 + m))) {
          const int n = h_nb[(long)(4 + m) * V + l];
          DDA_ASSERT(n >= 0 && n < 256 && ((n < 128) != (l < 128)) && slot[m][n] < 192);
          nbk |= (unsigned)(n & 127) << (8 * m); lb |= (unsigned)slot[m][n] << (8 * m);
        }
      }
      tab[5 * l] = in; tab[5 * l + 1] = nf; tab[5 * l + 2] = nbk; tab[5 * l + 3] = lf; tab[5 * l + 4] = lb;
    }
    d_saptab = dev_upload(tab);
    std::vector<int> slotsite(4 * 192, 0);
    for (int l = 0; l < 256; l++) for (int m = 0; m < 4; m++) if (slot[m][l] < 192) slotsite[m * 192 + slot[m][l]] = l;
    d_sapslotsite = dev_upload(slotsite);
  }
  // sites / blocks on the rank boundary (used to overlap the halo exchange with interior work)
  {
    std::vector<char> isb(V, 0);
    for (long k = 0; k < V; k++) for (int d = 0; d < 8; d++) if (h_nb[(long)d * V + k] >= V) isb[k] = 1;
    std::vector<int> bl;
    for (long k = 0; k < V; k++) if (isb[k]) bl.push_back((int)k);
    nbnd = (long)bl.size();
    d_bnd = bl.empty() ? nullptr : dev_upload(bl);
    std::vector<int> li[2], lb[2];
    for (int b = 0; b < nblocks; b++) {
      bool on = false;
      for (int i = 0; i < bs && !on; i++) on = isb[(long)b * bs + i];
      (on ? lb : li)[block_color[b]].push_back(b);
    }
    for (int c = 0; c < 2; c++) {
      nblk_int[c] = (int)li[c].size(); nblk_bnd[c] = (int)lb[c].size();
      d_blocklist_int[c] = dev_upload(li[c]); d_blocklist_bnd[c] = dev_upload(lb[c]);
    }
  }
}

void Geometry::destroy() {
  dev_free(d_nb); dev_free(d_blkflag); dev_free(d_aggflag); dev_free(d_lex2nat); dev_free(d_nat2lex);
  dev_free(d_blocklist[0]); dev_free(d_blocklist[1]); dev_free(d_agg2coarse);
  for (int d = 0; d < 8; d++) { dev_free(d_slice[d]); d_slice[d] = nullptr; }
  dev_free(d_bnd); d_bnd = nullptr;
  dev_free(d_nbg); d_nbg = nullptr;
  dev_free(d_sapjobs); d_sapjobs = nullptr; nsapjobs = 0;
  dev_free(d_saptab); d_saptab = nullptr;
  dev_free(d_sapslotsite); d_sapslotsite = nullptr;
  for (int c = 0; c < 2; c++) { dev_free(d_blocklist_int[c]); dev_free(d_blocklist_bnd[c]); d_blocklist_int[c] = d_blocklist_bnd[c] = nullptr; }
  d_nb = nullptr; d_blkflag = d_aggflag = nullptr; d_lex2nat = d_nat2lex = nullptr;
  d_blocklist[0] = d_blocklist[1] = nullptr; d_agg2coarse = nullptr;
}

}  // namespace dda

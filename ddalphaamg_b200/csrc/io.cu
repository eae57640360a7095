// io.cu -- test-vector files in the reference's native vector format, one file per vector "<name>.NN":
// optional text block "<header>\n ... </header>\n", then the GLOBAL lattice in lexicographic order (t, z, y, x; x fastest),
// 12 complex doubles per site (vector_io, io.c:704-845; header io.c:671-701; multi-file naming setup_generic.c:131-162).
// With "interpolation: 4" and "test vector io file name:" in the parameter file the setup iterations are replaced by
// reading the fine-level test vectors (iterative_PRECISION_setup, setup_generic.c:111-118 -> read_tv_from_file -> re_setup),
// so a hierarchy built on the device can be handed to the CPU reference and vice versa.
// Every rank reads / writes its own x-rows of the file at their global offsets (positional I/O on the shared file of one
// node); rank 0 creates the file.
#include "solver.h"
#include "comm.h"
#include <fcntl.h>
#include <unistd.h>

namespace dda {

static void rank_barrier() {
  if (!g_comm.active()) return;
  static double *buf = nullptr;
  if (!buf) buf = dev_alloc<double>(1);
  dev_zero(buf, sizeof(double));
  comm_allreduce_sum(buf, 1);
  dev_sync();
}

// byte offset of the binary payload (skips an optional <header> block)
static long payload_offset(const char *fn) {
  FILE *f = fopen(fn, "rb");
  if (!f) { fprintf(stderr, "dd_alpha_amg_b200: cannot open %s\n", fn); fatal("test vector io", __FILE__, __LINE__); }
  char line[512];
  long off = 0;
  if (fgets(line, sizeof(line), f) && strcmp(line, "<header>\n") == 0) {
    while (fgets(line, sizeof(line), f)) if (strcmp(line, "</header>\n") == 0) break;
    off = ftell(f);
  }
  fclose(f);
  return off;
}

// mode 0: read file -> host (local lexicographic, 24 doubles per site), 1: write
static void vector_file_io(const Solver &s, const char *fn, std::vector<double> &h, int mode, long off0) {
  const Geometry &g = s.lev[0].geo;
  const int *L = g.L; const int *G = s.p.global_lattice[0];
  const int fd = open(fn, mode ? O_WRONLY : O_RDONLY);
  if (fd < 0) { fprintf(stderr, "dd_alpha_amg_b200: cannot open %s\n", fn); fatal("test vector io", __FILE__, __LINE__); }
  const size_t bar = sizeof(double) * 24 * (size_t)L[3];
  long j = 0;
  for (int t = 0; t < L[0]; t++) for (int z = 0; z < L[1]; z++) for (int y = 0; y < L[2]; y++, j++) {
    const long tg = g.pc[0] * L[0] + t, zg = g.pc[1] * L[1] + z, yg = g.pc[2] * L[2] + y, xg = (long)g.pc[3] * L[3];
    const long site = xg + (long)G[3] * (yg + (long)G[2] * (zg + (long)G[1] * tg));
    const off_t off = off0 + (off_t)site * 24 * sizeof(double);
    char *p = (char *)(h.data() + j * 24 * (long)L[3]);
    const ssize_t rc = mode ? pwrite(fd, p, bar, off) : pread(fd, p, bar, off);
    if (rc != (ssize_t)bar) { fprintf(stderr, "dd_alpha_amg_b200: short %s on %s\n", mode ? "write" : "read", fn); fatal("test vector io", __FILE__, __LINE__); }
  }
  close(fd);
}

void tv_write(Solver &s, const char *base) {
  DDA_ASSERT(s.setup_done && s.nlev > 1);
  Level &L = s.lev[0];
  const long V = L.geo.V;
  std::vector<double> h((size_t)V * 24);
  for (int k = 0; k < L.nv; k++) {
    char fn[1024];
    snprintf(fn, sizeof(fn), "%s.%02d", base, k);
    if (g_comm.rank == 0) {
      FILE *f = fopen(fn, "wb");
      if (!f) { fprintf(stderr, "dd_alpha_amg_b200: cannot create %s\n", fn); fatal("test vector io", __FILE__, __LINE__); }
      fclose(f);
    }
    rank_barrier();
    spinor_to_lex<float>(L.geo, s.lexbuf, L.tv[k], 12);
    d2h(h.data(), s.lexbuf, sizeof(cd) * 12 * V);
    vector_file_io(s, fn, h, 1, 0);
    rank_barrier();
  }
}

void tv_read(Solver &s, const char *base) {
  DDA_ASSERT(s.setup_done && s.nlev > 1);
  Level &L = s.lev[0];
  const long V = L.geo.V;
  std::vector<double> h((size_t)V * 24);
  for (int k = 0; k < L.nv; k++) {
    char fn[1024];
    snprintf(fn, sizeof(fn), "%s.%02d", base, k);
    vector_file_io(s, fn, h, 0, payload_offset(fn));
    h2d(s.lexbuf, h.data(), sizeof(cd) * 12 * V);
    spinor_from_lex<float>(L.geo, L.tv[k], s.lexbuf, 12);
  }
  dev_sync();
}

}  // namespace dda

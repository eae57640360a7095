"""Worker of the world_size-2 tests: one process per rank, lattice split along T.  argv: <lib path> <backend> <out json>.
Backend gloo + the host-emulation library in the GPU-less container; backend nccl + the CUDA library on a multi-GPU box.
Every rank holds its own instance of the oracle (single-rank reference on the GLOBAL lattice) and compares its local part."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ddalphaamg_b200 import DDalphaAMG, read_conf, STAT  # noqa: E402
from ddalphaamg_b200.interface import comm_init, comm_finalize  # noqa: E402
import parity_common as pc  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    lib, backend, outp = sys.argv[1], sys.argv[2], sys.argv[3]
    levels = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(backend, rank=rank, world_size=world)
    comm_init(lib)
    dims, plaq, U = read_conf(os.path.join(ROOT, "tests", "golden", "conf_8x8x8x8b6.0000id3n1"))
    mass = {}
    if len(sys.argv) > 6:
        # synthetic field on a lattice of its own (argv[6] = "T,Z,Y,X"): lets a rank own interior AND boundary 4^4 blocks
        from ddalphaamg_b200 import random_gauge_field
        dims = [int(v) for v in sys.argv[6].split(",")]
        U = random_gauge_field(dims, seed=77, eps=0.3)
        plaq = None
        mass = dict(m0=-0.3, csw=1.0)
    # process grid PT x PZ (argv[5] = "PT,PZ", default: all ranks along T); rank = cT * PZ + cZ (T slowest)
    PT, PZ = (int(v) for v in sys.argv[5].split(",")) if len(sys.argv) > 5 else (world, 1)
    assert PT * PZ == world
    cT, cZ = rank // PZ, rank % PZ
    local = [dims[0] // PT, dims[1] // PZ] + dims[2:]
    if len(sys.argv) > 6:
        block = [4, 4, 4, 4]
        kw = dict(levels=levels, test_vectors=(20, 28)[:max(1, levels - 1)], setup_iter=(2, 1)[:max(1, levels - 1)], restart=20, **mass)
        if levels > 2:
            kw["coarse_block"] = [2, 2, 2, 2]
    elif levels == 2 and (world > 2 or PZ > 1):
        block = [2, 2, 2, 2]      # local T extent 8 / world = 2: one block in T per rank, coarse local T extent 1
        kw = dict(levels=2, test_vectors=(12,), setup_iter=(2,), restart=20)
    elif levels == 2:
        block = [4, 4, 4, 4]
        kw = dict(levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    else:
        block = [2, 2, 2, 2]
        kw = dict(levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=50, coarse_block=[2, 2, 2, 2])
    out = {"rank": rank}
    R = ref.Reference(dims, block, **kw)
    plaq_ref = R.set_conf(U)
    R.setup(kw["setup_iter"][0], nthreads=max(1, (os.cpu_count() or 1) // world) if len(sys.argv) > 6 else 1)
    S = DDalphaAMG(dims, block, lib=lib, local_lattice=local, **kw)
    def part(a4, c_t, c_z):    # this (or another) rank's part of an array whose first four axes are the global T,Z,Y,X
        lt, lz = a4.shape[0] // PT, a4.shape[1] // PZ
        return np.ascontiguousarray(a4[c_t * lt:(c_t + 1) * lt, c_z * lz:(c_z + 1) * lz])

    out["plaq_err"] = abs(S.set_conf(part(U, cT, cZ)) - (plaq if plaq is not None else plaq_ref))

    def loc(v, nc, depth=0):     # local part of a global lexicographic vector of level `depth` with nc entries per site
        g = [R.info(7 + m, depth) for m in range(4)]
        return part(v.reshape(g + [nc]), cT, cZ).reshape(-1)

    rng = np.random.default_rng(4321)
    v = pc.crandom(rng, R.V * 12)
    want = R.dw_double(v)
    out["dw_double"] = pc.rel(loc(want, 12), S.apply_dw(loc(v, 12)))
    out["dw_float"] = pc.rel(loc(want, 12), S.apply_dw(loc(v, 12), "float"))

    # own (distributed) setup: converges like the single-rank reference, residual checked with the reference operator
    S.setup(kw["setup_iter"][0])
    b = np.ones(R.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    xs, ress, sts = S.solve(loc(b, 12))
    parts = [torch.zeros(xs.size, dtype=torch.complex128) for _ in range(world)]
    if backend == "nccl":
        parts = [p.cuda() for p in parts]
        dist.all_gather(parts, torch.from_numpy(xs).cuda())
        parts = [p.cpu() for p in parts]
    else:
        dist.all_gather(parts, torch.from_numpy(xs))
    xg4 = np.zeros(dims + [12], dtype=np.complex128)
    for r_, p_ in enumerate(parts):
        a, b_ = r_ // PZ, r_ % PZ
        xg4[a * local[0]:(a + 1) * local[0], b_ * local[1]:(b_ + 1) * local[1]] = p_.numpy().reshape(local + [12])
    xg = xg4.reshape(-1)
    out["solve_own"] = {"iters": int(sts[0]), "ref_iters": int(str_[0]), "res": float(ress),
                        "res_ref_operator": pc.rel(b, R.dw_double(xg)) if False else float(np.linalg.norm(b - R.dw_double(xg)) / np.linalg.norm(b))}

    # the reference's interpolation imported: operator-by-operator parity of the distributed hierarchy
    for d in range(levels - 1):
        tt = R.translation(d)
        P = R.interpolation(d)
        Vd, nc = R.info(1, d), R.info(2, d)
        nv = P.shape[1]
        Plex = P.reshape(Vd, nc, nv)[tt]                         # global lexicographic
        S.set_interpolation(d, loc(Plex.reshape(-1), nc * nv, d).reshape(-1, nv))
    hier = {}
    for d in range(1, levels):
        Vd, nc = R.info(1, d), R.info(2, d)
        vv = pc.crandom(rng, Vd * nc, np.complex64)
        hier["coarse_apply_d%d" % d] = pc.rel(loc(R.coarse_apply(d, vv), nc, d), S.level_apply(d, loc(vv, nc, d)))
    for d in range(levels - 1):
        Vd, nc = R.info(1, d), R.info(2, d)
        Vc, ncc = R.info(1, d + 1), R.info(2, d + 1)
        vf, vc, phi0 = pc.crandom(rng, Vd * nc, np.complex64), pc.crandom(rng, Vc * ncc, np.complex64), pc.crandom(rng, Vd * nc, np.complex64)
        hier["restrict_d%d" % d] = pc.rel(loc(R.restrict(d, vf), ncc, d + 1), S.restrict(d, loc(vf, nc, d)))
        hier["interpolate_d%d" % d] = pc.rel(loc(R.interpolate(d, vc), nc, d), S.interpolate(d, loc(vc, ncc, d + 1)))
        hier["smoother_d%d" % d] = pc.rel(loc(R.smoother(d, vf, 2, phi0), nc, d), S.smoother(d, loc(vf, nc, d), 2, loc(phi0, nc, d)))
        hier["vcycle_d%d" % d] = pc.rel(loc(R.vcycle(d, vf), nc, d), S.vcycle(d, loc(vf, nc, d)))
    Vl, ncl = R.info(1, levels - 1), R.info(2, levels - 1)
    vv = pc.crandom(rng, Vl * ncl, np.complex64)
    xr_c, itr = R.coarsest_solve(vv)
    hier["coarsest_solve"] = pc.rel(loc(xr_c, ncl, levels - 1), S.coarsest_solve(loc(vv, ncl, levels - 1)))
    hier["coarsest_iters"] = [int(itr), int(S.stat(STAT.COARSE_ITER))]
    w = pc.crandom(rng, R.V * 12)
    hier["preconditioner"] = pc.rel(loc(R.preconditioner(w), 12), S.preconditioner(loc(w, 12)))
    out["hierarchy"] = hier
    xs, ress, sts = S.solve(loc(b, 12))
    out["solve_imported"] = {"iters": int(sts[0]), "ref_iters": int(str_[0]), "res": float(ress)}
    S.free()
    R.free()
    comm_finalize(lib)
    with open(outp, "w") as f:
        json.dump(out, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

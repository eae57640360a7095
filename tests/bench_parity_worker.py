"""Worker of test_bench_parity.py::test_parity_two_ranks: bench.py's parity check with the lattice split along T over two gloo
ranks (emulation build): per-rank files, slab assembly across rank boundaries, sums over ranks.  argv: lib, workdir, out json."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field  # noqa: E402
from ddalphaamg_b200.interface import comm_init, comm_finalize  # noqa: E402


def main():
    lib, workdir, outp = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm_init(lib)
    lat = [16, 8, 8, 8]
    lt = lat[0] // world
    U = random_gauge_field(lat, seed=5, eps=0.3, t_range=(rank * lt, (rank + 1) * lt))
    S = DDalphaAMG(lat, [4, 4, 4, 4], lib=lib, local_lattice=[lt] + lat[1:], levels=2, test_vectors=(12,), setup_iter=(2,),
                   restart=10, m0=-0.3, mixed_precision=2)
    S.set_conf(U)
    np.save(os.path.join(workdir, "U_%d.npy" % rank), U)
    S.setup(2)
    b = np.ones(S.V * 12, dtype=np.complex128)
    x, res, st = S.solve(b)
    par = bench.parity_check(S, dict(m0=-0.3), lat, (world, 1), rank, world, workdir, x, dist)
    S.free()
    comm_finalize(lib)
    with open(outp, "w") as f:
        json.dump({"res": float(res), "iters": int(st[0]), "parity": par}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

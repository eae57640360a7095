"""bench.py's workload-size parity check (reference operator on T slabs with halo slices) exercised on CPU: emulation
build of the library on a 16 x 8^3 lattice, two slabs of 8 interior + 2 x 2 halo slices (the halo wraps around the
global lattice), reference in sub-processes exactly as in the benchmark."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field  # noqa: E402


def test_parity_slabs_match_reference_operator(emu_lib, oracle_ref, tmp_path):
    lat = [16, 8, 8, 8]
    w = dict(m0=-0.3)
    U = random_gauge_field(lat, seed=5, eps=0.3)
    S = DDalphaAMG(lat, [4, 4, 4, 4], lib=emu_lib, levels=2, test_vectors=(12,), setup_iter=(2,), restart=10, m0=-0.3, mixed_precision=2)
    try:
        S.set_conf(U)
        np.save(os.path.join(str(tmp_path), "U_0.npy"), U)
        S.setup(2)
        b = np.ones(S.V * 12, dtype=np.complex128)
        x, res, st = S.solve(b)
        assert st[0] > 0 and res < 1e-10
        par = bench.parity_check(S, w, lat, (1, 1), 0, 1, str(tmp_path), x, None)
        assert par.get("ok") is True, par
        assert par["coverage"].startswith("16 of 16"), par
        # a corrupted solution must be caught by the slab residual
        x2 = x.copy()
        x2[5 * 8 * 8 * 8 * 12 + 7] += 1e-6
        par2 = bench.parity_check(S, w, lat, (1, 1), 0, 1, str(tmp_path), x2, None)
        assert par2.get("ok") is False and par2["residual_ref_operator"] > 1.5e-10, par2
    finally:
        S.free()


def test_parity_two_ranks(emu_lib, oracle_ref, tmp_path):
    """The same check with the lattice split along T over two gloo ranks: every rank writes its part, slab j is evaluated by
    rank j mod 2 from files of both ranks (the halo slices of a slab lie on the other rank), the sums are all-reduced."""
    import json
    import socket
    import subprocess
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs, outs = [], []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS=str(max(1, (os.cpu_count() or 4) // 2)))
        o = str(tmp_path / ("rank%d.json" % r))
        outs.append(o)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "bench_parity_worker.py"), emu_lib, str(tmp_path), o],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=900)[0] for p in procs]
    for p, lg in zip(procs, logs):
        assert p.returncode == 0, lg[-3000:]
    res = [json.load(open(o)) for o in outs]
    for r in res:
        assert r["iters"] > 0 and r["res"] < 1e-10
        assert r["parity"].get("ok") is True and r["parity"]["coverage"].startswith("16 of 16"), r
    assert res[0]["parity"]["residual_ref_operator"] == res[1]["parity"]["residual_ref_operator"]

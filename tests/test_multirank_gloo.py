"""world_size-2 test of the multi-GPU path's host logic on CPU: lattice split along T over two processes (gloo),
host-emulation build of the library, halo exchanges and global sums through torch.distributed; every rank checks its
local part against the single-rank oracle (the operator is decomposition independent)."""
import json
import os
import socket
import subprocess
import sys

import pytest

import parity_common as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_ranks(lib, backend, levels, tmp_path, world=2, timeout=900, grid=None, lattice=None, extra_env=None):
    port = _free_port()
    procs, outs = [], []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS=str(max(1, (os.cpu_count() or 4) // world)))
        env.update(extra_env or {})
        o = str(tmp_path / ("rank%d.json" % r))
        outs.append(o)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multirank_worker.py"), lib, backend, o, str(levels)]
                                      + ([grid or "%d,1" % world] if (grid or lattice) else []) + ([lattice] if lattice else []),
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=timeout)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, lg in zip(procs, logs):
        assert p.returncode == 0, lg[-4000:]
    return [json.load(open(o)) for o in outs]


def check(results, imported_tol_iters=1):
    for out in results:
        assert out["plaq_err"] < 1e-12, out
        assert out["dw_double"] <= pc.TOL_DOUBLE and out["dw_float"] <= pc.TOL_FLOAT, out
        so = out["solve_own"]
        assert so["iters"] > 0 and so["res"] < 1e-10 and so["res_ref_operator"] < 1.5e-10, out
        assert abs(so["iters"] - so["ref_iters"]) <= 5, out
        pc.assert_hierarchy(out["hierarchy"])
        si = out["solve_imported"]
        assert si["res"] < 1e-10 and abs(si["iters"] - si["ref_iters"]) <= imported_tol_iters, out


@pytest.mark.parametrize("levels", [2, 3])
def test_two_ranks_split_T(emu_lib, oracle_ref, tmp_path, levels):
    check(run_ranks(emu_lib, "gloo", levels, tmp_path))


def test_four_ranks_split_T(emu_lib, oracle_ref, tmp_path):
    """Four ranks: the +T and -T neighbours of a rank are different processes (with two ranks they coincide)."""
    check(run_ranks(emu_lib, "gloo", 2, tmp_path, world=4), imported_tol_iters=1)


def test_two_ranks_split_Z(emu_lib, oracle_ref, tmp_path):
    """Partition along Z only (process grid 1 x 2)."""
    check(run_ranks(emu_lib, "gloo", 2, tmp_path, world=2, grid="1,2"))


def test_four_ranks_split_T_and_Z(emu_lib, oracle_ref, tmp_path):
    """Process grid 2 x 2 in T x Z: the clover term needs the corner sites x +- T +- Z of the extended ghost slabs."""
    check(run_ranks(emu_lib, "gloo", 2, tmp_path, world=4, grid="2,2"))


@pytest.mark.parametrize("dirs,levels", [("T", 2), ("TZ", 3)])
def test_forced_split_single_rank(emu_lib, oracle_ref, tmp_path, dirs, levels):
    """DDA_FORCE_SPLIT: ONE rank that carries ghost slabs and is its own periodic neighbour -- the partitioned code paths
    (pack kernel, ghost neighbour tables, interior / boundary lists, coarse ghost hops) without a second process."""
    check(run_ranks(emu_lib, "gloo", levels, tmp_path, world=1, extra_env={"DDA_FORCE_SPLIT": dirs}))


def test_two_ranks_partitioned_coarsest(emu_lib, oracle_ref, tmp_path):
    """Replication of the coarsest level switched off: the device-resident GMRES runs on the partitioned coarsest lattice
    (halo exchanges inside the Schur complement, all-reduces on the device-side Hessenberg buffers)."""
    check(run_ranks(emu_lib, "gloo", 3, tmp_path, extra_env={"DDA_COARSEST_REPLICATE_MAX": "0"}))

"""The oracle (unmodified reference, oracle/_ref) against the known answers recorded from the reference's own runs
 and against the committed golden vectors."""
import os

import numpy as np
import pytest

from conftest import CONF4, CONF8, GOLDEN
from ddalphaamg_b200 import read_conf


def test_plaquette_and_operator_known_answer(oracle_ref):
    dims, plaq, U = read_conf(CONF8)
    assert dims == [8, 8, 8, 8]
    R = oracle_ref.Reference(dims, [4, 4, 4, 4], levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    try:
        p = R.set_conf(U)
        assert abs(p - 1.7772950976130) < 1e-12 and abs(p - plaq) < 1e-12     # file header, io.c:504
        ones = np.ones(R.V * 12, dtype=np.complex128)
        out = R.dw_double(ones)
        assert abs(np.vdot(out, out).real - 8.042666972709764e+05) < 1e-6
        assert abs(out[0] - (0.2091708830936854 + 0.4800419460605345j)) < 1e-13
        assert abs(R.D().reshape(-1)[0] - (0.3316105802849674 - 0.008353665220814323j)) < 1e-15
        assert abs(R.clover()[0, 0] - 3.764387515415905) < 1e-13
    finally:
        R.free()


def test_end_to_end_known_answer_sample_ini(oracle_ref):
    # the reference's own sample.ini on conf/8x8x8x8b6.0000id3n1 (3 levels, blocks 2^4, 28 test vectors, 4 setup
    # iterations, restart 50, mixed precision 1, m0 -0.5, csw 1, anti-periodic, rhs ones, tol 1e-10):
    # 11 outer iterations, ||r||/||b|| = 1.399e-11 (1.399044e-11 with 1 thread, 1.399092e-11 with 8)
    dims, plaq, U = read_conf(CONF8)
    R = oracle_ref.Reference(dims, [2, 2, 2, 2], levels=3, test_vectors=(28, 28), setup_iter=(4, 3), restart=50,
                             max_restart=20, coarse_iter=100, coarse_restart=5, mixed_precision=1)
    try:
        R.set_conf(U)
        R.setup(4, nthreads=min(8, os.cpu_count() or 1))
        x, res, st = R.solve(np.ones(R.V * 12, dtype=np.complex128))
        assert st[0] == 11 and abs(res - 1.399e-11) < 1e-13
    finally:
        R.free()


def test_golden_vectors(oracle_ref):
    g = np.load(os.path.join(GOLDEN, "golden_4x4x4x4.npz"))
    dims, plaq, U = read_conf(CONF4)
    R = oracle_ref.Reference(dims, [2, 2, 2, 2], levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    try:
        R.set_conf(U)
        assert np.linalg.norm(R.dw_double(g["dw_in"]) - g["dw_out"]) / np.linalg.norm(g["dw_out"]) < 1e-14
    finally:
        R.free()

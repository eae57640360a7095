"""Parity tests proper: the CUDA path (libdd_alpha_amg.so, through the C ABI) against the oracle = unmodified reference
(oracle/_ref, built in the container and shipped to the GPU box), against the committed golden vectors, and at
BASELINE.json's full sizes through size-independent properties."""
import os

import numpy as np
import pytest

from conftest import CONF4, CONF8, GOLDEN
from ddalphaamg_b200 import DDalphaAMG, read_conf, random_gauge_field, INFO, STAT, OPT
import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair8(oracle_ref, cuda_lib):
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    R = oracle_ref.Reference(dims, [4, 4, 4, 4], **kw)
    R.set_conf(U)
    R.setup(2)
    S = DDalphaAMG(dims, [4, 4, 4, 4], lib=cuda_lib, **kw)
    assert not S.emulated
    assert abs(S.set_conf(U) - plaq) < 1e-12
    S.setup(2)
    yield R, S
    S.free()
    R.free()


def test_fine_operator_vs_reference(pair8):
    R, S = pair8
    for fast in (1, 0):
        S.set_option(OPT.USE_FAST, fast)
        pc.check_fine_operator(R, S)
    S.set_option(OPT.USE_FAST, 1)


def test_own_setup_converges_like_reference(pair8):
    R, S = pair8
    b = np.ones(S.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    launches0 = S.stat(STAT.LAUNCHES)
    xs, ress, sts = S.solve(b)
    assert S.stat(STAT.LAUNCHES) > launches0          # our kernels did the work
    assert sts[0] > 0 and ress < 1e-10
    assert abs(int(sts[0]) - int(str_[0])) <= 4
    assert pc.rel(xr, xs) < 1e-8
    assert pc.rel(b, R.dw_double(xs)) < 1.5e-10       # residual checked with the reference's operator


def test_hierarchy_vs_reference_two_level(pair8):
    R, S = pair8
    for fast in (0, 1):
        S.set_option(OPT.USE_FAST, fast)
        out = pc.check_hierarchy(R, S, 2)
        pc.assert_hierarchy(out)
    b = np.ones(S.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    xs, ress, sts = S.solve(b)
    assert abs(int(sts[0]) - int(str_[0])) <= 1 and ress < 1e-10


def test_golden_vectors_4x4x4x4(cuda_lib):
    g = np.load(os.path.join(GOLDEN, "golden_4x4x4x4.npz"))
    dims, plaq, U = read_conf(CONF4)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=cuda_lib, levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    try:
        assert abs(S.set_conf(U) - plaq) < 1e-12
        assert pc.rel(g["dw_out"], S.apply_dw(g["dw_in"])) <= pc.TOL_DOUBLE
        assert pc.rel(g["dw_out"], S.apply_dw(g["dw_in"], "float")) <= pc.TOL_FLOAT
        S.setup(0)
        S.set_interpolation(0, g["P"])
        assert pc.rel(g["coarse_out"], S.level_apply(1, g["coarse_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["restrict_out"], S.restrict(0, g["restrict_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["interpolate_out"], S.interpolate(0, g["interpolate_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["smoother_out"], S.smoother(0, g["smoother_eta"], 2, g["smoother_phi0"])) <= pc.TOL_FLOAT
        assert pc.rel(g["coarsest_out"], S.coarsest_solve(g["coarsest_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["prec_out"], S.preconditioner(g["prec_in"])) <= pc.TOL_FLOAT
        x, res, st = S.solve(np.ones(S.V * 12, dtype=np.complex128))
        assert abs(int(st[0]) - int(g["solve_iters"][0])) <= 1 and res < 1e-10
    finally:
        S.free()


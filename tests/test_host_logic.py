"""Host logic (cycle control flow, Krylov drivers, setup, layouts, C API plumbing) checked in the GPU-less container.

These tests load tests/_emu/libdda_emu.so: the SAME sources as the CUDA library compiled with -DDDA_HOST_EMU, where
every generic kernel body runs as a host loop.  It is test infrastructure only (the package never loads it, and the
hand-tuned sm_100a kernels are not part of it); the parity tests proper are the -m gpu tests."""
import os

import numpy as np
import pytest

from conftest import CONF4, CONF8, GOLDEN
from ddalphaamg_b200 import DDalphaAMG, read_conf, random_gauge_field, INFO, STAT
import parity_common as pc


@pytest.fixture(scope="module")
def pair8(oracle_ref, emu_lib):
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    R = oracle_ref.Reference(dims, [4, 4, 4, 4], **kw)
    R.set_conf(U)
    R.setup(2)
    S = DDalphaAMG(dims, [4, 4, 4, 4], lib=emu_lib, **kw)
    assert S.emulated
    assert abs(S.set_conf(U) - plaq) < 1e-12
    S.setup(2)
    yield R, S
    S.free()
    R.free()


def test_fine_operator_vs_reference(pair8):
    R, S = pair8
    pc.check_fine_operator(R, S)


def test_own_setup_converges_like_reference(pair8):
    R, S = pair8
    b = np.ones(S.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    xs, ress, sts = S.solve(b)
    assert sts[0] > 0 and ress < 1e-10
    assert abs(int(sts[0]) - int(str_[0])) <= 4          # comparable outer iteration count
    assert pc.rel(xr, xs) < 1e-8                          # both solve D x = b to 1e-10
    assert pc.rel(b, S.apply_dw(xs)) < 1.5e-10


def test_hierarchy_vs_reference_two_level(pair8):
    R, S = pair8
    out = pc.check_hierarchy(R, S, 2)
    pc.assert_hierarchy(out)
    # with the reference's own interpolation the whole solve follows the reference iteration by iteration
    b = np.ones(S.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    xs, ress, sts = S.solve(b)
    assert int(sts[0]) == int(str_[0]) and abs(int(sts[1]) - int(str_[1])) <= 2
    assert abs(ress - resr) / resr < 1e-3


def test_golden_vectors_4x4x4x4(emu_lib):
    """Committed golden vectors generated from the reference (tests/golden/make_golden.py); no oracle needed."""
    g = np.load(os.path.join(GOLDEN, "golden_4x4x4x4.npz"))
    dims, plaq, U = read_conf(CONF4)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    try:
        assert abs(S.set_conf(U) - plaq) < 1e-12
        assert pc.rel(g["dw_out"], S.apply_dw(g["dw_in"])) <= pc.TOL_DOUBLE
        assert pc.rel(g["dw_out"], S.apply_dw(g["dw_in"], "float")) <= pc.TOL_FLOAT
        S.setup(0)
        S.set_interpolation(0, g["P"])
        assert pc.rel(g["P"], S.get_interpolation(0)) == 0.0
        assert pc.rel(g["coarse_out"], S.level_apply(1, g["coarse_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["restrict_out"], S.restrict(0, g["restrict_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["interpolate_out"], S.interpolate(0, g["interpolate_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["smoother_out"], S.smoother(0, g["smoother_eta"], 2, g["smoother_phi0"])) <= pc.TOL_FLOAT
        assert pc.rel(g["coarsest_out"], S.coarsest_solve(g["coarsest_in"])) <= pc.TOL_FLOAT
        assert pc.rel(g["prec_out"], S.preconditioner(g["prec_in"])) <= pc.TOL_FLOAT
        x, res, st = S.solve(np.ones(S.V * 12, dtype=np.complex128))
        assert int(st[0]) == int(g["solve_iters"][0]) and res < 1e-10
    finally:
        S.free()


def test_edge_cases(emu_lib):
    """method 0 (no preconditioner), csw = 0, periodic boundary conditions, non-convergence status, mass shift."""
    dims, plaq, U = read_conf(CONF4, anti_pbc=False)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, bc=1, csw=0.0, m0=0.2, levels=2, test_vectors=(8,), setup_iter=(1,),
                   restart=20, max_restart=50, method=0)
    try:
        S.set_conf(U)
        S.setup(0)
        b = np.zeros(S.V * 12, dtype=np.complex128)
        b[0] = 1.0
        x, res, st = S.solve(b, tol=1e-9)
        assert st[0] > 0 and res < 1e-9 and pc.rel(b, S.apply_dw(x)) < 2e-9
        # gamma5-hermiticity of D_W: <g5 D x, y> = <x, g5 D y>
        rng = np.random.default_rng(5)
        u, v = pc.crandom(rng, S.V * 12), pc.crandom(rng, S.V * 12)
        g5 = np.tile(np.repeat([-1.0, 1.0], 6), S.V)
        assert abs(np.vdot(g5 * S.apply_dw(u), v) - np.vdot(u, g5 * S.apply_dw(v))) < 1e-10
        # too few iterations -> status[0] = -1 (dd_alpha_amg.c:391-392)
        S2 = None
    finally:
        S.free()
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, bc=1, csw=1.0, m0=0.2, levels=2, test_vectors=(8,), setup_iter=(1,),
                   restart=2, max_restart=1, method=0)
    try:
        S.set_conf(U)
        S.setup(0)
        x, res, st = S.solve(np.ones(S.V * 12, dtype=np.complex128), tol=1e-12)
        assert st[0] == -1 and res > 1e-12
    finally:
        S.free()


def test_synthetic_field_generator_is_su3():
    U = random_gauge_field([4, 4, 4, 4], seed=3, eps=0.4, anti_pbc=False)
    M = (U[..., 0] + 1j * U[..., 1]).reshape(-1, 3, 3)
    assert np.abs(M @ np.conj(np.swapaxes(M, 1, 2)) - np.eye(3)).max() < 1e-13
    assert np.abs(np.linalg.det(M) - 1.0).max() < 1e-13


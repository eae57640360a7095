"""BASELINE.json config[1] size (16^3 x 32, 2 levels) on the GPU: D_W against the reference's operator on the same
synthetic gauge field, size-independent properties, and a full solve whose residual is checked with the reference's
operator."""
import numpy as np
import pytest

from ddalphaamg_b200 import DDalphaAMG, random_gauge_field, STAT, OPT
import parity_common as pc

pytestmark = pytest.mark.gpu

LATTICE = [32, 16, 16, 16]


def test_fullsize_16x16x16x32(oracle_ref, cuda_lib):
    U = random_gauge_field(LATTICE, seed=20261018, eps=0.3)
    kw = dict(levels=2, test_vectors=(20,), setup_iter=(3,), restart=10, m0=-0.1)
    R = oracle_ref.Reference(LATTICE, [4, 4, 4, 4], **kw)
    S = DDalphaAMG(LATTICE, [4, 4, 4, 4], lib=cuda_lib, **kw)
    try:
        pr, ps = R.set_conf(U), S.set_conf(U)
        assert abs(pr - ps) < 1e-12
        rng = np.random.default_rng(99)
        n = S.V * 12
        u, v = pc.crandom(rng, n), pc.crandom(rng, n)
        Du = S.apply_dw(u)
        assert pc.rel(R.dw_double(u), Du) <= pc.TOL_DOUBLE
        assert pc.rel(R.dw_double(u), S.apply_dw(u, "float")) <= pc.TOL_FLOAT
        # linearity and gamma5-hermiticity
        a = 0.3 - 1.7j
        assert pc.rel(Du + a * S.apply_dw(v), S.apply_dw(u + a * v)) < 1e-13
        g5 = np.tile(np.repeat([-1.0, 1.0], 6), S.V)
        lhs, rhs = np.vdot(g5 * Du, v), np.vdot(u, g5 * S.apply_dw(v))
        assert abs(lhs - rhs) / abs(lhs) < 1e-12
        # setup + solve on the device; residual verified with the reference's D_W
        S.setup(3)
        b = np.ones(n, dtype=np.complex128)
        x, res, st = S.solve(b)
        assert st[0] > 0 and res < 1e-10
        assert pc.rel(b, R.dw_double(x)) < 1.5e-10
        # Galerkin identity on the device hierarchy: P^H D P v_c = D_c v_c, and P^H P = 1
        Vc, nc = S.level_shape(1)
        vc = pc.crandom(rng, Vc * nc, np.complex64)
        Pv = S.interpolate(0, vc)
        assert pc.rel(vc, S.restrict(0, Pv)) < 1e-5
        assert pc.rel(S.level_apply(1, vc), S.restrict(0, S.level_apply(0, Pv))) < 1e-5
    finally:
        S.free()
        R.free()


def test_config2_32x32x32x64_three_level(oracle_ref, cuda_lib):
    """BASELINE.json configs[2] size (32^3 x 64, 3 levels): fine operator against the reference on the same synthetic
    field, size-independent properties of the device hierarchy (P^H P = 1, Galerkin identity on both coarse levels,
    gamma5-hermiticity), and a solve to 1e-10 whose residual is checked with the reference's operator."""
    lat = [64, 32, 32, 32]
    U = random_gauge_field(lat, seed=20261018, eps=0.3)
    kw = dict(levels=3, test_vectors=(20, 28), setup_iter=(2, 2), restart=10, m0=-0.1, coarse_block=[2, 2, 2, 2])
    # the reference only supplies its fine operator here (2-level parameters keep its own setup cheap; no setup is run)
    R = oracle_ref.Reference(lat, [4, 4, 4, 4], levels=2, test_vectors=(20,), setup_iter=(1,), restart=10, m0=-0.1)
    S = DDalphaAMG(lat, [4, 4, 4, 4], lib=cuda_lib, **kw)
    try:
        pr, ps = R.set_conf(U), S.set_conf(U)
        assert abs(pr - ps) < 1e-12
        rng = np.random.default_rng(5)
        n = S.V * 12
        u, v = pc.crandom(rng, n), pc.crandom(rng, n)
        want = R.dw_double(u)
        Du = S.apply_dw(u)
        assert pc.rel(want, Du) <= pc.TOL_DOUBLE
        assert pc.rel(want, S.apply_dw(u, "float")) <= pc.TOL_FLOAT
        g5 = np.tile(np.repeat([-1.0, 1.0], 6), S.V)
        lhs, rhs = np.vdot(g5 * Du, v), np.vdot(u, g5 * S.apply_dw(v))
        assert abs(lhs - rhs) / abs(lhs) < 1e-12
        S.setup(2)
        for d in (0, 1):
            Vc, nc = S.level_shape(d + 1)
            vc = pc.crandom(rng, Vc * nc, np.complex64)
            Pv = S.interpolate(d, vc)
            assert pc.rel(vc, S.restrict(d, Pv)) < 1e-5                                         # P^H P = 1
            assert pc.rel(S.level_apply(d + 1, vc), S.restrict(d, S.level_apply(d, Pv))) < 2e-5   # P^H D P = D_c
        b = np.ones(n, dtype=np.complex128)
        x, res, st = S.solve(b)
        assert st[0] > 0 and res < 1e-10
        assert pc.rel(b, R.dw_double(x)) < 1.5e-10
    finally:
        S.free()
        R.free()

"""Shared parity checks: the same assertions run against the CUDA library on the GPU box (-m gpu) and against the
host-emulation build of the same host logic in the GPU-less container (host-logic tests).  The checker is always the
oracle = the unmodified reference built into oracle/_ref (tests only)."""
import numpy as np

from ddalphaamg_b200 import DDalphaAMG, read_conf, INFO, STAT

TOL_DOUBLE = 1e-12   # relative L2 per operator apply (BASELINE.json north_star)
TOL_FLOAT = 1e-5


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(a))


def crandom(rng, n, dtype=np.complex128):
    return (rng.uniform(-0.5, 0.5, n) + 1j * rng.uniform(-0.5, 0.5, n)).astype(dtype)


def import_interpolation(R, S, depth):
    """Copies the reference's prolongator of level `depth` into the device hierarchy (lexicographic site order)."""
    tt = R.translation(depth)
    P = R.interpolation(depth)
    V, nc = S.level_shape(depth)
    nv = P.shape[1]
    S.set_interpolation(depth, P.reshape(V, nc, nv)[tt].reshape(V * nc, nv))


def check_fine_operator(R, S):
    rng = np.random.default_rng(12345)
    v = crandom(rng, S.V * 12)
    ref = R.dw_double(v)
    assert rel(ref, S.apply_dw(v, "double")) <= TOL_DOUBLE
    assert rel(ref, S.apply_dw(v, "float")) <= TOL_FLOAT
    D, cl = S.operator_arrays()
    assert np.abs(D - R.D()).max() == 0.0
    assert np.abs(cl - R.clover()).max() <= 1e-13
    ones = np.ones(S.V * 12, dtype=np.complex128)
    assert rel(R.dw_double(ones), S.apply_dw(ones)) <= TOL_DOUBLE


def check_hierarchy(R, S, levels, rng=None):
    """Operator-by-operator parity with the reference's own hierarchy imported into the device solver."""
    rng = rng or np.random.default_rng(777)
    for d in range(levels - 1):
        import_interpolation(R, S, d)
    out = {}
    for d in range(1, levels):
        V, nc = S.level_shape(d)
        v = crandom(rng, V * nc, np.complex64)
        out["coarse_apply_d%d" % d] = rel(R.coarse_apply(d, v), S.level_apply(d, v))
    for d in range(levels - 1):
        V, nc = S.level_shape(d)
        Vc, ncc = S.level_shape(d + 1)
        vf, vc, phi0 = crandom(rng, V * nc, np.complex64), crandom(rng, Vc * ncc, np.complex64), crandom(rng, V * nc, np.complex64)
        out["restrict_d%d" % d] = rel(R.restrict(d, vf), S.restrict(d, vf))
        out["interpolate_d%d" % d] = rel(R.interpolate(d, vc), S.interpolate(d, vc))
        out["smoother_d%d" % d] = rel(R.smoother(d, vf, 2, phi0), S.smoother(d, vf, 2, phi0))
        out["vcycle_d%d" % d] = rel(R.vcycle(d, vf), S.vcycle(d, vf))
    V, nc = S.level_shape(levels - 1)
    v = crandom(rng, V * nc, np.complex64)
    xr, itr = R.coarsest_solve(v)
    xs = S.coarsest_solve(v)
    out["coarsest_solve"] = rel(xr, xs)
    out["coarsest_iters"] = (itr, int(S.stat(STAT.COARSE_ITER)))
    w = crandom(rng, S.V * 12)
    out["preconditioner"] = rel(R.preconditioner(w), S.preconditioner(w))
    return out


def assert_hierarchy(out):
    for k, v in out.items():
        if k == "coarsest_iters":
            assert abs(v[0] - v[1]) <= 1, out
        else:
            assert v <= TOL_FLOAT, out


def check_mass_shift_and_clover_scaling(R, S, m_new):
    """a21: shift_update (dirac.c:669-691) through dd_alpha_amg_update_parameters, and the per-parity clover scaling of
    dd_alpha_amg_wilson_solve (scale_clover dirac.c:646-667, dd_alpha_amg.c:354-373), both against the reference with
    the reference's prolongator imported.  R and S are set up 2-level pairs at the same mass."""
    rng = np.random.default_rng(2024)
    import_interpolation(R, S, 0)
    R.shift_mass(m_new)
    S.update_parameters(m_new)
    out = {}
    v = crandom(rng, S.V * 12)
    want = R.dw_double(v)
    out["dw_double"] = rel(want, S.apply_dw(v))
    out["dw_float"] = rel(want, S.apply_dw(v, "float"))
    V1, n1 = S.level_shape(1)
    vc = crandom(rng, V1 * n1, np.complex64)
    out["coarse_apply"] = rel(R.coarse_apply(1, vc), S.level_apply(1, vc))
    xr, itr = R.coarsest_solve(vc)
    out["coarsest_solve"] = rel(xr, S.coarsest_solve(vc))
    w = crandom(rng, S.V * 12)
    out["preconditioner"] = rel(R.preconditioner(w), S.preconditioner(w))
    b = np.ones(S.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    xs, ress, sts = S.solve(b)
    out["solve_shifted"] = (int(str_[0]), int(sts[0]), float(ress), rel(xr, xs))
    se, so = 1.1, 0.9
    xr, resr, str_ = R.solve(b, scale_even=se, scale_odd=so)
    xs, ress, sts = S.solve(b, scale_even=se, scale_odd=so)
    out["solve_scaled"] = (int(str_[0]), int(sts[0]), float(ress), rel(xr, xs))
    out["restored"] = rel(R.dw_double(v), S.apply_dw(v))
    assert out["dw_double"] <= TOL_DOUBLE and out["restored"] <= TOL_DOUBLE and out["dw_float"] <= TOL_FLOAT, out
    for k in ("coarse_apply", "coarsest_solve", "preconditioner"):
        assert out[k] <= TOL_FLOAT, out
    for k in ("solve_shifted", "solve_scaled"):
        ir, is_, res, dx = out[k]
        assert is_ > 0 and abs(ir - is_) <= 1 and res < 1e-10 and dx < 1e-8, out
    return out


def check_near_critical_solve(R, S):
    """Near-critical mass: the coarsest-level GMRES needs more than 64 Arnoldi steps per restart cycle (the regime that
    dominates a near-physical-mass solve).  Reference prolongator imported; outer iterations +-1, coarsest iterations
    within 5 %, residual verified with the reference operator."""
    import_interpolation(R, S, 0)
    b = np.ones(S.V * 12, dtype=np.complex128)
    xr, resr, str_ = R.solve(b)
    xs, ress, sts = S.solve(b)
    out = {"ref": [int(v) for v in str_], "mine": [int(v) for v in sts], "res": float(ress), "res_ref_op": rel(b, R.dw_double(xs))}
    assert out["mine"][0] > 0 and abs(out["ref"][0] - out["mine"][0]) <= 1, out
    assert out["ref"][1] / out["ref"][0] > 64, out
    assert abs(out["ref"][1] - out["mine"][1]) <= 0.05 * out["ref"][1], out
    assert out["res"] < 1e-10 and out["res_ref_op"] < 1.5e-10, out
    return out

"""Behaviour of the reference-facing C API beyond the plain solve (host logic, emulation build, no oracle instance):
struct-route initialisation, mass update, clover scaling by site parity, setup update, raw gauge pointer access."""
import numpy as np

from conftest import CONF4
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


def clover_apply(cl, v):
    """y = C v with the reference's packed clover array [site][42] (dirac.c:386-398, dirac_generic.h:723-799)."""
    V = cl.shape[0]
    v = v.reshape(V, 12)
    y = cl[:, :12] * v
    for b in range(2):
        m = 12 + 15 * b
        for i in range(6):
            for j in range(i + 1, 6):
                c = cl[:, m]
                y[:, 6 * b + i] += c * v[:, 6 * b + j]
                y[:, 6 * b + j] += np.conj(c) * v[:, 6 * b + i]
                m += 1
    return y.reshape(-1)


def site_parity(dims):
    t, z, y, x = np.meshgrid(*[np.arange(d) for d in dims], indexing="ij")
    return ((t + z + y + x) & 1).reshape(-1)


def make(emu_lib, **kw):
    dims, plaq, U = read_conf(CONF4)
    base = dict(levels=2, test_vectors=(12,), setup_iter=(2,), restart=20, max_restart=20, m0=-0.2)
    base.update(kw)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, **base)
    assert abs(S.set_conf(U) - plaq) < 1e-12
    return dims, U, S


def test_mass_update_shifts_every_level(emu_lib):
    dims, U, S = make(emu_lib)
    try:
        rng = np.random.default_rng(2)
        v = pc.crandom(rng, S.V * 12)
        d0 = S.apply_dw(v)
        S.setup(2)
        vc = pc.crandom(rng, np.prod(S.level_shape(1)), np.complex64)
        c0 = S.level_apply(1, vc)
        S.update_parameters(-0.05)                       # solver mass -0.2 -> -0.05
        assert pc.rel(d0 + 0.15 * v, S.apply_dw(v)) < 1e-14
        assert pc.rel(c0 + np.complex64(0.15) * vc, S.level_apply(1, vc)) < 1e-5    # P^H 1 P = 1 on the coarse level
        b = np.ones(S.V * 12, dtype=np.complex128)
        x, res, st = S.solve(b)
        assert st[0] > 0 and res < 1e-10 and pc.rel(b, S.apply_dw(x)) < 1.5e-10
    finally:
        S.free()


def test_clover_scaling_by_parity_in_wilson_solve(emu_lib):
    """scale_even / scale_odd of dd_alpha_amg_wilson_solve (dd_alpha_amg.c:354-373, scale_clover dirac.c:646-667): the
    clover term (diagonal shift included) is multiplied per site parity for this solve and restored afterwards."""
    dims, U, S = make(emu_lib)
    try:
        S.setup(2)
        rng = np.random.default_rng(3)
        v = pc.crandom(rng, S.V * 12)
        before = S.apply_dw(v)
        D, cl = S.operator_arrays()
        se, so = 1.1, 0.9
        b = np.ones(S.V * 12, dtype=np.complex128)
        x, res, st = S.solve(b, scale_even=se, scale_odd=so)
        assert st[0] > 0 and res < 1e-10
        fac = np.where(site_parity(dims) == 0, se, so).repeat(12)
        Dx = S.apply_dw(x) + (fac - 1.0) * clover_apply(cl, x)       # scaled operator applied with the unscaled library state
        assert pc.rel(b, Dx) < 1.5e-10
        assert pc.rel(before, S.apply_dw(v)) < 1e-15                 # restored
        x2, res2, st2 = S.solve(b)
        assert st2[0] > 0 and res2 < 1e-10 and pc.rel(b, S.apply_dw(x2)) < 1.5e-10
    finally:
        S.free()


def test_setup_update_and_raw_gauge_access(emu_lib):
    dims, U, S = make(emu_lib)
    try:
        st = S.setup(1)
        assert st[0] == 1
        st = S.setup_update(1)
        assert st[0] == 1
        b = np.ones(S.V * 12, dtype=np.complex128)
        x, res, sts = S.solve(b)
        assert sts[0] > 0 and res < 1e-10
        # raw pointer: D = U/2 in the reference's array format; writing it back unchanged leaves the operator unchanged
        g = S.gauge_pointer()
        Uc = (U[..., 0] + 1j * U[..., 1]).reshape(-1)
        assert np.abs((g[0::2] + 1j * g[1::2]) - 0.5 * Uc).max() == 0.0
        rng = np.random.default_rng(4)
        v = pc.crandom(rng, S.V * 12)
        before = S.apply_dw(v)
        S.fields_updated()
        assert pc.rel(before, S.apply_dw(v)) < 1e-15
        # flipping the sign of every link flips the hopping term: D' v = 2 C v - D v
        g *= -1.0
        S.fields_updated()
        D, cl = S.operator_arrays()
        assert pc.rel(2.0 * clover_apply(cl, v) - before, S.apply_dw(v)) < 1e-13
    finally:
        S.free()


def test_struct_route_initialisation(emu_lib):
    """dd_alpha_amg_init_external_threading: geometry from dd_alpha_amg_parameters (X,Y,Z,T order)."""
    dims, plaq, U = read_conf(CONF4)
    S = DDalphaAMG.from_struct(dims, [2, 2, 2, 2], levels=2, test_vectors=(12,), setup_iter=(2,), m0=-0.2, lib=emu_lib)
    try:
        assert abs(S.set_conf(U) - plaq) < 1e-12
        S.setup(2)
        assert S.level_shape(0) == (256, 12) and S.level_shape(1) == (16, 24)
        b = np.ones(S.V * 12, dtype=np.complex128)
        x, res, st = S.solve(b)
        assert st[0] > 0 and res < 1e-10 and pc.rel(b, S.apply_dw(x)) < 1.5e-10
    finally:
        S.free()

"""Dirichlet boundary conditions in time (bc = 0 of dd_alpha_amg_par, dd_alpha_amg.c:206-235): the time links on the
slices 0, T-2, T-1 enter only the clover term, the links of the last slice must be zero.  Own module: one live
reference instance per process."""
import numpy as np

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


def test_dirichlet_boundary_conditions_vs_reference(oracle_ref, emu_lib):
    dims, plaq, U = read_conf(CONF8, anti_pbc=False)
    U[-1, :, :, :, 0] = 0.0                       # expected by the interface for bc = 0
    kw = dict(levels=2, test_vectors=(12,), setup_iter=(1,), restart=20)
    R = oracle_ref.Reference(dims, [4, 4, 4, 4], bc=0, anti_pbc=0, **kw)
    S = DDalphaAMG(dims, [4, 4, 4, 4], lib=emu_lib, bc=0, **kw)
    try:
        pr, ps = R.set_conf(U), S.set_conf(U)
        assert abs(pr - ps) < 1e-12
        D, cl = S.operator_arrays()
        assert np.abs(D - R.D()).max() == 0.0
        assert np.abs(cl - R.clover()).max() <= 1e-13
        rng = np.random.default_rng(8)
        v = pc.crandom(rng, S.V * 12)
        assert pc.rel(R.dw_double(v), S.apply_dw(v)) <= pc.TOL_DOUBLE
        S.setup(1)
        b = np.ones(S.V * 12, dtype=np.complex128)
        x, res, st = S.solve(b)
        assert st[0] > 0 and res < 1e-10 and pc.rel(b, R.dw_double(x)) < 1.5e-10
    finally:
        S.free()
        R.free()

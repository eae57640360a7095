"""`mixed precision: 2` outer solver (own module: the reference keeps process-global state, one live instance at a time)."""
import numpy as np
import pytest

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


def test_mixed_precision_2_outer_solver(oracle_ref, emu_lib):
    """`mixed precision: 2` (fgmres_MP, linsolve.c:153-424): float Arnoldi inside double restarts.  Same iteration
    count as the reference run with the same parameter and the reference's own interpolation imported."""
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=2, test_vectors=(20,), setup_iter=(2,), restart=10, mixed_precision=2)
    R = oracle_ref.Reference(dims, [4, 4, 4, 4], **kw)
    S = DDalphaAMG(dims, [4, 4, 4, 4], lib=emu_lib, **kw)
    try:
        R.set_conf(U)
        R.setup(2)
        S.set_conf(U)
        S.setup(0)
        pc.import_interpolation(R, S, 0)
        b = np.ones(S.V * 12, dtype=np.complex128)
        xr, resr, str_ = R.solve(b)
        xs, ress, sts = S.solve(b)
        assert sts[0] > 0 and ress < 1e-10
        assert abs(int(sts[0]) - int(str_[0])) <= 1
        assert pc.rel(b, R.dw_double(xs)) < 1.5e-10
        assert pc.rel(xr, xs) < 1e-8
    finally:
        S.free()
        R.free()

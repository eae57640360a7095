"""Three-level K-cycle host logic against the reference (see test_host_logic.py for what the emulation build is).
Kept in its own module: the reference holds process-global state, one live instance at a time."""
import numpy as np
import pytest

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


@pytest.mark.parametrize("single_reduction", ["0", "1"])
def test_three_level_kcycle_vs_reference(oracle_ref, emu_lib, single_reduction, monkeypatch):
    # "1": opt-in single-reduction Arnoldi of the coarse-level solvers (reference: SINGLE_ALLREDUCE_ARNOLDI)
    monkeypatch.setenv("DDA_SINGLE_REDUCTION", single_reduction)
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=50, coarse_block=[2, 2, 2, 2])
    R = oracle_ref.Reference(dims, [2, 2, 2, 2], **kw)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, **kw)
    try:
        R.set_conf(U)
        R.setup(1)
        S.set_conf(U)
        S.setup(0)
        out = pc.check_hierarchy(R, S, 3)
        pc.assert_hierarchy(out)
        b = np.ones(S.V * 12, dtype=np.complex128)
        xr, resr, str_ = R.solve(b)
        xs, ress, sts = S.solve(b)
        assert abs(int(sts[0]) - int(str_[0])) <= (0 if single_reduction == "0" else 1) and ress < 1e-10
    finally:
        S.free()
        R.free()



"""Three-level K-cycle on the GPU against the reference (own module: one live reference instance per process state)."""
import numpy as np
import pytest

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("teams", ["1", "4"])
def test_three_level_kcycle_vs_reference(oracle_ref, cuda_lib, teams, monkeypatch):
    # both thread-team configurations of the fused coarse-level SAP kernel (chosen by block count in production)
    monkeypatch.setenv("DDA_SAP_TEAMS", teams)
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=50, coarse_block=[2, 2, 2, 2])
    R = oracle_ref.Reference(dims, [2, 2, 2, 2], **kw)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=cuda_lib, **kw)
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
        assert abs(int(sts[0]) - int(str_[0])) <= 1 and ress < 1e-10
        # own setup, 3 levels: converges in a comparable number of iterations
        S.setup(2)
        xs, ress, sts = S.solve(b)
        assert sts[0] > 0 and ress < 1e-10 and abs(int(sts[0]) - int(str_[0])) <= 5
    finally:
        S.free()
        R.free()

"""Three-level K-cycle host logic against the reference (see test_host_logic.py for what the emulation build is).
Kept in its own module: the reference holds process-global state, one live instance at a time."""
import numpy as np

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


def test_three_level_kcycle_vs_reference(oracle_ref, emu_lib):
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
        assert int(sts[0]) == int(str_[0])
    finally:
        S.free()
        R.free()



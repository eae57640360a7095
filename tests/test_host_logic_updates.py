"""a21 (mass shift, clover scaling by site parity) and a near-critical solve against the reference -- host logic on the
emulation build; the same checks run on the CUDA library in test_gpu_api.py."""
import pytest

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


def pair(oracle_ref, lib, block, nv, m0):
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=2, test_vectors=(nv,), setup_iter=(2,), restart=20, m0=m0)
    R = oracle_ref.Reference(dims, block, **kw)
    R.set_conf(U)
    R.setup(2)
    S = DDalphaAMG(dims, block, lib=lib, **kw)
    S.set_conf(U)
    S.setup(0)
    return R, S


def test_mass_shift_and_clover_scaling_vs_reference(emu_lib, oracle_ref):
    R, S = pair(oracle_ref, emu_lib, [4, 4, 4, 4], 20, -0.5)
    try:
        pc.check_mass_shift_and_clover_scaling(R, S, -0.62)
    finally:
        S.free()
        R.free()


def test_near_critical_solve_vs_reference(emu_lib, oracle_ref):
    R, S = pair(oracle_ref, emu_lib, [2, 2, 2, 2], 12, -0.85)
    try:
        pc.check_near_critical_solve(R, S)
    finally:
        S.free()
        R.free()

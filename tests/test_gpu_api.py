"""GPU parity of the update paths: mass shift on every level, clover scaling by site parity inside
dd_alpha_amg_wilson_solve, and a near-critical solve whose coarsest-level GMRES runs > 64 steps per restart -- all
through the C ABI against the unmodified reference."""
import pytest

import parity_common as pc
from test_host_logic_updates import pair

pytestmark = pytest.mark.gpu


def test_mass_shift_and_clover_scaling_vs_reference_gpu(cuda_lib, oracle_ref):
    R, S = pair(oracle_ref, cuda_lib, [4, 4, 4, 4], 20, -0.5)
    try:
        assert not S.emulated
        pc.check_mass_shift_and_clover_scaling(R, S, -0.62)
    finally:
        S.free()
        R.free()


def test_near_critical_solve_vs_reference_gpu(cuda_lib, oracle_ref):
    R, S = pair(oracle_ref, cuda_lib, [2, 2, 2, 2], 12, -0.85)
    try:
        assert not S.emulated
        pc.check_near_critical_solve(R, S)
    finally:
        S.free()
        R.free()

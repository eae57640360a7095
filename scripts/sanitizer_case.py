"""Small end-to-end case sized for compute-sanitizer (racecheck / memcheck; the tool is closed on this pool in round 1): 8^4 fixture, 3 levels, initial setup + one solve.
Exercises k_dw_full, k_sap_fine, k_coarse_full, k_coarse_sap_mr (both team configurations via DDA_SAP_TEAMS),
k_restrict_fine and the generic kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ddalphaamg_b200 import DDalphaAMG, read_conf  # noqa: E402

dims, plaq, U = read_conf(os.path.join(ROOT, "tests", "golden", "conf_8x8x8x8b6.0000id3n1"))
S = DDalphaAMG(dims, [2, 2, 2, 2], levels=3, test_vectors=(8, 12), setup_iter=(1, 1), restart=10, coarse_block=[2, 2, 2, 2])
S.set_conf(U)
S.setup(0)
x, res, st = S.solve(np.ones(S.V * 12, dtype=np.complex128), tol=1e-6)
print("sanitizer case: iterations", st, "residual", res)
S.free()
# 4^4 blocks on the fine level -> the fused fine SAP kernel
S = DDalphaAMG(dims, [4, 4, 4, 4], levels=2, test_vectors=(8,), setup_iter=(1,), restart=10)
S.set_conf(U)
S.setup(0)
x, res, st = S.solve(np.ones(S.V * 12, dtype=np.complex128), tol=1e-6)
print("sanitizer case (fused fine SAP): iterations", st, "residual", res)
S.free()

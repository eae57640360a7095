"""12 right-hand sides at once on the tensor cores (tcgen05, TF32 x 3): every column must equal the single-RHS coarse
operator -- the reference's apply_coarse_operator_float with the reference's prolongator imported, and this library's own
single-RHS kernel -- within the float tolerance of BASELINE.json (1e-5 relative L2 per operator apply)."""
import numpy as np
import pytest

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc

pytestmark = pytest.mark.gpu


def test_multi_rhs_coarse_apply_vs_reference(oracle_ref, cuda_lib):
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=50, coarse_block=[2, 2, 2, 2])
    R = oracle_ref.Reference(dims, [2, 2, 2, 2], **kw)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=cuda_lib, **kw)
    try:
        R.set_conf(U)
        R.setup(1)
        S.set_conf(U)
        S.setup(0)
        for d in range(2):
            pc.import_interpolation(R, S, d)
        rng = np.random.default_rng(31)
        for d in (1, 2):                                   # n = 40 (256 sites) and n = 56 (16 sites)
            V, nc = S.level_shape(d)
            vs = np.stack([pc.crandom(rng, V * nc, np.complex64) for _ in range(12)])
            out, _ = S.level_apply_mrhs(d, vs)
            for j in range(12):
                assert pc.rel(R.coarse_apply(d, vs[j]), out[j]) <= pc.TOL_FLOAT, (d, j)
                assert pc.rel(S.level_apply(d, vs[j]), out[j]) <= pc.TOL_FLOAT, (d, j)
    finally:
        S.free()
        R.free()

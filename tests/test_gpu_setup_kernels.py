"""Setup kernels on the GPU: the fused aggregate orthonormalisation (CholeskyQR2 in double, one CTA per aggregate and
chirality) against the reference's gram_schmidt_on_aggregates (linalg_generic.c:400-454, reached through the reference's own
"interpolation: 4" test-vector reader) and against this library's generic Gram-Schmidt kernels; orthonormality as a property
(P^H P = 1 on every level)."""
import os

import numpy as np
import pytest

from conftest import CONF8
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc

pytestmark = pytest.mark.gpu


def _orthonormality(S, depth, rng):
    vc = pc.crandom(rng, int(np.prod(S.level_shape(depth + 1))), np.complex64)
    return pc.rel(S.restrict(depth, S.interpolate(depth, vc)), vc)


def test_aggregate_orthonormalisation_vs_reference(cuda_lib, oracle_ref, tmp_path, monkeypatch):
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=2, test_vectors=(20,), setup_iter=(1,), restart=20, m0=-0.3)
    S = DDalphaAMG(dims, [4, 4, 4, 4], lib=cuda_lib, **kw)
    base = os.path.join(str(tmp_path), "tv")
    rng = np.random.default_rng(11)
    try:
        S.set_conf(U)
        S.setup(1)
        P = S.get_interpolation(0)
        assert _orthonormality(S, 0, rng) < 2e-6
        S.write_test_vectors(base)
        R = oracle_ref.Reference(dims, [4, 4, 4, 4], interpolation=4, tv_file=base, **kw)
        try:
            R.set_conf(U)
            R.setup(1)
            tt = R.translation(0)
            V, nc = S.level_shape(0)
            Pr = R.interpolation(0).reshape(V, nc, -1)[tt].reshape(V * nc, -1)
            assert pc.rel(Pr, P) < 1e-5
        finally:
            R.free()
        # the generic Gram-Schmidt kernels on the same test vectors
        monkeypatch.setenv("DDA_GS_FAST", "0")
        S.read_test_vectors(base)
        Pg = S.get_interpolation(0)
        monkeypatch.delenv("DDA_GS_FAST")
        assert pc.rel(Pg, P) < 1e-5
        S.read_test_vectors(base)
        assert pc.rel(S.get_interpolation(0), P) < 1e-6
    finally:
        S.free()


def test_coarse_level_orthonormalisation(cuda_lib, monkeypatch):
    dims, plaq, U = read_conf(CONF8)
    kw = dict(levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=20, m0=-0.3, coarse_block=[2, 2, 2, 2])
    rng = np.random.default_rng(12)
    Ps = []
    for fast in ("1", "0"):
        monkeypatch.setenv("DDA_GS_FAST", fast)
        S = DDalphaAMG(dims, [2, 2, 2, 2], lib=cuda_lib, **kw)
        try:
            S.set_conf(U)
            S.setup(0)                                  # no bootstrap iteration: P = orthonormalised smoothed random vectors
            for d in (0, 1):
                assert _orthonormality(S, d, rng) < 3e-6, (fast, d)
            Ps.append(S.get_interpolation(0))
        finally:
            S.free()
    assert pc.rel(Ps[0], Ps[1]) < 1e-5

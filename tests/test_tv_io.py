"""Test-vector files in the reference's vector format: vectors written by this library are read by the UNMODIFIED
reference through its own "interpolation: 4" path (vector_io, io.c:704; read_tv_from_file, setup_generic.c:131), and the
interpolation operator the reference builds from them must equal this library's; the library's own reader (parameter-file
route and dda_read_test_vectors) must reproduce the hierarchy it was written from.  Plus the struct-route setup policy
(dd_alpha_amg.c:85-93)."""
import os

import numpy as np

from conftest import CONF4
from ddalphaamg_b200 import DDalphaAMG, read_conf
import parity_common as pc


def test_test_vector_files_reference_format(emu_lib, oracle_ref, tmp_path):
    dims, plaq, U = read_conf(CONF4)
    kw = dict(levels=2, test_vectors=(12,), setup_iter=(2,), restart=20, m0=-0.3)
    S = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, **kw)
    base = os.path.join(str(tmp_path), "tv")
    try:
        S.set_conf(U)
        S.setup(2)
        P = S.get_interpolation(0)
        rng = np.random.default_rng(8)
        vc = pc.crandom(rng, np.prod(S.level_shape(1)), np.complex64)
        Dc = S.level_apply(1, vc)
        S.write_test_vectors(base)
        assert os.path.getsize(base + ".00") == S.V * 24 * 8 and os.path.exists(base + ".11")
        # the reference reads the files (its own reader) and rebuilds its interpolation from them
        R = oracle_ref.Reference(dims, [2, 2, 2, 2], interpolation=4, tv_file=base, **kw)
        try:
            R.set_conf(U)
            R.setup(1)
            tt = R.translation(0)
            Pr = R.interpolation(0)
            V, nc = S.level_shape(0)
            Pr_lex = Pr.reshape(V, nc, -1)[tt].reshape(V * nc, -1)
            assert pc.rel(Pr_lex, P) < 2e-6
            assert pc.rel(R.coarse_apply(1, vc), Dc) <= pc.TOL_FLOAT
        finally:
            R.free()
        # own reader, explicit call: perturb the hierarchy, read back, same operators again
        S.setup(1)
        assert pc.rel(P, S.get_interpolation(0)) > 1e-3
        S.read_test_vectors(base)
        assert pc.rel(P, S.get_interpolation(0)) < 1e-6 and pc.rel(Dc, S.level_apply(1, vc)) < 1e-5
    finally:
        S.free()
    # own reader, parameter-file route ("interpolation: 4")
    S2 = DDalphaAMG(dims, [2, 2, 2, 2], lib=emu_lib, interpolation=4, tv_file=base, **kw)
    try:
        S2.set_conf(U)
        S2.setup(1)
        assert pc.rel(P, S2.get_interpolation(0)) < 1e-6
        b = np.ones(S2.V * 12, dtype=np.complex128)
        x, res, st = S2.solve(b)
        assert st[0] > 0 and res < 1e-10
    finally:
        S2.free()


def test_setup_policy_struct_route(emu_lib):
    dims, plaq, U = read_conf(CONF4)
    S = DDalphaAMG.from_struct(dims, [2, 2, 2, 2], levels=2, test_vectors=(12,), setup_iter=(1,), m0=-0.2, lib=emu_lib)
    try:
        S.set_conf(U)
        assert S.setup_if_necessary() == 2          # counters start at their thresholds (init.c:899-900)
        assert S.setup_if_necessary() == 2          # discard_setup_after = 0 in a zeroed struct: every check is a full setup
    finally:
        S.free()

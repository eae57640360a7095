"""Generates tests/golden/golden_4x4x4x4.npz from the unmodified reference (oracle/_ref/libddref.so built by
oracle/build_ref.sh from /root/reference).  Run in the build container: python tests/golden/make_golden.py

Contents (4^4 fixture conf/4x4x4x4b6.0000id3n1, anti-periodic, m0 = -0.5, csw = 1, 2 levels, blocks 2^4, 20 test
vectors, 2 setup iterations; all vectors in lexicographic site order):
  dw_in, dw_out             d_plus_clover_double                        (dirac_generic.c:159)
  P                         is_float.operator, [site][12][Nv]           (interpolation_generic.c:74-90)
  coarse_in, coarse_out     coarse operator D_c                         (coarse_operator_generic.c:383)
  restrict_in/out, interpolate_in/out                                   (interpolation_generic.c:130-207)
  smoother_eta/phi0/out     red_black_schwarz_float, 2 iterations, _RES (schwarz_generic.c:1260)
  prec_in, prec_out         preconditioner = one float V-cycle          (preconditioner.c:25)
  coarsest_in/out           coarse_solve_odd_even_float                 (coarse_oddeven_generic.c:1139)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from ddalphaamg_b200 import read_conf  # noqa: E402


def crandom(rng, n, dtype):
    return (rng.uniform(-0.5, 0.5, n) + 1j * rng.uniform(-0.5, 0.5, n)).astype(dtype)


def main():
    dims, plaq, U = read_conf(os.path.join(HERE, "conf_4x4x4x4b6.0000id3n1"))
    R = ref.Reference(dims, [2, 2, 2, 2], levels=2, test_vectors=(20,), setup_iter=(2,), restart=10)
    R.set_conf(U)
    R.setup(2)
    rng = np.random.default_rng(4444)
    V = R.V
    Vc, nc = R.info(1, 1), R.info(2, 1)
    g = {}
    g["dw_in"] = crandom(rng, V * 12, np.complex128)
    g["dw_out"] = R.dw_double(g["dw_in"])
    tt = R.translation(0)
    P = R.interpolation(0)
    g["P"] = P.reshape(V, 12, P.shape[1])[tt].reshape(V * 12, P.shape[1])
    g["coarse_in"] = crandom(rng, Vc * nc, np.complex64)
    g["coarse_out"] = R.coarse_apply(1, g["coarse_in"])
    g["restrict_in"] = crandom(rng, V * 12, np.complex64)
    g["restrict_out"] = R.restrict(0, g["restrict_in"])
    g["interpolate_in"] = crandom(rng, Vc * nc, np.complex64)
    g["interpolate_out"] = R.interpolate(0, g["interpolate_in"])
    g["smoother_eta"] = crandom(rng, V * 12, np.complex64)
    g["smoother_phi0"] = crandom(rng, V * 12, np.complex64)
    g["smoother_out"] = R.smoother(0, g["smoother_eta"], 2, g["smoother_phi0"])
    g["prec_in"] = crandom(rng, V * 12, np.complex128)
    g["prec_out"] = R.preconditioner(g["prec_in"])
    g["coarsest_in"] = crandom(rng, Vc * nc, np.complex64)
    g["coarsest_out"], _ = R.coarsest_solve(g["coarsest_in"])
    x, res, st = R.solve(np.ones(V * 12, dtype=np.complex128))
    g["solve_iters"] = np.array([st[0], st[1]])
    g["solve_res"] = np.array([res])
    np.savez_compressed(os.path.join(HERE, "golden_4x4x4x4.npz"), **g)
    R.free()
    print("written", {k: v.shape for k, v in g.items()}, "solve", st, res)


if __name__ == "__main__":
    main()

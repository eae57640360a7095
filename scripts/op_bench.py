#!/usr/bin/env python
"""Operator timings of one workload with the library named by DDA_LIBRARY (default: the product build): A/B measurements
of kernel variants on the SAME box.  Prints one JSON line: milliseconds per application (CUDA events inside the library)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field, BENCH, INFO, library_path  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="32^3x64-L3")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    w = dict(bench.WORKLOADS[args.workload])
    lat = w["lattice"]
    kw = bench.solver_kwargs(w)
    S = DDalphaAMG(lat, [4, 4, 4, 4], **kw)
    S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
    S.setup(1)
    nlev = S.info(INFO.NUM_LEVELS)
    out = {"tag": args.tag, "library": os.path.basename(library_path()), "workload": args.workload, "env": {k: v for k, v in os.environ.items() if k.startswith("DDA_")}}
    for rnd in range(2):       # second round = the reported one (clocks settled)
        out["dw_double"] = S.bench_op(BENCH.DW_DOUBLE, 0, args.reps)
        out["dw_float"] = S.bench_op(BENCH.DW_FLOAT, 0, args.reps)
        for d in range(1, nlev):
            out["coarse_apply_d%d" % d] = S.bench_op(BENCH.LEVEL_APPLY, d, args.reps)
        for d in range(nlev - 1):
            out["restrict_d%d" % d] = S.bench_op(BENCH.RESTRICT, d, args.reps)
            out["interpolate_d%d" % d] = S.bench_op(BENCH.INTERPOLATE, d, args.reps)
            out["smoother_d%d" % d] = S.bench_op(BENCH.SMOOTHER, d, 5)
        out["coarsest_schur"] = S.bench_op(BENCH.COARSEST_SCHUR, nlev - 1, 50)
    if os.environ.get("DDA_BENCH_MRHS", "0") != "0":
        # 12 right-hand sides at once on the tensor cores: ms per 12-RHS application and the single-RHS time next to it
        rng = np.random.default_rng(3)
        for d in range(1, nlev):
            V, nc = S.level_shape(d)
            vs = (rng.standard_normal((12, V * nc)) + 1j * rng.standard_normal((12, V * nc))).astype(np.complex64)
            o, ms = S.level_apply_mrhs(d, vs, reps=10)
            one = S.level_apply(d, vs[5])
            out["mrhs12_apply_d%d_ms" % d] = ms
            out["mrhs12_apply_d%d_relerr_col5" % d] = float(np.linalg.norm(o[5] - one) / np.linalg.norm(one))
    b = np.ones(S.V * 12, dtype=np.complex128)
    S.solve_device(b)
    res, st, ms = S.solve_device(b)
    out["solve_ms"] = ms
    out["iterations"] = [int(st[0]), int(st[1])]
    S.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

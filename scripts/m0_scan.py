#!/usr/bin/env python
"""Mass scan on one GPU: ONE setup, then device-resident solves at several m0 (mass shift on every level through
dd_alpha_amg_update_parameters, reference shift_update dirac.c:669-691).  Prints outer / coarsest iterations, seconds per
solve and the time share + milliseconds per iteration of the coarsest-level solve.  Used to place bench.py's m0 at
20-50 coarsest iterations per cycle."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field, STAT, OPT, INFO  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--setup-m0", type=float, default=-0.3)
    ap.add_argument("--m0", type=float, nargs="+", default=[-0.1, -0.2, -0.3, -0.35, -0.4, -0.45])
    args = ap.parse_args()
    w = dict(bench.WORKLOADS[args.workload])
    w["m0"] = args.setup_m0
    lat = w["lattice"]
    kw = bench.solver_kwargs(w)
    U = random_gauge_field(lat, seed=20261018, eps=0.3)
    S = DDalphaAMG(lat, [4, 4, 4, 4], **kw)
    S.set_conf(U)
    del U
    t0 = time.time()
    S.setup(w["setup_iter"][0])
    print(json.dumps({"setup_seconds": time.time() - t0, "setup_m0": args.setup_m0}), flush=True)
    b = np.ones(S.V * 12, dtype=np.complex128)
    nlev = S.info(INFO.NUM_LEVELS)
    for m0 in args.m0:
        S.update_parameters(m0, setup_iter=tuple(list(w["setup_iter"]) + [2, 2])[:4])
        S.solve_device(b)
        res, st, ms = S.solve_device(b)
        S.set_option(OPT.PROFILE, 1)
        S.reset_stats()
        S.solve_device(b)
        tc = S.stat(STAT.T_COARSEST)
        ts = [S.stat(STAT.T_SMOOTH0 + d) for d in range(nlev - 1)]
        S.set_option(OPT.PROFILE, 0)
        cycles = max(1, int(st[0]))
        print(json.dumps({"m0": m0, "outer_iterations": int(st[0]), "coarsest_iterations": int(st[1]),
                          "coarsest_per_cycle": int(st[1]) / cycles, "solve_seconds": ms / 1e3, "residual": res,
                          "coarsest_seconds_profiled": tc, "coarsest_ms_per_iteration": 1e3 * tc / max(1, int(st[1])),
                          "smoother_seconds_profiled": ts}), flush=True)
    S.free()


if __name__ == "__main__":
    main()

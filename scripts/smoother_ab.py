#!/usr/bin/env python
"""A/B timing of the level-1 SAP smoother (k_coarse_sap_mr variants chosen by DDA_SAPMR_* / DDA_SAP_TEAMS) on one workload."""
import os
import sys

import torch  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field, BENCH  # noqa: E402

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "48^3x96-L3"])
lat = w["lattice"]
kw = bench.solver_kwargs(w)
kw["setup_iter"] = (0, 0)
S = DDalphaAMG(lat, [4, 4, 4, 4], **kw)
S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
S.setup(0)
S.bench_op(BENCH.SMOOTHER, 1, 3)
print({k: v for k, v in os.environ.items() if k.startswith("DDA_")}, "smoother_d1 ms", S.bench_op(BENCH.SMOOTHER, 1, 10),
      "apply_d1 ms", S.bench_op(BENCH.LEVEL_APPLY, 1, 10))
S.free()

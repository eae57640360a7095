#!/usr/bin/env python
"""Setup phase timers (DDA_SETUP_PROFILE=1) of one benchmark workload, full setup as bench.py runs it."""
import os
import sys
import time

import torch  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DDA_SETUP_PROFILE", "1")
import bench  # noqa: E402
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field  # noqa: E402

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "48^3x96-L3"])
lat = w["lattice"]
S = DDalphaAMG(lat, [4, 4, 4, 4], **bench.solver_kwargs(w))
S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
t0 = time.time()
S.setup(w["setup_iter"][0])
print({k: v for k, v in os.environ.items() if k.startswith("DDA_")}, "setup seconds", time.time() - t0)
S.free()

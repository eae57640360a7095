#!/usr/bin/env python
"""Small, fixed launch sequences for ncu captures (run once plain, then under ncu with -k regex:<kernel>).
    python scripts/ncu_case.py sap      fine-level SAP smoother calls on 32^3x64 (k_sap_fine2: 4096 block visits per launch)
    python scripts/ncu_case.py hier     3-level 32^3x64: D_W, coarse applies, level-1 SAP, restrict / interpolate, Schur complement
"""
import os
import sys

import numpy as np
import torch  # noqa: F401  (before the library: torch bundles its own NCCL, which must be the first one the process loads)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ddalphaamg_b200 import DDalphaAMG, random_gauge_field, BENCH, INFO  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "sap"
    lat = [64, 32, 32, 32]
    if case == "sap48":
        lat = [96, 48, 48, 48]          # the benchmark's lattice: 20736 block visits per launch
        case = "sap"
    if case == "mrhs":
        S = DDalphaAMG(lat, [4, 4, 4, 4], levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=10, m0=-0.35, csw=1.0,
                       mixed_precision=2, coarse_block=[2, 2, 2, 2])
        S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
        S.setup(0)
        rng = np.random.default_rng(3)
        for d in (1, 2):
            V, nc = S.level_shape(d)
            vs = (rng.standard_normal((12, V * nc)) + 1j * rng.standard_normal((12, V * nc))).astype(np.complex64)
            o, ms = S.level_apply_mrhs(d, vs, reps=10)
            print("mrhs depth", d, "ms per 12-RHS apply", ms, "single", S.bench_op(BENCH.LEVEL_APPLY, d, 10))
        S.free()
        return
    if case == "galerkin":
        # setup kernels (fused Galerkin construction, aggregate orthonormalisation) on 32^3x64: capture with -k regex:<kernel>
        S = DDalphaAMG(lat, [4, 4, 4, 4], levels=3, test_vectors=(20, 28), setup_iter=(0, 0), restart=10, m0=-0.35, csw=1.0,
                       mixed_precision=2, coarse_block=[2, 2, 2, 2])
        S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
        S.setup(0)
        S.free()
        return
    if case == "sapmr48":
        # level-1 SAP block solves at the benchmark size (24 x 12^3 coarse lattice, 3x2x2x2 blocks: 864 blocks per launch)
        lat = [96, 48, 48, 48]
        S = DDalphaAMG(lat, [4, 4, 4, 4], levels=3, test_vectors=(20, 28), setup_iter=(0, 0), restart=10, m0=-0.35, csw=1.0,
                       mixed_precision=2, coarse_block=[3, 2, 2, 2])
        S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
        S.setup(0)
        torch.cuda.profiler.start()
        print("sap d1", S.bench_op(BENCH.SMOOTHER, 1, 2))
        torch.cuda.profiler.stop()
        S.free()
        return
    if case == "sap":
        S = DDalphaAMG(lat, [4, 4, 4, 4], levels=2, test_vectors=(4,), setup_iter=(0,), restart=10, m0=-0.35, csw=1.0, mixed_precision=2)
        S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
        S.setup(0)                     # 4 test vectors x (1+2+3) smoother iterations x 2 colours = 48 launches of the SAP kernel
        print("smoother ms", S.bench_op(BENCH.SMOOTHER, 0, 3))
    else:
        S = DDalphaAMG(lat, [4, 4, 4, 4], levels=3, test_vectors=(20, 28), setup_iter=(1, 1), restart=10, m0=-0.35, csw=1.0,
                       mixed_precision=2, coarse_block=[2, 2, 2, 2])
        S.set_conf(random_gauge_field(lat, seed=20261018, eps=0.3))
        S.setup(0)
        nlev = S.info(INFO.NUM_LEVELS)
        torch.cuda.profiler.start()      # ncu --profile-from-start off: only the operator applications below
        print("dw double", S.bench_op(BENCH.DW_DOUBLE, 0, 2), "float", S.bench_op(BENCH.DW_FLOAT, 0, 2))
        for d in range(1, nlev):
            print("apply", d, S.bench_op(BENCH.LEVEL_APPLY, d, 2))
        for d in range(nlev - 1):
            print("restrict", d, S.bench_op(BENCH.RESTRICT, d, 2), "interpolate", S.bench_op(BENCH.INTERPOLATE, d, 2))
        print("sap d1", S.bench_op(BENCH.SMOOTHER, 1, 2))
        print("schur", S.bench_op(BENCH.COARSEST_SCHUR, nlev - 1, 4))
        torch.cuda.profiler.stop()
    S.free()


if __name__ == "__main__":
    main()

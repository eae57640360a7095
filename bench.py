#!/usr/bin/env python
"""bench.py -- time-to-solution of the device-resident DDalphaAMG solve + operator HBM GB/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one complete solve (outer FGMRES preconditioned by the K-cycle, relative residual 1e-10) of
D_W x = b on a synthetic SU(3) gauge field of the workload's lattice; the multigrid hierarchy is set up once,
untimed, by the library's own device-side setup.  `value` = seconds per solve with the source resident in HBM
(CUDA events on the library's stream, max over ranks); `e2e` = seconds per solve through the reference-facing
C ABI call dd_alpha_amg_wilson_solve with pinned HOST source/solution buffers (host<->device copies inside the
timed region).  `roofline` describes the dominant kernel of the solve, `operators` lists algorithmic HBM GB/s of
the fine D_W, coarse-operator, restrict and interpolate kernels (SURVEY.md section 8d formulas).  `cpu_baseline`
(and the whole `--impl reference` arm) time the UNMODIFIED reference (oracle/_ref, SSE flavour = its default build)
on the box's host cores on a bounded sample lattice and scale seconds-per-site to the workload's volume.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "solve time @48^3x96 + D_W/coarse-apply HBM GB/s vs peak, at 1/2/4/8 B200"

# lattice order T Z Y X.  m0 chosen per synthetic field (warm SU(3), eps 0.3, plaquette ~1.83/3) so that the coarsest
# solver needs a few tens of iterations per cycle.
WORKLOADS = {
    "16^3x32-L2": dict(lattice=[32, 16, 16, 16], levels=2, test_vectors=(20,), setup_iter=(3,), m0=-0.1,
                       config="configs[1]: 16^3x32 synthetic random SU(3) gauge field, 2-level AMG"),
    "32^3x64-L3": dict(lattice=[64, 32, 32, 32], levels=3, test_vectors=(20, 28), setup_iter=(3, 2), m0=-0.1,
                       coarse_block=[2, 2, 2, 2],
                       config="configs[2]: 32^3x64 synthetic gauge field, 3-level AMG, mixed float/double"),
    # level-1 blocks 3x2x2x2: the T extents 96 -> 24 -> 8 stay divisible by 8 ranks on every level (T-partition)
    "48^3x96-L3": dict(lattice=[96, 48, 48, 48], levels=3, test_vectors=(20, 28), setup_iter=(3, 2), m0=-0.1,
                       coarse_block=[3, 2, 2, 2],
                       config="configs[3]: 48^3x96 synthetic gauge field, 3-level AMG"),
    "64^3x128-L3": dict(lattice=[128, 64, 64, 64], levels=3, test_vectors=(20, 28), setup_iter=(3, 2), m0=-0.1,
                        coarse_block=[2, 2, 2, 2],
                        config="configs[4]: 64^3x128 synthetic gauge field, 3-level AMG (needs >= 4 GPUs; single RHS)"),
}
DEFAULT_WORKLOAD = "48^3x96-L3"
CPU_SAMPLE = {2: [16, 8, 8, 8], 3: [16, 16, 16, 16]}


def solver_kwargs(w):
    kw = dict(levels=w["levels"], test_vectors=w["test_vectors"], setup_iter=w["setup_iter"], restart=10,
              max_restart=50, m0=w["m0"], csw=1.0, tol=1e-10, mixed_precision=w.get("mixed_precision", 2))
    if w["levels"] > 2:
        kw["coarse_block"] = w.get("coarse_block", [2, 2, 2, 2])
    return kw


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.index, self.proc = [], index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def reference_solve_time(levels, w, steps, warmup):
    """Times the unmodified reference (oracle/_ref) on a bounded sample lattice, all host cores.  Returns
    (seconds per solve on the sample, sample description, cores, flavour, iterations, D_W seconds per apply)."""
    from oracle import ref
    from ddalphaamg_b200 import random_gauge_field
    flavour = "_sse" if ref.available("_sse") else ""
    if not ref.available(flavour):
        raise RuntimeError("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    cores = os.cpu_count() or 1
    lat = CPU_SAMPLE[levels]
    kw = solver_kwargs(w)
    if levels > 2:
        kw["coarse_block"] = [2, 2, 2, 2]       # the sample lattice has its own (smaller) geometry
    U = random_gauge_field(lat, seed=20261018, eps=0.3)
    R = ref.Reference(lat, [4, 4, 4, 4], flavour=flavour, nthreads=cores, **{k: v for k, v in kw.items() if k not in ("csw", "m0")},
                      m0=w["m0"], csw=1.0)
    R.set_conf(U)
    R.setup(w["setup_iter"][0], nthreads=cores)
    b = np.ones(R.V * 12, dtype=np.complex128)
    secs, its = [], None
    for i in range(warmup + steps):
        x, res, st, sec = R.solve_mt(b)
        if res > 1e-10 or st[0] < 0:
            raise RuntimeError("reference solve did not converge: %g %s" % (res, st))
        its = [int(st[0]), int(st[1])]
        if i >= warmup:
            secs.append(sec)
    dw = R.dw_time(10, cores)
    V = R.V
    R.free()
    desc = "%dx%dx%dx%d (TxZxYxX) lattice, %d-level, same solver parameters and gauge recipe; %d solves, mean" % (
        lat[0], lat[1], lat[2], lat[3], levels, steps)
    return float(np.mean(secs)), desc, cores, flavour, its, dw, V


def run_reference(args, w, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Vw = int(np.prod(w["lattice"]))
    sec, desc, cores, flavour, its, dw, Vs = reference_solve_time(w["levels"], w, args.steps, args.warmup)
    val = sec * Vw / Vs
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * val, "higher_is_better": False, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64 outer / f32 cycle", "data": "synthetic",
           "config": {"workload": name, "detail": w["config"],
                      "note": "CPU reference (unmodified DDalphaAMG%s, 1 rank x %d OpenMP threads) timed on the sample and "
                              "scaled by volume (%d/%d sites); no MPI on this box, so no multi-rank CPU run" % (
                                  " SSE build" if flavour else "", cores, Vw, Vs)},
           "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "reference", "sample": desc,
                            "sample_seconds": sec, "sample_iterations": its,
                            "dw_double_gbs": 1632.0 * Vs / dw / 1e9},
           "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------
def run_native(args, w, name):
    import torch
    from ddalphaamg_b200 import DDalphaAMG, random_gauge_field, STAT, OPT, BENCH, INFO
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    os.environ["DDA_DEVICE"] = str(local_rank)
    lat = w["lattice"]
    kw = solver_kwargs(w)
    peak, peak_src = peaks()
    if world > 1:
        # one process per GPU, lattice partitioned along T; torch.distributed only carries the NCCL id and the timing max
        import torch.distributed as dist
        from ddalphaamg_b200.interface import comm_init
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        comm_init()
        if lat[0] % world:
            raise RuntimeError("T extent %d is not divisible by %d ranks" % (lat[0], world))
        kw["local_lattice"] = [lat[0] // world] + lat[1:]
    lt = lat[0] // world
    V = int(np.prod(lat)) // world           # local sites
    Vglob = int(np.prod(lat))

    def rank_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    t0 = time.time()
    U = random_gauge_field(lat, seed=20261018, eps=0.3, t_range=(rank * lt, (rank + 1) * lt))
    t_gauge = time.time() - t0
    S = DDalphaAMG(lat, [4, 4, 4, 4], **kw)
    plaq = S.set_conf(U)
    del U
    t0 = time.time()
    S.setup(w["setup_iter"][0])
    t_setup = time.time() - t0

    n = V * 12
    hb = torch.ones(n, dtype=torch.complex128).pin_memory()
    hx = torch.zeros(n, dtype=torch.complex128).pin_memory()
    b, x = hb.numpy(), hx.numpy()

    # ---- device-resident solves
    for _ in range(args.warmup):
        res, st, ms = S.solve_device(b)
    clk = ClockSampler(local_rank)
    clk.start()
    S.reset_stats()
    barrier()
    torch.cuda.profiler.start()      # ncu --profile-from-start off: timed solves + operator benchmarks only
    tw0 = time.time()
    tot_ms, its = 0.0, None
    for _ in range(args.steps):
        res, st, ms = S.solve_device(b)
        tot_ms += ms
        its = [int(st[0]), int(st[1])]
    barrier()
    wall = rank_max(time.time() - tw0)
    tot_ms = rank_max(tot_ms)        # device time (CUDA events on the library's stream), max over ranks
    launches = int(S.stat(STAT.LAUNCHES))
    clocks = clk.stop()
    if res > 1e-10 or its[0] < 0:
        raise RuntimeError("solve did not converge: %g %s" % (res, its))
    sec = tot_ms / 1e3 / args.steps

    # ---- end to end through the reference-facing C ABI, host buffers
    for _ in range(min(args.warmup, 2)):
        S.solve(b, out=x)
    barrier()
    te = time.time()
    for _ in range(args.steps):
        _, res_e, st_e = S.solve(b, out=x)
    barrier()
    e2e = rank_max(time.time() - te) / args.steps

    # ---- operator throughput (CUDA events inside the library, device-resident)
    nlev = S.info(INFO.NUM_LEVELS)
    ops = {}
    reps = 20

    def add(key, ms_, bytes_):
        # per-rank algorithmic bytes of the local volume, slowest rank's time: GB/s and roofline fraction PER GPU
        ms_ = rank_max(ms_)
        g = bytes_ / (ms_ * 1e-3) / 1e9
        ops[key] = {"ms": ms_, "algorithmic_bytes": bytes_, "gbs": g, "frac_of_peak": g / peak}

    add("dw_double", S.bench_op(BENCH.DW_DOUBLE, 0, reps), 1632.0 * V)
    add("dw_float", S.bench_op(BENCH.DW_FLOAT, 0, reps), 816.0 * V)
    for d in range(1, nlev):
        Vc, nc = S.level_shape(d)
        add("coarse_apply_d%d_n%d" % (d, nc), S.bench_op(BENCH.LEVEL_APPLY, d, reps), (4 * nc * nc + nc * (nc + 1) // 2 + 2 * nc) * 8.0 * Vc)
    for d in range(nlev - 1):
        Vd, nc = S.level_shape(d)
        Vc, ncc = S.level_shape(d + 1)
        nv = ncc // 2
        tb = (nc * nv + nc) * 8.0 * Vd + ncc * 8.0 * Vc
        add("restrict_d%d" % d, S.bench_op(BENCH.RESTRICT, d, reps), tb)
        add("interpolate_d%d" % d, S.bench_op(BENCH.INTERPOLATE, d, reps), tb)
    # SAP smoother of the fine level: per block visit 256*(288+336) + 5*256*96 bytes, 2 colours x post_smooth_iter
    nblk = S.info(INFO.NUM_BLOCKS, 0)
    bs = S.info(INFO.BLOCK_SITES, 0)
    sap_ms = S.bench_op(BENCH.SMOOTHER, 0, 5)
    add("sap_smoother_d0", sap_ms, 2 * nblk * (bs * (288 + 336) + 5 * bs * 96.0))
    for d in range(1, nlev - 1):
        # coarse-level SAP (2 iterations x 2 colours): per colour the block residual (operator on half the sites) and
        # block_iter block-operator applications (self coupling + in-block hops ~ 3 blocks of n^2 per site)
        Vd, nc = S.level_shape(d)
        per_colour = (Vd / 2) * ((4 * nc * nc + nc * (nc + 1) // 2) + 4 * 3 * nc * nc) * 8.0
        add("sap_smoother_d%d" % d, S.bench_op(BENCH.SMOOTHER, d, 3), 4 * per_colour)

    torch.cuda.profiler.stop()

    # ---- time shares of one profiled solve (device-synchronising timers per operator class)
    S.set_option(OPT.PROFILE, 1)
    S.reset_stats()
    _, _, ms_prof = S.solve_device(b)
    share = {"smoother_d%d" % d: S.stat(STAT.T_SMOOTH0 + d) for d in range(nlev - 1)}
    share.update({"op_d%d" % d: S.stat(STAT.T_OP0 + d) for d in range(nlev)})
    share.update({"coarsest_solve": S.stat(STAT.T_COARSEST), "restrict": S.stat(STAT.T_RESTRICT), "interpolate": S.stat(STAT.T_INTERPOLATE)})
    S.set_option(OPT.PROFILE, 0)
    dev_bytes = S.stat(STAT.DEVICE_BYTES)
    S.free()

    # dominant kernel of the solve: k_sap_fine (one launch = the block visits of one colour; a smoother call = 2 iterations
    # x 2 colours = 4 launches).  Algorithmic bytes per block visit: 256*(288+336) + 5*256*96 = 282 624 B.
    # DRAM traffic per block visit measured by ncu --set full (profiles/r1_ncu_full_k_sap_fine_final.txt): 1.545 GB per
    # launch of 4096 visits = 377 KB.
    dom = ops["sap_smoother_d0"]
    visits_per_launch = nblk / 2.0
    launch_ms = dom["ms"] / 4.0
    roof = {"kernel": "k_sap_fine: fused fine-level SAP block solve, one CTA per 4^4 Schwarz block (%d block visits per launch, "
                      "avg launch %.3f ms from CUDA events over %d launches)" % (int(visits_per_launch), launch_ms, 4 * 5),
            "bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak,
            "traffic": 377.2e3 * visits_per_launch, "algorithmic_bytes_per_launch": 282624.0 * visits_per_launch,
            "peak_source": peak_src,
            "note": "arithmetic intensity 8.5 flop/B: the kernel is fp32-issue bound (ncu: 52 % issue slots, 18 % DRAM), "
                    "see DESIGN.md section 4; D_W and coarse-operator GB/s (the metric's named kernels) are in `operators`"}

    if world > 1:
        from ddalphaamg_b200.interface import comm_finalize
        comm_finalize()
        dist.barrier()
        dist.destroy_process_group()

    out = {"metric": METRIC, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * sec, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64 restarts + f32 Arnoldi/cycle (mixed precision %d)" % w.get("mixed_precision", 2), "data": "synthetic",
           "config": {"workload": name, "detail": w["config"], "lattice_TZYX": lat, "local_lattice_TZYX": [lt] + lat[1:],
                      "partition": "T split over %d GPU(s), NCCL send/recv halos + allreduce" % world, "levels": w["levels"],
                      "test_vectors": list(w["test_vectors"]), "mixed_precision": w.get("mixed_precision", 2), "m0": w["m0"], "csw": 1.0, "tol": 1e-10,
                      "gauge": "U=exp(i*0.3*H), H Gaussian traceless Hermitian, seed 20261018, plaquette %.6f" % plaq,
                      "l2": "working set %.1f GB per solve >> 126 MB L2 (inputs larger than L2, no flush)" % (dev_bytes / 1e9),
                      "iterations": its, "setup_seconds_untimed": t_setup},
           "e2e": {"value": e2e, "unit": "s", "h2d_bytes_per_step": n * 16 * world, "d2h_bytes_per_step": n * 16 * world},
           "gpu_launches": launches, "clocks": clocks, "roofline": roof, "operators": ops,
           "time_share_seconds_profiled_solve": share, "wall_seconds_timed_region": wall}

    if rank == 0 and world == 1 and not args.no_cpu:
        # the reference runs in its own process (it aborts the process on any error, main.h:424-439)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
                                "--workload", name, "--m0", str(w["m0"])], capture_output=True, text=True, timeout=900)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            out["cpu_baseline"] = json.loads(line)["cpu_baseline"]
        except Exception as e:  # the baseline is reported, never required for the product number
            out["cpu_baseline"] = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}
    if rank == 0:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--m0", type=float, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--mixed-precision", type=int, default=2, choices=[1, 2],
                    help="reference parameter `mixed precision` (both arms): 2 = fgmres_MP, the reference's default "
                         "(init.c:581-962); 1 = double outer FGMRES")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    w["mixed_precision"] = args.mixed_precision
    if args.m0 is not None:
        w["m0"] = args.m0
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_native(args, w, args.workload)


if __name__ == "__main__":
    main()

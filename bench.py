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

# lattice order T Z Y X.  The synthetic field (U = exp(i 0.3 H), plaquette ~1.83/3) has its critical mass near -0.5 at
# these volumes.  m0 = -0.35 is where the 48^3x96 solve needs 20-50 coarsest-level iterations per cycle (SURVEY.md
# section 8d; measured by scripts/m0_scan.py on the B200, profiles/r2_m0_scan.jsonl: -0.1 -> 4, -0.3 -> 15, -0.35 -> 36,
# -0.4 -> 149, -0.45 -> 317 per cycle): the "near-physical mass" regime of BASELINE.json configs[3].  `--m0` overrides.
WORKLOADS = {
    "16^3x32-L2": dict(lattice=[32, 16, 16, 16], levels=2, test_vectors=(20,), setup_iter=(3,), m0=-0.35,
                       config="configs[1]: 16^3x32 synthetic random SU(3) gauge field, 2-level AMG"),
    "32^3x64-L3": dict(lattice=[64, 32, 32, 32], levels=3, test_vectors=(20, 28), setup_iter=(3, 2), m0=-0.35,
                       coarse_block=[2, 2, 2, 2],
                       config="configs[2]: 32^3x64 synthetic gauge field, 3-level AMG, mixed float/double"),
    # level-1 blocks 3x2x2x2: lattices 96x48^3 -> 24x12^3 -> 8x6^3
    "48^3x96-L3": dict(lattice=[96, 48, 48, 48], levels=3, test_vectors=(20, 28), setup_iter=(3, 2), m0=-0.35,
                       coarse_block=[3, 2, 2, 2],
                       config="configs[3]: 48^3x96 synthetic gauge field, 3-level AMG near-physical mass"),
    "64^3x128-L3": dict(lattice=[128, 64, 64, 64], levels=3, test_vectors=(20, 28), setup_iter=(3, 2), m0=-0.3,
                        coarse_block=[2, 2, 2, 2],
                        config="configs[4]: 64^3x128 synthetic gauge field, 3-level AMG (needs >= 4 GPUs; single RHS)"),
}
DEFAULT_WORKLOAD = "48^3x96-L3"
CPU_SAMPLE = {2: [16, 8, 8, 8], 3: [16, 16, 16, 16]}
# process grid T x Z per GPU count (48^3x96 on 8 ranks = 4 x 2, local 24 x 24 x 48 x 48; a T-only
# split would leave local T = 12 = three block layers, two of them on the rank boundary)
GRIDS = {1: (1, 1), 2: (2, 1), 4: (4, 1), 8: (4, 2)}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # FFMA lanes x 2 flop x boost clock = 74.4


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
# Parity at the WORKLOAD size, untimed, written into the bench record: the unmodified reference's d_plus_clover_double
# (oracle/_ref, the checker) against the device D_W (double and float) on a seeded random vector, and the residual of the
# timed solve's solution under the reference operator.  D_W is local, so the reference evaluates it exactly on T slabs:
# a slab process builds the reference operator on [t0-2, t0+ti+2+pad) of the same gauge field as a lattice of its own;
# the wrap-around of that sub-lattice only touches the halo slices, the `ti` interior slices see the true neighbours.
# Ranks share their parts of U and x through files (one node), slab j is checked by rank j mod world, a few slab
# processes in parallel per rank (host cores / memory permitting); `coverage` says which fraction of T was checked.
PARITY_TOL = {"dw_double_rel": 1e-12, "dw_float_rel": 1e-5, "residual_ref_operator": 1.5e-10}
PARITY_HALO = 2


def parity_vector_slice(seed, t, shape_zyx):
    rng = np.random.default_rng([seed, int(t)])
    n = int(np.prod(shape_zyx)) * 12
    return (rng.uniform(-0.5, 0.5, n) + 1j * rng.uniform(-0.5, 0.5, n)).reshape(tuple(shape_zyx) + (12,))


def _assemble(d, name, ts, lat, grid, tail):
    """Global T slices `ts` of the field stored per rank as <name>_<rank>.npy with shape [lt][lz][Y][X] + tail."""
    PT, PZ = grid
    lt, lz = lat[0] // PT, lat[1] // PZ
    files = {}
    out = None
    for k, t in enumerate(ts):
        cT, tl = t // lt, t % lt
        for cZ in range(PZ):
            r = cT * PZ + cZ
            if r not in files:
                files[r] = np.load(os.path.join(d, "%s_%d.npy" % (name, r)), mmap_mode="r")
            a = files[r]
            if out is None:
                out = np.empty((len(ts), lat[1], lat[2], lat[3]) + tuple(tail), dtype=a.dtype)
            out[k, cZ * lz:(cZ + 1) * lz] = a[tl]
    return out


def _slab_extent(ti, halo):
    return -(-(ti + 2 * halo) // 8) * 8


def parity_slab_main(spec_path):
    """Subprocess: the reference operator on one T slab (see above).  Writes D_ref v on the interior slices and the partial
    sums of the residual check."""
    with open(spec_path) as f:
        sp = json.load(f)
    from oracle import ref
    lat, grid, d = sp["lat"], tuple(sp["grid"]), sp["dir"]
    T, t0, ti, H = lat[0], sp["t0"], sp["ti"], sp["halo"]
    Ts = _slab_extent(ti, H)                              # reference geometry: 4^4 blocks, even coarse lattice
    assert Ts <= T
    ts = [(t0 - H + k) % T for k in range(Ts)]
    U = _assemble(d, "U", ts, lat, grid, (4, 3, 3, 2))
    x = _assemble(d, "x", ts, lat, grid, (12,))
    v = np.stack([parity_vector_slice(sp["seed"], t, lat[1:]) for t in ts])
    sub = [Ts] + lat[1:]
    R = ref.Reference(sub, [4, 4, 4, 4], levels=2, test_vectors=(4,), setup_iter=(1,), restart=2, max_restart=2,
                      m0=sp["m0"], csw=sp["csw"])
    R.set_conf(U)
    del U
    Dv = R.dw_double(v.reshape(-1)).reshape(v.shape)[H:H + ti]
    Dx = R.dw_double(x.reshape(-1)).reshape(x.shape)[H:H + ti]
    np.save(os.path.join(d, "dref_%d.npy" % sp["slab"]), Dv)
    num = float(np.sum(np.abs(1.0 - Dx) ** 2))           # b = 1 on every site
    with open(os.path.join(d, "slab_%d.json" % sp["slab"]), "w") as f:
        json.dump({"res_num": num, "res_den": float(Dx.size), "sites": int(Dx.size // 12)}, f)


def parity_check(S, w, lat, grid, rank, world, workdir, x_local, dist):
    """Runs on every rank after the timed region.  Returns the parity dict (identical on all ranks)."""
    PT, PZ = grid
    lt, lz = lat[0] // PT, lat[1] // PZ
    cT, cZ = rank // PZ, rank % PZ
    seed = 4242
    per_slice = int(np.prod(lat[1:]))
    ti, halo = 0, PARITY_HALO
    for cand in range(1, lt + 1):                          # largest divisor of lt whose slab stays below ~2M sites
        if lt % cand == 0 and _slab_extent(cand, halo) * per_slice <= 2.2e6 and _slab_extent(cand, halo) <= lat[0]:
            ti = cand
    if ti == 0:
        if world == 1 and lat[0] % 8 == 0:
            ti, halo = lat[0], 0                           # small lattice: the reference runs on all of it
        else:
            return {"skipped": "no T slab of this lattice fits the reference's geometry rules"}
    nslab = lat[0] // ti
    mine = [j for j in range(nslab) if j % world == rank]
    cores = os.cpu_count() or 1
    try:
        with open("/proc/meminfo") as f:
            avail = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] * 1024.0
    except Exception:
        avail = 32e9
    slab_bytes = _slab_extent(ti, halo) * per_slice * 4.5e3
    par = int(max(0, min(len(mine), max(1, cores // (2 * world)), (0.5 * avail / world) // slab_bytes)))
    mine = mine[:par]                                       # one wave of slab processes: bounded time
    v = np.stack([parity_vector_slice(seed, cT * lt + t, lat[1:])[cZ * lz:(cZ + 1) * lz] for t in range(lt)])
    dvd = S.apply_dw(v.reshape(-1)).reshape(v.shape)
    dvf = S.apply_dw(v.reshape(-1), "float").reshape(v.shape)
    del v
    np.save(os.path.join(workdir, "x_%d.npy" % rank), x_local.reshape(lt, lz, lat[2], lat[3], 12))
    if world > 1:
        dist.barrier()
    procs = []
    for j in mine:
        spec = os.path.join(workdir, "spec_%d.json" % j)
        with open(spec, "w") as f:
            json.dump({"dir": workdir, "lat": lat, "grid": [PT, PZ], "t0": j * ti, "ti": ti, "halo": halo, "slab": j, "seed": seed,
                       "m0": w["m0"], "csw": 1.0}, f)
        env = dict(os.environ, OMP_NUM_THREADS=str(max(1, cores // max(1, world * par))))
        procs.append((j, subprocess.Popen([sys.executable, os.path.abspath(__file__), "--parity-slab", spec], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    fails = []
    for j, pr in procs:
        out_, _ = pr.communicate()
        if pr.returncode != 0:
            fails.append("slab %d: %s" % (j, out_[-500:]))
    if world > 1:
        dist.barrier()
    acc = np.zeros(8)
    if not fails:
        for j in range(nslab):                              # every rank compares the slabs that fall into its own part
            fn = os.path.join(workdir, "dref_%d.npy" % j)
            t0 = j * ti
            if t0 // lt != cT or not os.path.exists(fn):
                continue
            ref_ = np.load(fn)[:, cZ * lz:(cZ + 1) * lz]
            tl = t0 % lt
            acc[0] += np.sum(np.abs(ref_ - dvd[tl:tl + ti]) ** 2)
            acc[1] += np.sum(np.abs(ref_ - dvf[tl:tl + ti]) ** 2)
            acc[2] += np.sum(np.abs(ref_) ** 2)
            if cZ == 0:
                with open(os.path.join(workdir, "slab_%d.json" % j)) as f:
                    sj = json.load(f)
                acc[3] += sj["res_num"]; acc[4] += sj["res_den"]; acc[5] += ti
    acc[6] = len(fails)
    if world > 1:
        import torch
        t = torch.tensor(acc, dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        acc = t.cpu().numpy()
    if acc[6] > 0 or acc[5] == 0:
        return {"skipped": "reference slab process failed or no slab fitted the host (%d failures, %s)" % (int(acc[6]), "; ".join(fails)[:300])}
    par_ = {"dw_double_rel": float(np.sqrt(acc[0] / acc[2])), "dw_float_rel": float(np.sqrt(acc[1] / acc[2])),
            "residual_ref_operator": float(np.sqrt(acc[3] / acc[4])),
            "coverage": "%d of %d T slices (slabs of %d + %d halo slices)" % (int(acc[5]), lat[0], ti, 2 * halo),
            "checker": "unmodified reference d_plus_clover_double (oracle/_ref) on the workload's gauge field",
            "tolerances": PARITY_TOL}
    par_["ok"] = all(par_[k] <= PARITY_TOL[k] for k in PARITY_TOL)
    return par_

# ----------------------------------------------------------------------------------------------------------------
def run_native(args, w, name):
    import tempfile
    import shutil
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
    dist = None
    PT, PZ = args.grid if args.grid else GRIDS.get(world, (world, 1))
    if PT * PZ != world or lat[0] % PT or lat[1] % PZ:
        raise RuntimeError("process grid %d x %d does not fit %d ranks / lattice %s" % (PT, PZ, world, lat))
    if world > 1:
        # one process per GPU, lattice partitioned T x Z; torch.distributed only carries the NCCL id, the timing max and the
        # parity sums
        import torch.distributed as dist
        from ddalphaamg_b200.interface import comm_init
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        comm_init()
    lt, lz = lat[0] // PT, lat[1] // PZ
    cT, cZ = rank // PZ, rank % PZ                    # rank = cT * PZ + cZ, T slowest (ghost.c:47-66)
    local = [lt, lz] + lat[2:]
    if world > 1:
        kw["local_lattice"] = local
    V = int(np.prod(local))                           # local sites

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

    # scratch directory shared by the ranks of this node (parity check): rank 0 creates it
    workdir = None
    if not args.no_parity:
        if rank == 0:
            base = "/dev/shm" if shutil.disk_usage("/dev/shm").free > 16e9 + 1e3 * int(np.prod(lat)) else tempfile.gettempdir()
            workdir = tempfile.mkdtemp(prefix="dda_parity_", dir=base)
        if world > 1:
            obj = [workdir]
            dist.broadcast_object_list(obj, src=0)
            workdir = obj[0]

    t0 = time.time()
    U = random_gauge_field(lat, seed=20261018, eps=0.3, t_range=(cT * lt, (cT + 1) * lt))
    if PZ > 1:
        U = np.ascontiguousarray(U[:, cZ * lz:(cZ + 1) * lz])
    t_gauge = time.time() - t0
    S = DDalphaAMG(lat, [4, 4, 4, 4], **kw)
    plaq = S.set_conf(U)
    if workdir:
        np.save(os.path.join(workdir, "U_%d.npy" % rank), U)
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
    # coarsest-level Schur complement (one GMRES iteration's operator): every hop matrix twice, S_ee and Soo^-1 once
    Vl, ncl = S.level_shape(nlev - 1)
    Vl_solve = Vl * world if S.info(INFO.COARSEST_REPLICATED) else Vl
    schur_ms = S.bench_op(BENCH.COARSEST_SCHUR, nlev - 1, 50)
    add("coarsest_schur_n%d" % ncl, schur_ms, Vl_solve * (2 * 4 + 1) * ncl * ncl * 8.0)

    # ---- 12 right-hand sides through the tensor-core kernel (single rank: the entry point takes whole-lattice
    # vectors).  Checked against the single-RHS kernel on one column, timed next to 12 single-RHS applications.
    tensor_core = None
    if world == 1 and nlev > 1:
        try:
            rng = np.random.default_rng(5)
            Vc, nc = S.level_shape(1)
            vs = (rng.standard_normal((12, Vc * nc)) + 1j * rng.standard_normal((12, Vc * nc))).astype(np.complex64)
            o12, ms12 = S.level_apply_mrhs(1, vs, reps=reps)
            one = S.level_apply(1, vs[5])
            single_ms = ops["coarse_apply_d1_n%d" % nc]["ms"]
            tb = 9 * nc * nc * 8.0 * Vc + 12 * 2 * nc * 8.0 * Vc            # operator images once + 12 vectors in and out
            fl = 12 * 9 * 8.0 * nc * nc * Vc                                # complex multiply-adds of 9 blocks x 12 columns
            tensor_core = {"kernel": "mrhs::k_coarse_mrhs: D_c on 12 right-hand sides, tcgen05.mma kind::tf32 (TF32x3 split), TMEM accumulators",
                           "depth": 1, "n": nc, "sites": Vc, "ms_per_12rhs_apply": ms12, "ms_single_rhs_apply": single_ms,
                           "speedup_vs_12_single_rhs_applies": 12 * single_ms / ms12,
                           "relerr_column5_vs_single_rhs_kernel": float(np.linalg.norm(o12[5] - one) / np.linalg.norm(one)),
                           "algorithmic_bytes": tb, "gbs": tb / (ms12 * 1e-3) / 1e9, "frac_of_peak": tb / (ms12 * 1e-3) / 1e9 / peak,
                           "fp32_equivalent_tflops": fl / (ms12 * 1e-3) / 1e12}
            del vs, o12, one
        except Exception as e:      # a shape the kernel does not take: the record says so, the benchmark goes on
            tensor_core = {"error": repr(e)}

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
    coarsest_ms_it = 1e3 * share["coarsest_solve"] / max(1, its[1])

    # ---- parity at the workload size with the reference's operator (untimed)
    parity = {"skipped": "--no-parity"}
    if workdir:
        S.solve_device(b)
        xs = S.download_solution()
        parity = parity_check(S, w, lat, (PT, PZ), rank, world, workdir, xs, dist)
        del xs
        barrier()
        if rank == 0:
            shutil.rmtree(workdir, ignore_errors=True)
    S.free()

    # dominant kernel of the solve: the fused fine-level SAP block visit (one launch = the block visits of one colour; a
    # smoother call = 2 iterations x 2 colours = 4 launches).  Algorithmic bytes per block visit:
    # 256*(288+336) + 5*256*96 = 282 624 B.  DRAM traffic per visit: ncu --set full capture of THIS workload, recorded in
    # profiles/r2_traffic.json (null when no capture of the workload exists).  Flops per visit: 6 block-operator
    # applications x 256 sites x ~1600 flop = 2.46 MFLOP.
    dom = ops["sap_smoother_d0"]
    visits_per_launch = nblk / 2.0
    launch_ms = dom["ms"] / 4.0
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f).get("k_sap_fine2", {})
        if tj.get("workload") == name and tj.get("n_gpus", 1) == world:
            traffic, traffic_src = tj["dram_bytes_per_block_visit"] * visits_per_launch, tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    flops_visit = 6 * 256 * 1600.0
    tflops = flops_visit * visits_per_launch / (launch_ms * 1e-3) / 1e12
    roof = {"kernel": "k_sap_fine2: fused fine-level SAP block solve, one CTA per 4^4 Schwarz block (%d block visits per launch, "
                      "avg launch %.3f ms from CUDA events over %d launches)" % (int(visits_per_launch), launch_ms, 4 * 5),
            "bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak,
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 282624.0 * visits_per_launch,
            "peak_source": peak_src,
            "fp32": {"achieved_tflops": tflops, "peak_tflops": FP32_PEAK_TFLOPS, "frac": tflops / FP32_PEAK_TFLOPS,
                     "note": "arithmetic intensity 8.7 flop/B puts the kernel on the fp32-issue side of the ridge: both fractions are reported"},
            "note": "D_W and coarse-operator GB/s (the metric's named kernels) are in `operators`"}

    # ---- north-star target configuration, driver visible: 64^3 x 128 on the full 8 x B200 box, with its own parity record
    target = None
    if world == 8 and name == DEFAULT_WORKLOAD and not args.no_target:
        tw = dict(WORKLOADS["64^3x128-L3"])
        tw["mixed_precision"] = w.get("mixed_precision", 2)
        tlat = tw["lattice"]
        tkw = solver_kwargs(tw)
        tl_t, tl_z = tlat[0] // PT, tlat[1] // PZ
        tlocal = [tl_t, tl_z] + tlat[2:]
        tkw["local_lattice"] = tlocal
        twork = None
        if not args.no_parity:
            if rank == 0:
                twork = tempfile.mkdtemp(prefix="dda_parity_t_", dir=tempfile.gettempdir())
            obj = [twork]
            dist.broadcast_object_list(obj, src=0)
            twork = obj[0]
        U = random_gauge_field(tlat, seed=20261018, eps=0.3, t_range=(cT * tl_t, (cT + 1) * tl_t))
        if PZ > 1:
            U = np.ascontiguousarray(U[:, cZ * tl_z:(cZ + 1) * tl_z])
        T_ = DDalphaAMG(tlat, [4, 4, 4, 4], **tkw)
        tplaq = T_.set_conf(U)
        if twork:
            np.save(os.path.join(twork, "U_%d.npy" % rank), U)
        del U
        t0 = time.time()
        T_.setup(tw["setup_iter"][0])
        tsetup = time.time() - t0
        tb = np.ones(int(np.prod(tlocal)) * 12, dtype=np.complex128)
        for _ in range(2):
            T_.solve_device(tb)
        barrier()
        tms = 0.0
        for _ in range(3):
            tres, tst, ms_ = T_.solve_device(tb)
            tms += ms_
        barrier()
        tsec = rank_max(tms) / 3e3
        te = time.time()
        T_.solve(tb)
        barrier()
        te2e = rank_max(time.time() - te)
        tpar = {"skipped": "--no-parity"}
        if twork:
            T_.solve_device(tb)
            txs = T_.download_solution()
            tpar = parity_check(T_, tw, tlat, (PT, PZ), rank, world, twork, txs, dist)
            del txs
            barrier()
            if rank == 0:
                shutil.rmtree(twork, ignore_errors=True)
        T_.free()
        target = {"workload": "64^3x128-L3", "detail": tw["config"], "value": tsec, "unit": "s", "e2e": te2e, "n_gpus": world,
                  "local_lattice_TZYX": tlocal, "m0": tw["m0"], "iterations": [int(tst[0]), int(tst[1])], "residual": tres,
                  "plaquette": tplaq, "setup_seconds_untimed": tsetup, "steps": 3, "parity": tpar}

    if world > 1:
        from ddalphaamg_b200.interface import comm_finalize
        comm_finalize()
        dist.barrier()
        dist.destroy_process_group()

    out = {"metric": METRIC, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * sec, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64 restarts + f32 Arnoldi/cycle (mixed precision %d)" % w.get("mixed_precision", 2), "data": "synthetic",
           "config": {"workload": name, "detail": w["config"], "lattice_TZYX": lat, "local_lattice_TZYX": local,
                      "partition": "process grid %d x %d (T x Z) over %d GPU(s), NCCL send/recv halos + allreduce, coarsest level "
                                   "gathered and solved on every rank" % (PT, PZ, world), "levels": w["levels"],
                      "test_vectors": list(w["test_vectors"]), "mixed_precision": w.get("mixed_precision", 2), "m0": w["m0"], "csw": 1.0, "tol": 1e-10,
                      "gauge": "U=exp(i*0.3*H), H Gaussian traceless Hermitian, seed 20261018, plaquette %.6f" % plaq,
                      "l2": "working set %.1f GB per solve >> 126 MB L2 (inputs larger than L2, no flush)" % (dev_bytes / 1e9),
                      "iterations": its, "coarsest_iterations_per_cycle": its[1] / max(1, its[0]),
                      "coarsest_ms_per_iteration": coarsest_ms_it, "setup_seconds_untimed": t_setup},
           "e2e": {"value": e2e, "unit": "s", "h2d_bytes_per_step": n * 16 * world, "d2h_bytes_per_step": n * 16 * world},
           "gpu_launches": launches, "clocks": clocks, "roofline": roof, "operators": ops, "parity": parity,
           "time_share_seconds_profiled_solve": share, "wall_seconds_timed_region": wall}
    if target is not None:
        out["target_64c128"] = target
    if tensor_core is not None:
        out["tensor_core_12rhs"] = tensor_core

    if rank == 0 and world == 1 and not args.no_cpu:
        # the reference runs in its own process (it aborts the process on any error, main.h:424-439)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
                                "--workload", name, "--m0", str(w["m0"])], capture_output=True, text=True, timeout=900)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            out["cpu_baseline"] = json.loads(line)["cpu_baseline"]
        except Exception as e:  # the baseline is reported, never required for the product number
            out["cpu_baseline"] = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}
        # measured pair: the GPU library on the CPU arm's own sample lattice (same field, same parameters) -- both numbers
        # really run, no volume scaling
        try:
            cb = out["cpu_baseline"]
            slat = CPU_SAMPLE[w["levels"]]
            kws = solver_kwargs(w)
            if w["levels"] > 2:
                kws["coarse_block"] = [2, 2, 2, 2]
            S2 = DDalphaAMG(slat, [4, 4, 4, 4], **kws)
            S2.set_conf(random_gauge_field(slat, seed=20261018, eps=0.3))
            S2.setup(w["setup_iter"][0])
            bs_ = np.ones(S2.V * 12, dtype=np.complex128)
            for _ in range(3):
                S2.solve(bs_)
            ts0 = time.time()
            for _ in range(5):
                _, rs_, sts_ = S2.solve(bs_)
            gsec = (time.time() - ts0) / 5
            S2.free()
            out["sample_pair"] = {"lattice_TZYX": slat, "gpu_seconds_e2e": gsec, "gpu_iterations": [int(sts_[0]), int(sts_[1])],
                                  "cpu_seconds": cb.get("sample_seconds"), "cpu_iterations": cb.get("sample_iterations"),
                                  "cpu_cores": cb.get("cores"), "sample_ratio": (cb.get("sample_seconds") or 0) / gsec,
                                  "note": "same lattice, same field recipe, same solver parameters, both arms measured on this box"}
        except Exception as e:
            out["sample_pair"] = {"failed": repr(e)}
    if rank == 0:
        print(json.dumps(out))
    for par_ in (parity, (target or {}).get("parity")):
        if isinstance(par_, dict) and par_.get("ok") is False:
            raise RuntimeError("parity against the reference operator failed: %s" % json.dumps(par_))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--m0", type=float, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed reference-operator parity check")
    ap.add_argument("--no-target", action="store_true", help="N = 8 only: skip the extra 64^3x128 target-configuration block")
    ap.add_argument("--grid", type=lambda v: tuple(int(q) for q in v.split("x")), default=None, help="process grid TxZ, e.g. 4x2")
    ap.add_argument("--parity-slab", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--mixed-precision", type=int, default=2, choices=[1, 2],
                    help="reference parameter `mixed precision` (both arms): 2 = fgmres_MP, the reference's default "
                         "(init.c:581-962); 1 = double outer FGMRES")
    args = ap.parse_args()
    if args.parity_slab:
        parity_slab_main(args.parity_slab)
        return
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

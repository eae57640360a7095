"""ctypes host side of libdd_alpha_amg.so -- mirrors the reference's C library interface.

Reference: include/dd_alpha_amg.h:29-83 (dd_alpha_amg_par, dd_alpha_amg_init/set_conf/setup/setup_update/
wilson_solve/free), include/dd_alpha_amg_parameters.h:25-51, parameter file keys src/init.c:592-962.
"""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
STRINGLENGTH = 500
MAX_MG_LEVELS = 4


class INFO:
    NUM_LEVELS, SITES, SITE_VARS, TEST_VECTORS, BLOCK_SITES, NUM_BLOCKS, EMULATION, COARSEST_REPLICATED = range(8)


class OPT:
    USE_FAST, PROFILE, SEED, PRINT = range(4)


class STAT:
    LAUNCHES, DEVICE_BYTES, PLAQUETTE, ITER, COARSE_ITER, T_COARSEST, T_RESTRICT, T_INTERPOLATE = range(8)
    T_SMOOTH0 = 10
    T_OP0 = 20


class OP:
    APPLY, RESTRICT, INTERPOLATE, SMOOTHER, VCYCLE, COARSEST_SOLVE = range(6)


class BENCH:
    DW_DOUBLE, DW_FLOAT, LEVEL_APPLY, RESTRICT, INTERPOLATE, SMOOTHER, VCYCLE, COARSEST_SCHUR = range(8)


CONF_INDEX_FCT = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
VECTOR_INDEX_FCT = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
GLOBAL_TIME_FCT = C.CFUNCTYPE(C.c_int, C.c_int)


class AmgParameters(C.Structure):
    """struct dd_alpha_amg_parameters (reference include/dd_alpha_amg_parameters.h:26-51)."""
    _fields_ = [("number_of_levels", C.c_int),
                ("global_lattice", (C.c_int * 4) * MAX_MG_LEVELS),
                ("local_lattice", (C.c_int * 4) * MAX_MG_LEVELS),
                ("block_lattice", (C.c_int * 4) * MAX_MG_LEVELS),
                ("mg_basis_vectors", C.c_int * MAX_MG_LEVELS),
                ("setup_iterations", C.c_int * MAX_MG_LEVELS),
                ("discard_setup_after", C.c_int),
                ("update_setup_iterations", C.c_int * MAX_MG_LEVELS),
                ("update_setup_after", C.c_int),
                ("post_smooth_iterations", C.c_int * MAX_MG_LEVELS),
                ("post_smooth_block_iterations", C.c_int * MAX_MG_LEVELS),
                ("coarse_grid_iterations", C.c_int),
                ("coarse_grid_maximum_number_of_restarts", C.c_int),
                ("coarse_grid_tolerance", C.c_double),
                ("solver_mass", C.c_double),
                ("setup_mass", C.c_double),
                ("c_sw", C.c_double)]


class Par(C.Structure):
    """dd_alpha_amg_par (reference include/dd_alpha_amg.h:29-39), passed by value."""
    _fields_ = [("param_file_path", C.c_char * STRINGLENGTH),
                ("conf_index_fct", CONF_INDEX_FCT),
                ("vector_index_fct", VECTOR_INDEX_FCT),
                ("global_time", GLOBAL_TIME_FCT),
                ("bc", C.c_int),
                ("m0", C.c_double),
                ("csw", C.c_double),
                ("setup_m0", C.c_double),
                ("amg_params", AmgParameters)]


EXPORTS = ["dd_alpha_amg_init", "dd_alpha_amg_init_external_threading", "dd_alpha_amg_get_gauge_pointer",
           "dd_alpha_amg_get_clover_pointer", "dd_alpha_amg_fields_updated", "dd_alpha_amg_set_conf",
           "dd_alpha_amg_update_parameters", "dd_alpha_amg_setup", "dd_alpha_amg_setup_external_threading",
           "dd_alpha_amg_setup_update", "dd_alpha_amg_setup_update_external_threading", "dd_alpha_amg_wilson_solve",
           "dd_alpha_amg_preconditioner", "dd_alpha_amg_preconditioner_external_threading", "dd_alpha_amg_free",
           "DDalphaAMG_initialize", "DDalphaAMG_update_parameters", "DDalphaAMG_setup", "DDalphaAMG_solve",
           "DDalphaAMG_finalize",
           "dda_info", "dda_set_option", "dda_get_stat", "dda_reset_stats", "dda_apply_dw", "dda_get_operator",
           "dda_set_interpolation", "dda_get_interpolation", "dda_level_op", "dda_bench_op", "dda_upload_source",
           "dda_solve_device", "dda_download_solution", "dda_level_apply_mrhs", "dda_write_test_vectors", "dda_read_test_vectors",
           "dda_setup_if_necessary",
           "dda_comm_unique_id", "dda_comm_init", "dda_comm_finalize", "dda_comm_rank", "dda_comm_size",
           "dda_comm_init_callbacks", "dda_is_emulation"]

SENDRECV_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int)
ALLREDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int)


def library_path():
    """The CUDA library.  DDA_LIBRARY (a path) selects another CUDA build of the same sources for A/B measurements."""
    return os.environ.get("DDA_LIBRARY") or os.path.join(_HERE, "libdd_alpha_amg.so")


_LIBS = {}


def load_library(path=None):
    """Loads the C-ABI library.  Default = the CUDA build; raises if it is missing (no fallback of any kind)."""
    path = path or library_path()
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise RuntimeError("%s not found: build the CUDA library with `python -m ddalphaamg_b200.build` "
                           "(there is no CPU fallback)" % path)
    if os.path.basename(os.path.dirname(os.path.abspath(path))) != "_emu":
        # The CUDA library links the system NCCL; torch bundles a newer one.  If torch is going to be used in this process
        # (bench.py, the multi-rank tests) it has to be imported BEFORE the library, or libtorch_cuda fails to resolve its
        # NCCL symbols.  A C caller without Python is not affected.
        try:
            import torch  # noqa: F401
        except ImportError:
            pass
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    L.dda_is_emulation.restype = C.c_int
    if L.dda_is_emulation() and os.path.basename(os.path.dirname(os.path.abspath(path))) != "_emu":
        # the g++ host-emulation build of the host logic is test infrastructure (tests/_emu); it is never a product path
        raise RuntimeError("%s is a host-emulation test build; the product library is %s" % (path, library_path()))
    dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.dd_alpha_amg_init.argtypes = [Par]
    L.dd_alpha_amg_init_external_threading.argtypes = [Par, C.c_int, C.c_int]
    L.dd_alpha_amg_set_conf.argtypes = [dp]
    L.dd_alpha_amg_set_conf.restype = C.c_double
    L.dd_alpha_amg_get_gauge_pointer.restype = dp
    L.dd_alpha_amg_get_clover_pointer.restype = dp
    L.dd_alpha_amg_update_parameters.argtypes = [C.POINTER(AmgParameters)]
    L.dd_alpha_amg_setup.argtypes = [C.c_int, ip]
    L.dd_alpha_amg_setup_update.argtypes = [C.c_int, ip]
    L.dd_alpha_amg_wilson_solve.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, ip]
    L.dd_alpha_amg_wilson_solve.restype = C.c_double
    L.dd_alpha_amg_preconditioner.argtypes = [dp, dp, C.c_double, C.c_double, ip]
    L.DDalphaAMG_solve.argtypes = [dp, dp, C.c_double, ip]
    L.DDalphaAMG_solve.restype = C.c_double
    L.dda_info.argtypes = [C.c_int, C.c_int]
    L.dda_info.restype = C.c_int
    L.dda_set_option.argtypes = [C.c_int, C.c_double]
    L.dda_get_stat.argtypes = [C.c_int]
    L.dda_get_stat.restype = C.c_double
    L.dda_apply_dw.argtypes = [C.c_int, dp, dp]
    L.dda_get_operator.argtypes = [dp, dp]
    L.dda_set_interpolation.argtypes = [C.c_int, fp]
    L.dda_get_interpolation.argtypes = [C.c_int, fp]
    L.dda_level_op.argtypes = [C.c_int, C.c_int, fp, fp, C.c_int, C.c_int]
    L.dda_bench_op.argtypes = [C.c_int, C.c_int, C.c_int]
    L.dda_bench_op.restype = C.c_double
    L.dda_upload_source.argtypes = [dp]
    L.dda_solve_device.argtypes = [C.c_double, ip, dp]
    L.dda_solve_device.restype = C.c_double
    L.dda_download_solution.argtypes = [dp]
    L.dda_write_test_vectors.argtypes = [C.c_char_p]
    L.dda_read_test_vectors.argtypes = [C.c_char_p]
    L.dda_setup_if_necessary.restype = C.c_int
    L.dda_level_apply_mrhs.argtypes = [C.c_int, fp, fp, C.c_int, dp]
    L.dda_level_apply_mrhs.restype = C.c_int
    L.dda_comm_unique_id.argtypes = [C.c_char_p, C.c_int]
    L.dda_comm_unique_id.restype = C.c_int
    L.dda_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.dda_comm_init_callbacks.argtypes = [C.c_int, C.c_int, SENDRECV_FN, ALLREDUCE_FN]
    _LIBS[path] = L
    return L


_COMM_KEEPALIVE = []


def comm_init(lib=None, device=None):
    """Binds this process (one per GPU) into the library's communicator.  torch.distributed must be initialised: it
    only carries the NCCL unique id from rank 0 to the others (what MPI_Bcast does for an MPI caller).  With the
    host-emulation test library the exchanges themselves are delegated to torch.distributed (gloo)."""
    import torch
    import torch.distributed as dist
    L = load_library(lib)
    rank, size = dist.get_rank(), dist.get_world_size()
    if L.dda_is_emulation():
        def sendrecv(send, recv, nbytes, to, frm):
            src = torch.frombuffer((C.c_char * nbytes).from_address(send), dtype=torch.uint8).clone()
            dst = torch.empty(nbytes, dtype=torch.uint8)
            if to == rank and frm == rank:
                dst.copy_(src)
            else:
                reqs = [dist.isend(src, to), dist.irecv(dst, frm)]
                for r in reqs:
                    r.wait()
            C.memmove(recv, dst.data_ptr(), nbytes)

        def allreduce(buf, n):
            t = torch.frombuffer((C.c_double * n).from_address(C.addressof(buf.contents)), dtype=torch.float64)
            dist.all_reduce(t)

        cbs = (SENDRECV_FN(sendrecv), ALLREDUCE_FN(allreduce))
        _COMM_KEEPALIVE.append(cbs)
        L.dda_comm_init_callbacks(rank, size, cbs[0], cbs[1])
        return rank, size
    buf = C.create_string_buffer(128)
    if rank == 0:
        n = L.dda_comm_unique_id(buf, 128)
        assert n > 0
    obj = [buf.raw if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    L.dda_comm_init(rank, size, obj[0], int(device))
    return rank, size


def comm_finalize(lib=None):
    load_library(lib).dda_comm_finalize()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def write_ini(path, lattice, block, levels=2, test_vectors=(20, 28), setup_iter=(4, 3), post_smooth=(2, 2),
              block_iter=(4, 4), m0=-0.5, csw=1.0, tol=1e-10, restart=50, max_restart=20, coarse_tol=5e-2,
              coarse_iter=100, coarse_restart=5, mixed_precision=1, anti_pbc=1, method=2, kcycle=1,
              coarse_lattice=None, coarse_block=None, odd_even=1, local_lattice=None, ncycle=(1, 1), relax=(1.0, 1.0),
              interpolation=2, tv_file=None):
    """Parameter file in the reference's "key: value" format (keys: src/init.c:592-962, sample.ini)."""
    loc = local_lattice or lattice
    lines = ["configuration: none", "format: 0", "right hand side: 0",
             "antiperiodic boundary conditions: %d" % anti_pbc, "number of levels: %d" % levels,
             "number of openmp threads: 1",
             "d0 global lattice: %d %d %d %d" % tuple(lattice), "d0 local lattice: %d %d %d %d" % tuple(loc),
             "d0 block lattice: %d %d %d %d" % tuple(block)]
    for d in range(max(levels - 1, 1)):
        lines += ["d%d post smooth iter: %d" % (d, post_smooth[min(d, len(post_smooth) - 1)]),
                  "d%d block iter: %d" % (d, block_iter[min(d, len(block_iter) - 1)]),
                  "d%d test vectors: %d" % (d, test_vectors[min(d, len(test_vectors) - 1)]),
                  "d%d setup iter: %d" % (d, setup_iter[min(d, len(setup_iter) - 1)]),
                  "d%d preconditioner cycles: %d" % (d, ncycle[min(d, len(ncycle) - 1)]),
                  "d%d relaxation factor: %.16g" % (d, relax[min(d, len(relax) - 1)])]
    if levels > 2:
        cl = coarse_lattice or [a // b for a, b in zip(lattice, block)]
        cloc = [c * l // g for c, l, g in zip(cl, loc, lattice)]
        lines += ["d1 global lattice: %d %d %d %d" % tuple(cl), "d1 local lattice: %d %d %d %d" % tuple(cloc)]
        if coarse_block is not None:
            lines += ["d1 block lattice: %d %d %d %d" % tuple(coarse_block)]
    lines += ["m0: %.16g" % m0, "csw: %.16g" % csw, "tolerance for relative residual: %g" % tol,
              "iterations between restarts: %d" % restart, "maximum of restarts: %d" % max_restart,
              "coarse grid tolerance: %g" % coarse_tol, "coarse grid iterations: %d" % coarse_iter,
              "coarse grid restarts: %d" % coarse_restart, "print mode: 0", "method: %d" % method,
              "mixed precision: %d" % mixed_precision, "randomize test vectors: 0",
              "odd even preconditioning: %d" % odd_even, "kcycle: %d" % kcycle, "kcycle length: 5",
              "kcycle restarts: 2", "kcycle tolerance: 1E-1", "interpolation: %d" % interpolation]
    if tv_file is not None:
        lines += ["test vector io from single file: 0", "test vector io file name: %s" % tv_file]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def read_conf(path, anti_pbc=True):
    """Reference native configuration format (src/io.c:486-506): 4 x int32 [T,Z,Y,X], one double (plaquette), then
    doubles [t][z][y][x][mu=T,Z,Y,X][3][3][re,im].  anti_pbc flips the time links of the last slice, which is what
    the reference's reader does (io.c:535-541) and what dd_alpha_amg_set_conf expects from its caller."""
    with open(path, "rb") as f:
        dims = np.fromfile(f, dtype=np.int32, count=4)
        plaq = float(np.fromfile(f, dtype=np.float64, count=1)[0])
        n = int(np.prod(dims)) * 72
        data = np.fromfile(f, dtype=np.float64, count=n)
    U = data.reshape(tuple(int(d) for d in dims) + (4, 3, 3, 2)).copy()
    if anti_pbc:
        U[-1, :, :, :, 0] *= -1.0
    return [int(d) for d in dims], plaq, U


def _mm3(A, B):
    """3x3 complex matrix product on component arrays (lists of 9 arrays of equal length)."""
    return [A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j] for i in range(3) for j in range(3)]


def _gauge_chunk(args):
    """exp(i eps H) for one chunk of links by scaling (2^-3) and squaring of a 10th-order Taylor polynomial, all
    arithmetic as elementwise array operations (unitary to ~1e-15).  The Gaussian numbers always come from numpy
    (same field everywhere); the arithmetic runs through torch on the GPU when one is present (input plumbing)."""
    ss, n, eps, use_cuda = args
    rng = np.random.default_rng(ss)
    g = rng.standard_normal((9, n))
    if use_cuda:
        import torch
        g = torch.from_numpy(g).cuda()
    tr = (g[0] + g[1] + g[2]) / 3.0
    r = float(np.sqrt(0.5))
    H = [None] * 9
    H[0], H[4], H[8] = g[0] - tr + 0j, g[1] - tr + 0j, g[2] - tr + 0j
    H[1] = r * (g[3] + 1j * g[4]); H[3] = H[1].conj()
    H[2] = r * (g[5] + 1j * g[6]); H[6] = H[2].conj()
    H[5] = r * (g[7] + 1j * g[8]); H[7] = H[5].conj()
    X = [(1j * eps / 8.0) * h for h in H]
    R = [x / 10.0 for x in X]
    for k in (0, 4, 8):
        R[k] = R[k] + 1.0
    for order in range(9, 0, -1):
        R = _mm3(X, R)
        R = [x / float(order) for x in R]
        for k in (0, 4, 8):
            R[k] = R[k] + 1.0
    for _ in range(3):
        R = _mm3(R, R)
    if use_cuda:
        import torch
        return torch.stack(R, dim=1).cpu().numpy().reshape(n, 3, 3)
    return np.stack(R, axis=1).reshape(n, 3, 3)


def random_gauge_field(lattice, seed=20261018, eps=0.3, anti_pbc=True, t_range=None):
    """Deterministic synthetic SU(3) field U = exp(i eps H), H Gaussian traceless Hermitian (eps -> inf: "hot").
    Shape [T][Z][Y][X][4][3][3][2] doubles, the reference's native order.  Chunks of 2^16 links carry independent
    streams spawned from `seed` (same field for any thread count); chunks are generated by a thread pool.
    t_range=(t0, t1): only the time slices [t0, t1) of the same global field (one rank's part)."""
    from concurrent.futures import ThreadPoolExecutor
    T = int(lattice[0])
    t0, t1 = (0, T) if t_range is None else (int(t_range[0]), int(t_range[1]))
    per_t = int(np.prod(lattice[1:])) * 4
    n = T * per_t
    a0, a1 = t0 * per_t, t1 * per_t
    chunk = 1 << 16
    nch = (n + chunk - 1) // chunk
    seeds = np.random.SeedSequence(seed).spawn(nch)
    use_cuda = False
    if a1 - a0 > (1 << 20):
        try:
            import torch
            use_cuda = torch.cuda.is_available()
        except ImportError:
            pass
    c0, c1 = a0 // chunk, (a1 + chunk - 1) // chunk
    jobs = [(seeds[i], min(chunk, n - i * chunk), eps, use_cuda) for i in range(c0, c1)]
    U = np.empty((a1 - a0, 3, 3, 2), dtype=np.float64)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        for i, blk in zip(range(c0, c1), ex.map(_gauge_chunk, jobs)):
            lo, hi = max(a0, i * chunk), min(a1, i * chunk + blk.shape[0])
            U[lo - a0:hi - a0, :, :, 0] = blk[lo - i * chunk:hi - i * chunk].real
            U[lo - a0:hi - a0, :, :, 1] = blk[lo - i * chunk:hi - i * chunk].imag
    U = U.reshape((t1 - t0,) + tuple(lattice[1:]) + (4, 3, 3, 2))
    if anti_pbc and t1 == T:
        U[-1, :, :, :, 0] *= -1.0
    return U


class DDalphaAMG:
    """One live solver instance (the library keeps process-global state like the reference: one per process)."""

    def __init__(self, lattice, block, m0=-0.5, csw=1.0, bc=2, setup_m0=None, lib=None, ini_path=None, **ini_kw):
        """lattice: global lattice T,Z,Y,X; local_lattice (keyword): this rank's part (multi-GPU, after comm_init).
        All host vectors of this object are the rank's LOCAL lexicographic parts."""
        self.L = load_library(lib)
        self.global_lattice = [int(x) for x in lattice]
        self.lattice = [int(x) for x in (ini_kw.get("local_lattice") or lattice)]
        self.V = int(np.prod(self.lattice))
        self._tmp = None
        if ini_path is None:
            self._tmp = tempfile.NamedTemporaryFile(suffix=".ini", delete=False)
            self._tmp.close()
            ini_path = write_ini(self._tmp.name, lattice, block, m0=m0, csw=csw, anti_pbc=1 if bc == 2 else 0, **ini_kw)
        p = Par()
        p.param_file_path = ini_path.encode()
        LT, LZ, LY, LX = self.lattice
        # the caller's layouts: native conf order / lexicographic spinors (offsets in doubles)
        self._cf = CONF_INDEX_FCT(lambda t, z, y, x, mu: 18 * (4 * (x + LX * (y + LY * (z + LZ * t))) + mu))
        self._vf = VECTOR_INDEX_FCT(lambda t, z, y, x: 24 * (x + LX * (y + LY * (z + LZ * t))))
        self._gt = GLOBAL_TIME_FCT(lambda t: t)
        # NULL index functions select the same default layouts inside the library without a Python callback per site
        p.conf_index_fct = CONF_INDEX_FCT(0)
        p.vector_index_fct = VECTOR_INDEX_FCT(0)
        p.global_time = GLOBAL_TIME_FCT(0)
        p.bc = bc
        p.m0 = m0
        p.csw = csw
        p.setup_m0 = m0 if setup_m0 is None else setup_m0
        self.L.dd_alpha_amg_init(p)
        self.emulated = bool(self.L.dda_info(INFO.EMULATION, 0))

    @classmethod
    def from_struct(cls, lattice, block, levels=2, test_vectors=(20, 28), setup_iter=(4, 3), post_smooth=(2, 2),
                    block_iter=(4, 4), m0=-0.5, csw=1.0, bc=2, coarse_iter=100, coarse_restart=5, coarse_tol=5e-2,
                    coarse_lattice=None, coarse_block=None, lib=None):
        """The struct route dd_alpha_amg_init_external_threading (include/dd_alpha_amg.h:44, src/dd_alpha_amg.c:135-173):
        geometry from dd_alpha_amg_parameters, lattice arrays in X,Y,Z,T order (reversed inside, init.c:821-823)."""
        self = cls.__new__(cls)
        self.L = load_library(lib)
        self.global_lattice = [int(x) for x in lattice]
        self.lattice = list(self.global_lattice)
        self.V = int(np.prod(self.lattice))
        self._tmp = None
        p = Par()
        ap = p.amg_params
        ap.number_of_levels = levels
        lat = list(lattice)
        blk = list(block)
        for i in range(levels):
            for m in range(4):
                ap.global_lattice[i][3 - m] = lat[m]
                ap.local_lattice[i][3 - m] = lat[m]
                ap.block_lattice[i][3 - m] = blk[m] if i < levels - 1 else 1
            ap.mg_basis_vectors[i] = test_vectors[min(i, len(test_vectors) - 1)]
            ap.setup_iterations[i] = setup_iter[min(i, len(setup_iter) - 1)]
            ap.post_smooth_iterations[i] = post_smooth[min(i, len(post_smooth) - 1)]
            ap.post_smooth_block_iterations[i] = block_iter[min(i, len(block_iter) - 1)]
            if i < levels - 1:
                lat = coarse_lattice if (i == 0 and coarse_lattice) else [a // b for a, b in zip(lat, blk)]
                blk = coarse_block or [2, 2, 2, 2]
        ap.coarse_grid_iterations = coarse_iter
        ap.coarse_grid_maximum_number_of_restarts = coarse_restart
        ap.coarse_grid_tolerance = coarse_tol
        ap.solver_mass = m0
        ap.setup_mass = m0
        ap.c_sw = csw
        p.conf_index_fct = CONF_INDEX_FCT(0)
        p.vector_index_fct = VECTOR_INDEX_FCT(0)
        p.global_time = GLOBAL_TIME_FCT(0)
        p.bc, p.m0, p.csw, p.setup_m0 = bc, m0, csw, m0
        self._par = p
        self.L.dd_alpha_amg_init_external_threading(p, 1, 1)
        self.emulated = bool(self.L.dda_info(INFO.EMULATION, 0))
        return self

    def update_parameters(self, solver_mass, post_smooth=(2, 2, 2, 2), block_iter=(4, 4, 4, 4), setup_iter=(4, 3, 2, 2)):
        """dd_alpha_amg_update_parameters (include/dd_alpha_amg.h:56, init.c:1139-1148): iteration counts + mass shift."""
        ap = AmgParameters()
        for i in range(MAX_MG_LEVELS):
            ap.post_smooth_iterations[i] = post_smooth[i]
            ap.post_smooth_block_iterations[i] = block_iter[i]
            ap.setup_iterations[i] = setup_iter[i]
        ap.solver_mass = solver_mass
        self.L.dd_alpha_amg_update_parameters(C.byref(ap))

    def gauge_pointer(self):
        """numpy view of the host mirror of D = U/2 (dd_alpha_amg_get_gauge_pointer, dirac.c:171-176)."""
        ptr = self.L.dd_alpha_amg_get_gauge_pointer()
        return np.ctypeslib.as_array(ptr, shape=(self.V * 72,))

    def fields_updated(self):
        self.L.dd_alpha_amg_fields_updated()

    # ---- reference interface
    def set_conf(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.size == self.V * 72
        return self.L.dd_alpha_amg_set_conf(_dp(U))

    def setup(self, iterations):
        st = np.zeros(2, dtype=np.int32)
        self.L.dd_alpha_amg_setup(int(iterations), _ip(st))
        return st

    def setup_update(self, iterations):
        st = np.zeros(2, dtype=np.int32)
        self.L.dd_alpha_amg_setup_update(int(iterations), _ip(st))
        return st

    def solve(self, b, tol=1e-10, scale_even=1.0, scale_odd=1.0, out=None):
        b = np.ascontiguousarray(b, dtype=np.complex128)
        x = np.zeros_like(b) if out is None else out
        st = np.zeros(2, dtype=np.int32)
        res = self.L.dd_alpha_amg_wilson_solve(_dp(x), _dp(b), tol, scale_even, scale_odd, _ip(st))
        return x, res, st

    def preconditioner(self, b):
        b = np.ascontiguousarray(b, dtype=np.complex128)
        x = np.zeros_like(b)
        st = np.zeros(2, dtype=np.int32)
        self.L.dd_alpha_amg_preconditioner(_dp(x), _dp(b), 1.0, 1.0, _ip(st))
        return x

    def free(self):
        self.L.dd_alpha_amg_free()
        if self._tmp is not None:
            try:
                os.unlink(self._tmp.name)
            except OSError:
                pass

    # ---- operator-level entry points (include/dd_alpha_amg_b200.h)
    def info(self, what, depth=0):
        return self.L.dda_info(what, depth)

    def set_option(self, what, value):
        self.L.dda_set_option(what, float(value))

    def stat(self, what):
        return self.L.dda_get_stat(what)

    def reset_stats(self):
        self.L.dda_reset_stats()

    def apply_dw(self, v, precision="double"):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros_like(v)
        self.L.dda_apply_dw(0 if precision == "double" else 1, _dp(out), _dp(v))
        return out

    def operator_arrays(self):
        D = np.zeros((self.V, 4, 3, 3), dtype=np.complex128)
        cl = np.zeros((self.V, 42), dtype=np.complex128)
        self.L.dda_get_operator(_dp(D), _dp(cl))
        return D, cl

    def level_shape(self, depth):
        return self.info(INFO.SITES, depth), self.info(INFO.SITE_VARS, depth)

    def set_interpolation(self, depth, P):
        P = np.ascontiguousarray(P, dtype=np.complex64)
        self.L.dda_set_interpolation(depth, _fp(P))

    def get_interpolation(self, depth):
        V, nc = self.level_shape(depth)
        P = np.zeros((V * nc, self.info(INFO.TEST_VECTORS, depth)), dtype=np.complex64)
        self.L.dda_get_interpolation(depth, _fp(P))
        return P

    def _level_op(self, op, depth, vin, out_depth, iparam=0, flag=0, out=None):
        vin = np.ascontiguousarray(vin, dtype=np.complex64)
        V, nc = self.level_shape(out_depth)
        if out is None:
            out = np.zeros(V * nc, dtype=np.complex64)
        self.L.dda_level_op(op, depth, _fp(out), _fp(vin), iparam, flag)
        return out

    def level_apply(self, depth, v):
        return self._level_op(OP.APPLY, depth, v, depth)

    def restrict(self, depth, v):
        return self._level_op(OP.RESTRICT, depth, v, depth + 1)

    def interpolate(self, depth, vc):
        return self._level_op(OP.INTERPOLATE, depth, vc, depth)

    def smoother(self, depth, eta, n, phi0=None):
        out = None if phi0 is None else np.ascontiguousarray(phi0, dtype=np.complex64).copy()
        return self._level_op(OP.SMOOTHER, depth, eta, depth, iparam=n, flag=0 if phi0 is None else 1, out=out)

    def vcycle(self, depth, eta):
        return self._level_op(OP.VCYCLE, depth, eta, depth)

    def coarsest_solve(self, b):
        d = self.info(INFO.NUM_LEVELS) - 1
        return self._level_op(OP.COARSEST_SOLVE, d, b, d)

    def write_test_vectors(self, basename):
        self.L.dda_write_test_vectors(basename.encode())

    def read_test_vectors(self, basename):
        self.L.dda_read_test_vectors(basename.encode())

    def setup_if_necessary(self):
        return self.L.dda_setup_if_necessary()

    def level_apply_mrhs(self, depth, vs, reps=0):
        """12 right-hand sides at once on the tensor cores: vs [12][sites * site vars] -> (D_c vs[j])_j, ms per application."""
        vs = np.ascontiguousarray(vs, dtype=np.complex64)
        assert vs.shape[0] == 12
        out = np.zeros_like(vs)
        ms = np.zeros(1)
        rc = self.L.dda_level_apply_mrhs(depth, _fp(out), _fp(vs), int(reps), _dp(ms))
        if rc != 0:
            raise RuntimeError("dda_level_apply_mrhs: level %d not supported" % depth)
        return out, float(ms[0])

    def bench_op(self, op, depth=0, reps=10):
        return self.L.dda_bench_op(op, depth, reps)

    def solve_device(self, b, tol=1e-10):
        b = np.ascontiguousarray(b, dtype=np.complex128)
        self.L.dda_upload_source(_dp(b))
        st = np.zeros(2, dtype=np.int32)
        ms = np.zeros(1)
        res = self.L.dda_solve_device(tol, _ip(st), _dp(ms))
        return res, st, float(ms[0])

    def download_solution(self):
        x = np.zeros(self.V * 12, dtype=np.complex128)
        self.L.dda_download_solution(_dp(x))
        return x

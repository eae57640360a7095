"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/_ref/libddref.so = the UNMODIFIED reference (mrottmann/DDalphaAMG) built by
oracle/build_ref.sh against a single-rank MPI shim.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the product (ddalphaamg_b200)
never does.

The reference keeps process-global state (global_struct g, static level_struct l,
src/dd_alpha_amg.c:28-33), so one Python process can hold ONE reference instance at a time.
"""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path(flavour=""):
    return os.path.join(_HERE, "_ref", "libddref%s.so" % flavour)


def available(flavour=""):
    return os.path.exists(lib_path(flavour))


def load(flavour=""):
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(lib_path(flavour), mode=C.RTLD_LOCAL)
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    L.ref_init.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_int, C.c_int]
    L.ref_set_conf.argtypes = [dp]; L.ref_set_conf.restype = C.c_double
    L.ref_setup.argtypes = [C.c_int, ip]
    L.ref_setup_mt.argtypes = [C.c_int, C.c_int, ip]
    L.ref_solve.argtypes = [dp, dp, C.c_double, ip]; L.ref_solve.restype = C.c_double
    L.ref_solve_scaled.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, ip]; L.ref_solve_scaled.restype = C.c_double
    L.ref_shift_mass.argtypes = [C.c_double]
    L.ref_solve_mt.argtypes = [dp, dp, C.c_double, ip, dp]; L.ref_solve_mt.restype = C.c_double
    L.ref_info.argtypes = [C.c_int, C.c_int]; L.ref_info.restype = C.c_int
    L.ref_get_D.argtypes = [dp]; L.ref_get_clover.argtypes = [dp]
    L.ref_dw_double.argtypes = [dp, dp]; L.ref_dw_float.argtypes = [dp, dp]
    L.ref_dw_double_time.argtypes = [C.c_int, C.c_int]; L.ref_dw_double_time.restype = C.c_double
    L.ref_preconditioner.argtypes = [dp, dp]
    L.ref_get_translation.argtypes = [C.c_int, ip]
    L.ref_get_interpolation.argtypes = [C.c_int, fp]
    L.ref_coarse_apply.argtypes = [C.c_int, fp, fp]
    L.ref_restrict.argtypes = [C.c_int, fp, fp]
    L.ref_interpolate.argtypes = [C.c_int, fp, fp]
    L.ref_smoother.argtypes = [C.c_int, fp, fp, C.c_int, C.c_int]
    L.ref_coarsest_solve.argtypes = [fp, fp]; L.ref_coarsest_solve.restype = C.c_int
    L.ref_vcycle.argtypes = [C.c_int, fp, fp]
    L.ref_plaquette.restype = C.c_double
    L.ref_norm_res.restype = C.c_double
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def write_ini(path, lattice, block, levels=2, test_vectors=(20, 28), setup_iter=(4, 3), post_smooth=(2, 2),
              block_iter=(4, 4), m0=-0.5, csw=1.0, tol=1e-10, restart=50, max_restart=20, coarse_tol=5e-2,
              coarse_iter=100, coarse_restart=5, mixed_precision=1, anti_pbc=1, method=2, kcycle=1,
              coarse_lattice=None, coarse_block=None, nthreads=1, local_lattice=None, odd_even=1, ncycle=(1, 1), relax=(1.0, 1.0),
              interpolation=2, tv_file=None):
    """Writes a .ini in the reference's key:value format (keys: src/init.c:592-962)."""
    loc = local_lattice or lattice
    lines = ["configuration: none", "format: 0", "right hand side: 0",
             "antiperiodic boundary conditions: %d" % anti_pbc, "number of levels: %d" % levels,
             "number of openmp threads: %d" % nthreads,
             "d0 global lattice: %d %d %d %d" % tuple(lattice), "d0 local lattice: %d %d %d %d" % tuple(loc),
             "d0 block lattice: %d %d %d %d" % tuple(block)]
    for d in range(levels - 1):
        lines += ["d%d post smooth iter: %d" % (d, post_smooth[min(d, len(post_smooth) - 1)]),
                  "d%d block iter: %d" % (d, block_iter[min(d, len(block_iter) - 1)]),
                  "d%d test vectors: %d" % (d, test_vectors[min(d, len(test_vectors) - 1)]),
                  "d%d setup iter: %d" % (d, setup_iter[min(d, len(setup_iter) - 1)]),
                  "d%d preconditioner cycles: %d" % (d, ncycle[min(d, len(ncycle) - 1)]),
                  "d%d relaxation factor: %.16g" % (d, relax[min(d, len(relax) - 1)])]
    if levels > 2:
        cl = coarse_lattice or [a // b for a, b in zip(lattice, block)]
        lines += ["d1 global lattice: %d %d %d %d" % tuple(cl), "d1 local lattice: %d %d %d %d" % tuple(cl)]
        if coarse_block is not None:
            lines += ["d1 block lattice: %d %d %d %d" % tuple(coarse_block)]
    lines += ["m0: %.16g" % m0, "csw: %.16g" % csw, "tolerance for relative residual: %g" % tol,
              "iterations between restarts: %d" % restart, "maximum of restarts: %d" % max_restart,
              "coarse grid tolerance: %g" % coarse_tol, "coarse grid iterations: %d" % coarse_iter,
              "coarse grid restarts: %d" % coarse_restart, "print mode: 1", "method: %d" % method,
              "mixed precision: %d" % mixed_precision, "randomize test vectors: 0",
              "odd even preconditioning: %d" % odd_even, "kcycle: %d" % kcycle, "kcycle length: 5",
              "kcycle restarts: 2", "kcycle tolerance: 1E-1", "interpolation: %d" % interpolation]
    if tv_file is not None:
        lines += ["test vector io from single file: 0", "test vector io file name: %s" % tv_file]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def read_conf(path, anti_pbc=True):
    """Reference native conf format (src/io.c:486-506): 4 x int32 [T,Z,Y,X], 1 double plaquette, then
    [t][z][y][x][mu][3][3][re,im] doubles.  anti_pbc flips U_T on the last time slice as io.c:535-541 does."""
    with open(path, "rb") as f:
        dims = np.fromfile(f, dtype=np.int32, count=4)
        plaq = float(np.fromfile(f, dtype=np.float64, count=1)[0])
        n = int(np.prod(dims)) * 72
        data = np.fromfile(f, dtype=np.float64, count=n)
    U = data.reshape(tuple(int(d) for d in dims) + (4, 3, 3, 2)).copy()
    if anti_pbc:
        U[-1, :, :, :, 0] *= -1.0
    return [int(d) for d in dims], plaq, U


class Reference:
    """One live instance of the reference solver (library route, src/dd_alpha_amg.c)."""

    def __init__(self, lattice, block, m0=-0.5, csw=1.0, print_mode=-1, flavour="", **ini_kw):
        self.L = load(flavour)
        self.lattice = list(lattice)
        self.V = int(np.prod(lattice))
        self._tmp = tempfile.NamedTemporaryFile(suffix=".ini", delete=False)
        self._tmp.close()
        bc = ini_kw.pop("bc", None)
        write_ini(self._tmp.name, lattice, block, m0=m0, csw=csw, **ini_kw)
        if bc is None:
            bc = 2 if ini_kw.get("anti_pbc", 1) else 1
        self.L.ref_init(self._tmp.name.encode(), m0, csw, bc, print_mode)

    def set_conf(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        return self.L.ref_set_conf(_dp(U))

    def setup(self, iters, nthreads=1):
        st = np.zeros(2, dtype=np.int32)
        if nthreads > 1:
            self.L.ref_setup_mt(iters, nthreads, _ip(st))
        else:
            self.L.ref_setup(iters, _ip(st))
        return st

    def solve(self, b, tol=1e-10, scale_even=1.0, scale_odd=1.0):
        b = np.ascontiguousarray(b, dtype=np.complex128)
        x = np.zeros_like(b)
        st = np.zeros(2, dtype=np.int32)
        if scale_even != 1.0 or scale_odd != 1.0:
            res = self.L.ref_solve_scaled(_dp(x), _dp(b), tol, scale_even, scale_odd, _ip(st))
        else:
            res = self.L.ref_solve(_dp(x), _dp(b), tol, _ip(st))
        return x, res, st

    def shift_mass(self, m0):
        """shift_update (src/dirac.c:669-691): new mass on every level of the hierarchy."""
        self.L.ref_shift_mass(float(m0))

    def solve_mt(self, b, tol=1e-10):
        b = np.ascontiguousarray(b, dtype=np.complex128)
        x = np.zeros_like(b)
        st = np.zeros(2, dtype=np.int32)
        sec = np.zeros(1)
        res = self.L.ref_solve_mt(_dp(x), _dp(b), tol, _ip(st), _dp(sec))
        return x, res, st, float(sec[0])

    def info(self, what, depth=0):
        return self.L.ref_info(what, depth)

    def D(self):
        out = np.zeros((self.V, 4, 3, 3), dtype=np.complex128)
        self.L.ref_get_D(_dp(out))
        return out

    def clover(self):
        out = np.zeros((self.V, 42), dtype=np.complex128)
        self.L.ref_get_clover(_dp(out))
        return out

    def dw_double(self, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros_like(v)
        self.L.ref_dw_double(_dp(out), _dp(v))
        return out

    def dw_float(self, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros_like(v)
        self.L.ref_dw_float(_dp(out), _dp(v))
        return out

    def dw_time(self, reps=10, nthreads=1):
        return self.L.ref_dw_double_time(reps, nthreads)

    def preconditioner(self, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        out = np.zeros_like(v)
        self.L.ref_preconditioner(_dp(out), _dp(v))
        return out

    def translation(self, depth=0):
        n = self.info(1, depth)
        out = np.zeros(n, dtype=np.int32)
        self.L.ref_get_translation(depth, _ip(out))
        return out

    def interpolation(self, depth=0):
        n = self.info(1, depth) * self.info(2, depth)
        nv = self.info(3, depth)
        out = np.zeros((n, nv), dtype=np.complex64)
        self.L.ref_get_interpolation(depth, _fp(out))
        return out

    def coarse_apply(self, depth, v):
        v = np.ascontiguousarray(v, dtype=np.complex64)
        out = np.zeros_like(v)
        self.L.ref_coarse_apply(depth, _fp(out), _fp(v))
        return out

    def restrict(self, depth, v):
        v = np.ascontiguousarray(v, dtype=np.complex64)
        nc = self.info(1, depth + 1) * self.info(2, depth + 1)
        out = np.zeros(nc, dtype=np.complex64)
        self.L.ref_restrict(depth, _fp(out), _fp(v))
        return out

    def interpolate(self, depth, vc):
        vc = np.ascontiguousarray(vc, dtype=np.complex64)
        nf = self.info(1, depth) * self.info(2, depth)
        out = np.zeros(nf, dtype=np.complex64)
        self.L.ref_interpolate(depth, _fp(out), _fp(vc))
        return out

    def smoother(self, depth, eta, n, phi0=None):
        eta = np.ascontiguousarray(eta, dtype=np.complex64)
        phi = np.zeros_like(eta) if phi0 is None else np.ascontiguousarray(phi0, dtype=np.complex64).copy()
        self.L.ref_smoother(depth, _fp(phi), _fp(eta), n, 0 if phi0 is None else 1)
        return phi

    def coarsest_solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.complex64)
        x = np.zeros_like(b)
        it = self.L.ref_coarsest_solve(_fp(x), _fp(b))
        return x, it

    def vcycle(self, depth, eta):
        eta = np.ascontiguousarray(eta, dtype=np.complex64)
        out = np.zeros_like(eta)
        self.L.ref_vcycle(depth, _fp(out), _fp(eta))
        return out

    def free(self):
        self.L.ref_free()
        try:
            os.unlink(self._tmp.name)
        except OSError:
            pass

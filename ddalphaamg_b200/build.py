"""Builds libdd_alpha_amg.so (sm_100a, nvcc) in-tree.  `python -m ddalphaamg_b200.build [--emu]`.

--emu builds tests/_emu/libdda_emu.so instead: the same host logic compiled by g++ with every kernel body run as a
host loop (DDA_HOST_EMU).  That library exists only so that the control flow (cycles, Krylov, setup) can be unit
tested in a GPU-less container; the package never loads it.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdd_alpha_amg.so")
EMU_DIR = os.path.join(ROOT, "tests", "_emu")
EMU_LIB = os.path.join(EMU_DIR, "libdda_emu.so")
GPU_ONLY = ("dw_kernel.cu", "coarse_kernel.cu", "sap_kernel.cu", "transfer_kernel.cu", "schur_kernel.cu", "mrhs_kernel.cu", "galerkin_kernel.cu", "comm.cu")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile_all(srcs, objdir, cmd_for, jobs=8):
    os.makedirs(objdir, exist_ok=True)
    hdrs = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if _newer(o, [s] + hdrs):
            procs.append((s, subprocess.Popen(cmd_for(s, o))))
            while len([p for _, p in procs if p.poll() is None]) >= jobs:
                for _, p in procs:
                    if p.poll() is None:
                        p.wait()
                        break
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError("compilation failed: %s" % s)
    return objs


def build(verbose=False, nofuse=False):
    """nofuse=True: measurement build libdd_alpha_amg_nofuse.so with the round-1 (unfused) complex multiply-adds, loaded
    only through DDA_LIBRARY for A/B timings of the same kernels."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objdir = os.path.join(HERE, "build", "cuda_nofuse" if nofuse else "cuda")
    lib = os.path.join(HERE, "libdd_alpha_amg_nofuse.so") if nofuse else LIB
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
             "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "--use_fast_math=false", "-I" + os.path.join(ROOT, "include")]
    flags = [f for f in flags if f != "--use_fast_math=false"]
    if nofuse:
        flags += ["-DDDA_NO_FUSE"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    objs = _compile_all(srcs, objdir, lambda s, o: ["nvcc"] + flags + ["-c", s, "-o", o])
    if _newer(lib, objs):
        subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                       "-Xcompiler", "-fopenmp", "-lcudart", "-lnccl"])
    return lib


def build_emu():
    srcs = [s for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if os.path.basename(s) not in GPU_ONLY]
    objdir = os.path.join(EMU_DIR, "obj")
    flags = ["-x", "c++", "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-DDDA_HOST_EMU", "-w", "-I" + os.path.join(ROOT, "include")]
    objs = _compile_all(srcs, objdir, lambda s, o: ["g++"] + flags + ["-c", s, "-o", o])
    if _newer(EMU_LIB, objs):
        subprocess.check_call(["g++", "-shared", "-fopenmp", "-o", EMU_LIB] + objs)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu())
    else:
        print(build(verbose="-v" in sys.argv, nofuse="--nofuse" in sys.argv))

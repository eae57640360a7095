"""The C-ABI library loads and exports every symbol that include/*.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    names = set()
    for h in ("dd_alpha_amg.h", "dd_alpha_amg_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(dd_alpha_amg_\w+|DDalphaAMG_\w+|dda_\w+)\s*\(", src):
            names.add(m.group(1))
    names.discard("dd_alpha_amg_par")
    return sorted(names)


def test_library_builds_and_exports_declared_symbols():
    from ddalphaamg_b200 import build as B
    from ddalphaamg_b200.interface import EXPORTS
    lib = B.build()
    L = ctypes.CDLL(lib, mode=ctypes.RTLD_LOCAL)
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), name
    assert set(EXPORTS) == set(decl)


def test_cuda_library_is_sm100a_and_has_no_cpu_path():
    import subprocess
    from ddalphaamg_b200 import library_path
    out = subprocess.run(["cuobjdump", "-lelf", library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out

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


def test_product_package_never_touches_the_oracle_or_a_cpu_path():
    """Static check: nothing under ddalphaamg_b200/ imports or opens oracle/, and the package refuses to load a
    host-emulation build from anywhere but tests/_emu."""
    import glob
    for f in glob.glob(os.path.join(ROOT, "ddalphaamg_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        assert "oracle" not in src, f
    for f in glob.glob(os.path.join(ROOT, "ddalphaamg_b200", "csrc", "*")):
        assert "oracle" not in open(f).read(), f
    import shutil
    import tempfile
    import pytest
    from ddalphaamg_b200 import build as B
    from ddalphaamg_b200.interface import load_library
    emu = B.build_emu()
    with tempfile.TemporaryDirectory() as d:
        other = os.path.join(d, "libdd_alpha_amg.so")
        shutil.copy(emu, other)
        with pytest.raises(RuntimeError):
            load_library(other)


def test_reference_arm_of_bench_prints_the_contract_keys():
    """bench.py --impl reference runs on the host cores only; check the JSON contract on the small 2-level workload."""
    import json
    import subprocess
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libddref.so")):
        import pytest
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "16^3x32-L2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is False and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]

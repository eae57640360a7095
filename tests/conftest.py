import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def emu_lib():
    """Host-emulation build of the SAME host logic (kernels run as host loops); test-only, never loaded by the product."""
    from ddalphaamg_b200 import build as B
    return B.build_emu()


@pytest.fixture(scope="session")
def cuda_lib():
    from ddalphaamg_b200 import library_path
    p = library_path()
    if not os.path.exists(p):
        from ddalphaamg_b200 import build as B
        B.build()
    if not _have_gpu():
        pytest.skip("no CUDA device")
    return p


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle import ref
    if not ref.available():
        if os.path.isdir("/root/reference/src"):
            import subprocess
            subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
        else:
            pytest.skip("oracle/_ref not built")
    return ref


GOLDEN = os.path.join(ROOT, "tests", "golden")
CONF8 = os.path.join(GOLDEN, "conf_8x8x8x8b6.0000id3n1")
CONF4 = os.path.join(GOLDEN, "conf_4x4x4x4b6.0000id3n1")

"""ddalphaamg_b200 -- B200-native (sm_100a) implementation of DDalphaAMG's solve-phase hot path.

The product is the C-ABI shared library ``libdd_alpha_amg.so`` (hand-written CUDA, built by
``python -m ddalphaamg_b200.build``); it exports the reference's library interface (include/dd_alpha_amg.h) and the
operator-level entry points of include/dd_alpha_amg_b200.h.  This package is the thin ctypes host side used by the
tests and by bench.py; it mirrors the reference interface function by function (reference: include/dd_alpha_amg.h:43-83,
src/dd_alpha_amg.c:95-404).  There is no CPU fallback: loading fails loudly when the CUDA library has not been built.
"""
from .interface import (DDalphaAMG, load_library, library_path, write_ini, read_conf, random_gauge_field,  # noqa: F401
                        INFO, OPT, STAT, OP, BENCH)

"""pgmo-b200: B200-native implementation of Kimera-PGMO's data-parallel mesh path.

The product is `lib/libpgmo_b200.so` (hand-written sm_100a CUDA behind the C ABI declared in
include/pgmo_b200.h) plus the header-only C++ shim under include/kimera_pgmo_b200/ that mirrors
the reference's C++ API.  This Python package only carries the ctypes binding (`capi`) and the
synthetic workload generators (`synth`) used by tests/ and bench.py.
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]

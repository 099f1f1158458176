"""libiop_b200 — B200-native prover hot path for libiop (LDE / FRI fold / Merkle build).

The product is the C-ABI shared library libiop_b200/lib/libb200iop.so (include/b200_iop.h), built
from libiop_b200/csrc/*.cu for sm_100a, plus the C++ adapter in libiop_b200/host/ that keeps the
reference's function names.  This Python package is a thin ctypes binding used by the tests and
bench.py; it never computes anything itself and fails loudly when the CUDA library is missing.
"""
from .capi import B200, Domain, FpParams, LIB_PATH, load_library, B200Error  # noqa: F401

"""
sparse_dot_b200 — B200 (sm_100a) backend behind sparse_dot_mkl's sparse-matmul
surface: ``dot_product_mkl`` and ``gram_matrix_mkl`` over scipy CSR/CSC/BSR and
numpy arrays, computed by hand-written CUDA kernels in libsdb200.so through a
ctypes C-ABI (include/sdb200.h).  No CPU fallback: importing this package
without the built library raises ImportError.
"""
__version__ = "0.1.0"

import sys as _sys

# `python -m sparse_dot_b200.build` must be able to run when the library is missing or stale: it is the one
# entry point that does not need it.  Everything else fails loudly on import (no CPU fallback).
_BUILDING = "sparse_dot_b200.build" in getattr(_sys, "orig_argv", [])

if not _BUILDING:
    from .api import (  # noqa: F401
        dot_product_mkl,
        dot_product_transpose_mkl,
        get_version_string,
        gram_matrix_mkl,
        optimize,
        set_debug_mode,
    )
    from ._lib import device_count, kernel_launches, last_spmm_kernel, last_timing_ms  # noqa: F401
    from .resident import ResidentCSR  # noqa: F401

__all__ = [
    "dot_product_mkl",
    "dot_product_transpose_mkl",
    "gram_matrix_mkl",
    "set_debug_mode",
    "get_version_string",
    "device_count",
    "kernel_launches",
    "last_timing_ms",
    "last_spmm_kernel",
    "ResidentCSR",
    "optimize",
]

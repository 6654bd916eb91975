"""
ref_pkg.py — loader for the UNMODIFIED reference package.  TEST / BASELINE INFRASTRUCTURE ONLY.

``load()`` imports ``sparse_dot_mkl`` 0.9.6 from ``baseline/_ref`` (installed there by
``__graft_entry__.build()`` with ``pip install --no-deps --target``; git-ignored, never part of this
repo's sources).  The package needs a ``libmkl_rt`` (sparse_dot_mkl/_mkl_interface/_load_library.py:31-96);
this image has none, so unless the caller's environment already names one (``$MKL_RT`` or a system
``libmkl_rt``) the loader points ``$MKL_RT`` at ``oracle/_ref/libmkl_fwd.so`` (oracle/mkl_fwd.c): glue that
resolves the sparse BLAS names to the REAL oneMKL 2024.2 embedded in torch's ``libtorch_cpu.so``.
What then runs is the reference's own Python (validation, handle creation, dispatch, export) over Intel's
``mkl_sparse_?_mm`` / ``mkl_sparse_spmm`` / ``?_spmmd`` / ``?_mv``.  Not available through that route
(absent from the embedded MKL): syrk/syrkd (gram), BSR export, dense syrk.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
"""
import ctypes.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
FWD_LIB = os.path.join(ROOT, "oracle", "_ref", "libmkl_fwd.so")

_cached = None


def load():
    """-> (module or None, one-line status)."""
    global _cached
    if _cached is not None:
        return _cached
    if not os.path.isdir(os.path.join(REF_DIR, "sparse_dot_mkl")):
        _cached = (None, "baseline/_ref/sparse_dot_mkl not installed")
        return _cached
    if "MKL_RT" in os.environ:
        how = "MKL_RT=" + os.environ["MKL_RT"]
    elif ctypes.util.find_library("mkl_rt") is not None:
        how = "system libmkl_rt"
    elif os.path.exists(FWD_LIB):
        os.environ["MKL_RT"] = FWD_LIB
        how = "MKL_RT=oracle/_ref/libmkl_fwd.so -> real oneMKL inside libtorch_cpu.so"
    else:
        how = "no libmkl_rt and oracle/_ref/libmkl_fwd.so not built"
    sys.path.insert(0, REF_DIR)
    try:
        import sparse_dot_mkl

        _cached = (sparse_dot_mkl, f"imported sparse_dot_mkl {sparse_dot_mkl.__version__} ({how}); "
                                   f"{sparse_dot_mkl.get_version_string()}")
    except Exception as e:  # ImportError (no libmkl_rt) / AttributeError (a partial MKL without the glue)
        _cached = (None, f"{type(e).__name__}: {str(e).splitlines()[0][:160]} ({how})")
    finally:
        sys.path.remove(REF_DIR)
    return _cached


def max_threads():
    mod, _ = load()
    if mod is None:
        return 0
    from sparse_dot_mkl._mkl_interface import mkl_get_max_threads

    return int(mkl_get_max_threads())


def set_threads(n):
    mod, _ = load()
    if mod is not None:
        from sparse_dot_mkl._mkl_interface import mkl_set_num_threads

        mkl_set_num_threads(int(n))

"""
libsdb200_mkl.so — the oneMKL symbol names the reference binds, over the libsdb200 C-ABI
(csrc/mkl_shim.cpp).

CPU: the shim exports every one of the 79 symbols `class MKL` binds at class-definition time
(sparse_dot_mkl/_mkl_interface/_cfunctions.py:43-168) — a missing one would be an AttributeError when the
reference is imported.
GPU: the UNMODIFIED reference package (pip-installed under baseline/_ref, which is git-ignored and never
part of this repo's sources) is imported with MKL_RT pointing at the shim and its OWN hot-path test files
are run on the B200 backend, including its dense x dense file: cblas_?gemm / cblas_?syrk are served by
sdb_gemm / sdb_syrk_dense (csrc/dense.cu) since round 2, so NO failure is allowed.
"""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "sparse_dot_b200", "libsdb200_mkl.so")
REF = os.path.join(ROOT, "baseline", "_ref")

BOUND_BY_REFERENCE = """
MKL_Get_Max_Threads MKL_Get_Version MKL_Get_Version_String MKL_Set_Interface_Layer MKL_Set_Num_Threads
MKL_Set_Num_Threads_Local cblas_cgemm cblas_csyrk cblas_dgemm cblas_dsyrk cblas_sgemm cblas_ssyrk cblas_zgemm
cblas_zsyrk dcg dcg_check dcg_get dcg_init dcgmrhs dcgmrhs_check dcgmrhs_get dcgmrhs_init dfgmres dfgmres_check
dfgmres_get dfgmres_init mkl_free_buffers mkl_sparse_c_create_bsr mkl_sparse_c_create_csc mkl_sparse_c_create_csr
mkl_sparse_c_export_bsr mkl_sparse_c_export_csc mkl_sparse_c_export_csr mkl_sparse_c_mm mkl_sparse_c_mv
mkl_sparse_c_spmmd mkl_sparse_c_syrkd mkl_sparse_convert_csr mkl_sparse_d_create_bsr mkl_sparse_d_create_csc
mkl_sparse_d_create_csr mkl_sparse_d_export_bsr mkl_sparse_d_export_csc mkl_sparse_d_export_csr mkl_sparse_d_mm
mkl_sparse_d_mv mkl_sparse_d_qr_factorize mkl_sparse_d_qr_solve mkl_sparse_d_spmmd mkl_sparse_d_syrkd
mkl_sparse_destroy mkl_sparse_order mkl_sparse_qr_reorder mkl_sparse_s_create_bsr mkl_sparse_s_create_csc
mkl_sparse_s_create_csr mkl_sparse_s_export_bsr mkl_sparse_s_export_csc mkl_sparse_s_export_csr mkl_sparse_s_mm
mkl_sparse_s_mv mkl_sparse_s_qr_factorize mkl_sparse_s_qr_solve mkl_sparse_s_spmmd mkl_sparse_s_syrkd
mkl_sparse_spmm mkl_sparse_syrk mkl_sparse_z_create_bsr mkl_sparse_z_create_csc mkl_sparse_z_create_csr
mkl_sparse_z_export_bsr mkl_sparse_z_export_csc mkl_sparse_z_export_csr mkl_sparse_z_mm mkl_sparse_z_mv
mkl_sparse_z_spmmd mkl_sparse_z_syrkd pardiso pardisoinit
""".split()



def test_shim_exports_every_symbol_the_reference_binds():
    assert len(BOUND_BY_REFERENCE) == 79
    lib = ctypes.CDLL(SHIM)
    missing = [n for n in BOUND_BY_REFERENCE if not hasattr(lib, n)]
    assert not missing, missing


def test_shim_refuses_64_bit_index_arrays_so_the_reference_probes_down_to_int32():
    """_mkl_interface/__init__.py:107-125 tries int64 first; an LP64 library must fail that attempt."""
    import numpy as np

    lib = ctypes.CDLL(SHIM)
    indptr = np.array([0, 2, 3], dtype=np.int64)
    idx = np.array([0, 1, 1], dtype=np.int64)
    val = np.ones(3, dtype=np.float32)
    h = ctypes.c_void_p()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    st = lib.mkl_sparse_s_create_csr(ctypes.byref(h), 0, ctypes.c_longlong(2), ctypes.c_longlong(2),
                                     p(indptr[:-1]), p(indptr[1:]), p(idx), p(val))
    assert st == 3 and not h.value  # SPARSE_STATUS_INVALID_VALUE before anything touches a GPU
    assert lib.mkl_sparse_destroy(None) == 1  # NULL handle -> NOT_INITIALIZED (tests/test_mkl.py:128-141)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_unmodified_reference_suite_runs_on_the_b200_backend():
    if not os.path.isdir(os.path.join(REF, "sparse_dot_mkl", "tests")):
        pytest.skip("baseline/_ref is not installed (pip install --no-deps --target baseline/_ref <reference>)")
    env = dict(os.environ, MKL_RT=SHIM, PYTHONPATH=REF)
    env.pop("MKL_INTERFACE_LAYER", None)
    files = ["test_mkl.py", "test_sparse_dense.py", "test_sparse_sparse.py", "test_sparse_vector.py",
             "test_gram_matrix.py", "test_dense_dense.py"]
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "--no-header", "-p", "no:cacheprovider", "-rf"]
        + [os.path.join("sparse_dot_mkl", "tests", f) for f in files],
        cwd=REF, env=env, capture_output=True, text=True, timeout=850)
    tail = r.stdout[-4000:]
    summary = re.search(r"(\d+) passed", r.stdout)
    assert summary, tail + r.stderr[-2000:]
    passed = int(summary.group(1))
    failed = [ln for ln in r.stdout.splitlines() if ln.startswith("FAILED")]
    assert not failed, "\n".join(failed) + "\n" + tail
    assert passed >= 940, tail  # 901 hot-path cases + the 44 of test_dense_dense.py on the round-2 box

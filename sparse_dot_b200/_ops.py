"""
Per-operation drivers: validate -> upload handle(s) -> ONE libsdb200 call ->
download.  Each mirrors the behaviour of a reference driver:

    dense_times_dense         _dense_dense.py:14-71     (_dense_matmul)
    dot_dense_dense           _dense_dense.py:74-88     (_dense_dot_dense)
    sparse_times_dense        _sparse_dense.py:34-132   (_sparse_dense_matmul)
    dot_sparse_dense          _sparse_dense.py:135-208  (_sparse_dot_dense)
    sparse_times_vector       _sparse_vector.py:28-102  (_sparse_dense_vector_mult)
    dot_sparse_vector         _sparse_vector.py:105-174 (_sparse_dot_vector)
    dot_sparse_sparse         _sparse_sparse.py:109-244 (_sparse_dot_sparse)
    gram                      _gram_matrix.py:252-335   (_gram_matrix)

There is no CPU arithmetic here: every product is computed by the CUDA library
(which raises through _lib.check when no GPU is usable).
"""
import ctypes as _ct
import time as _time

import numpy as np
import scipy.sparse as sps

from . import _handles as _h
from . import _lib
from . import _validate as _v
from ._lib import SDB, check, scalar_pair


def _ptr(a):
    return a.ctypes.data_as(_ct.c_void_p)


def _timer(msg=None, since=None):
    """debug_timer of the reference (_common.py:138-155): wall-clock phase
    prints, plus the device-side H2D / kernel / D2H split of the last call."""
    if not SDB.DEBUG:
        return None
    now = _time.time()
    if msg is not None and since is not None:
        print(f"{msg}: {now - since:.6f} seconds")
    return now


# ---------------------------------------------------------------- sparse x dense
def sparse_times_dense(a_sparse, b_dense, scalar=1.0, transpose=False, out=None, out_scalar=None, out_t=None):
    """alpha * op(A) @ B + beta * out  with A sparse and B a contiguous array.
    The result has B's memory order; beta defaults to 1 when ``out`` is given."""
    m = a_sparse.shape[1] if transpose else a_sparse.shape[0]
    shape = (m, b_dense.shape[1])
    layout, ldb = _v.dense_layout(b_dense, other=out)
    if layout == _lib.LAYOUT_C and not transpose and _v.is_csr(a_sparse):
        return _csr_times_rowmajor(a_sparse, b_dense, shape, scalar, out, out_scalar, out_t)
    handle, dbl, cplx = _h.create(a_sparse)
    with handle:
        dtype = _v.OUTPUT_DTYPES[(dbl, cplx)]
        order = "C" if layout == _lib.LAYOUT_C else "F"
        fresh = out is None
        result = _v.output_array(shape, dtype, order, out=out, out_t=out_t, zero=False)
        _, ldy = _v.dense_layout(result, other=b_dense)
        # a fresh result is never read: beta = 0 instead of the reference's zeros + 1.0
        beta = 0.0 if fresh else (1.0 if out_scalar is None else out_scalar)
        status = SDB.lib.sdb_spmm(
            _lib.OP_T if transpose else _lib.OP_N, scalar_pair(1.0 if scalar is None else scalar), handle.ref,
            layout, _ptr(b_dense), shape[1], ldb, scalar_pair(beta), _ptr(result), ldy,
        )
        check(status, "sdb_spmm")
        if SDB.DEBUG:
            h2d, krn, d2h = _lib.last_timing_ms()
            print(f"sdb_spmm device time: H2D {h2d:.3f} ms, kernels {krn:.3f} ms, D2H {d2h:.3f} ms")
    return result


def _csr_times_rowmajor(a, b, shape, scalar, out, out_scalar, out_t):
    """CSR @ C-ordered array: one library call does upload, product and download
    (sdb_spmm_csr_host; a three-stream row-chunk pipeline when the host arrays
    are page-locked, the create + spmm + destroy triple otherwise)."""
    dbl, cplx = _v.precision_flags(a)
    dtype = _v.OUTPUT_DTYPES[(dbl, cplx)]
    result = _v.output_array(shape, dtype, "C", out=out, out_t=out_t, zero=False)
    beta = 0.0 if out is None else (1.0 if out_scalar is None else out_scalar)
    indptr, indices, bits = _h._index_arrays(a)
    data = np.ascontiguousarray(a.data)
    if data.shape[0] != indices.shape[0] or indptr.shape[0] != a.shape[0] + 1:
        raise ValueError("Sparse matrix arrays are inconsistent with its shape")
    status = SDB.lib.sdb_spmm_csr_host(
        a.shape[0], a.shape[1], _ptr(indptr), _ptr(indices), bits, _ptr(data), _h._DTYPE_CODE[np.dtype(a.dtype)],
        scalar_pair(1.0 if scalar is None else scalar), _ptr(b), shape[1], shape[1], scalar_pair(beta),
        _ptr(result), shape[1],
    )
    check(status, "sdb_spmm_csr_host")
    if SDB.DEBUG:
        t0, t1, t2 = _lib.last_timing_ms()
        print(f"sdb_spmm_csr_host device time: {t0:.3f} / {t1:.3f} / {t2:.3f} ms (see sdb200.h)")
    return result


def dot_sparse_dense(a, b, cast=False, scalar=1.0, out=None, out_scalar=None):
    """One sparse and one dense operand, either side."""
    if not _v.is_supported_sparse(a) or not _v.is_supported_sparse(b):
        raise ValueError("Only CSR, CSC, and BSR-type sparse matrices are supported; COO is not")
    _v.check_shapes(a, b)
    if _v.product_is_empty(a, b):
        _v.debug_print("Skipping multiplication because A (dot) B must yield empty matrix")
        both_f32 = a.dtype == b.dtype and a.dtype == np.float32
        return _v.output_array((a.shape[0], b.shape[1]), np.float32 if both_f32 else np.float64, out=out)
    a, b = _v.unify_dtypes(a, b, cast=cast)
    n_sparse = int(sps.issparse(a)) + int(sps.issparse(b))
    if n_sparse != 1:
        raise ValueError("_sparse_dot_dense takes one sparse and one dense array")
    if sps.issparse(a):
        return sparse_times_dense(a, b, scalar=scalar, out=out, out_scalar=out_scalar)
    # dense @ sparse = (sparse^T @ dense^T)^T, all as views
    if out is not None:
        sparse_times_dense(b, a.T, scalar=scalar, transpose=True, out=out.T, out_scalar=out_scalar, out_t=True)
        return out
    return sparse_times_dense(b, a.T, scalar=scalar, transpose=True).T


# ---------------------------------------------------------------- dense x dense
_CBLAS_NOTRANS, _CBLAS_TRANS, _CBLAS_UPPER = 111, 112, 121


def dense_times_dense(a, b, scalar=1.0, out=None, out_scalar=None):
    """cblas_?gemm as the reference drives it (_dense_dense.py:14-71): the result takes A's memory order, B is
    passed as transposed when its order differs, a 1-d B is a column and the result is flattened."""
    dbl, cplx = _v.precision_flags(a)
    flatten = b.ndim == 1
    b = b.reshape(-1, 1) if flatten else b
    m, n, k = a.shape[0], b.shape[1], a.shape[1]
    layout_a, lda = _v.dense_layout(a)
    layout_b, ldb = _v.dense_layout(b)
    op_b = _CBLAS_TRANS if layout_b != layout_a else _CBLAS_NOTRANS
    order, ldc = ("C", n) if layout_a == _lib.LAYOUT_C else ("F", m)
    result = _v.output_array((m, n), _v.OUTPUT_DTYPES[(dbl, cplx)], order, out=out, zero=False)
    beta = 0.0 if out is None else (1.0 if out_scalar is None else out_scalar)
    check(
        SDB.lib.sdb_gemm(layout_a, _CBLAS_NOTRANS, op_b, m, n, k, scalar_pair(1.0 if scalar is None else scalar),
                         _ptr(a), lda, _ptr(b), ldb, scalar_pair(beta), _ptr(result), ldc,
                         _h._DTYPE_CODE[np.dtype(a.dtype)]),
        "sdb_gemm",
    )
    return result.ravel() if flatten else result


def dot_dense_dense(a, b, cast=False, scalar=1.0, out=None, out_scalar=None):
    _v.check_shapes(a, b, allow_vector=True)
    if _v.product_is_empty(a, b):
        _v.debug_print("Skipping multiplication because A (dot) B must yield an empty matrix")
        both_f32 = a.dtype == b.dtype and a.dtype == np.float32
        return _v.output_array((a.shape[0], b.shape[1]), np.float32 if both_f32 else np.float64, out=out)
    a, b = _v.unify_dtypes(a, b, cast=cast)
    return dense_times_dense(a, b, scalar=scalar, out=out, out_scalar=out_scalar)


# ---------------------------------------------------------------- sparse x vector
def sparse_times_vector(a_sparse, vec, scalar=1.0, transpose=False, out=None, out_scalar=None, out_t=None):
    """SpMV through the SpMM kernel with one column."""
    m = a_sparse.shape[1] if transpose else a_sparse.shape[0]
    shape = (m,) if vec.ndim == 1 else (m, 1)
    if _v.product_is_empty(a_sparse, vec):
        both_f32 = a_sparse.dtype == vec.dtype and a_sparse.dtype == np.float32
        return _v.output_array(shape, np.float32 if both_f32 else np.float64, out=out)
    handle, dbl, cplx = _h.create(a_sparse)
    with handle:
        x = np.ascontiguousarray(vec.ravel())
        fresh = out is None
        result = _v.output_array(shape, _v.OUTPUT_DTYPES[(dbl, cplx)], out=out, out_t=out_t, zero=False)
        beta = 0.0 if fresh else (1.0 if out_scalar is None else out_scalar)
        status = SDB.lib.sdb_spmm(
            _lib.OP_T if transpose else _lib.OP_N, scalar_pair(1.0 if scalar is None else scalar), handle.ref,
            _lib.LAYOUT_C, _ptr(x), 1, 1, scalar_pair(beta), _ptr(result), 1,
        )
        check(status, "sdb_spmm")
    return result


def dot_sparse_vector(a, b, cast=False, scalar=1.0, out=None, out_scalar=None):
    if not _v.is_supported_sparse(a) or not _v.is_supported_sparse(b):
        raise ValueError("Only CSR, CSC, and BSR-type sparse matrices are supported")
    _v.check_shapes(a, b, allow_vector=True)
    a, b = _v.unify_dtypes(a, b, cast=cast)
    if _v.is_dense_vector(b):
        return sparse_times_vector(a, b, scalar=scalar, out=out, out_scalar=out_scalar)
    if _v.is_dense_vector(a):
        if out is None:
            return sparse_times_vector(b, a.T, scalar=scalar, transpose=True).T
        sparse_times_vector(b, a.T, scalar=scalar, transpose=True, out=out.T, out_scalar=out_scalar, out_t=True)
        return out
    raise ValueError("Neither mv_a or mv_b is a dense vector")


# ---------------------------------------------------------------- sparse x sparse
def _spgemm_handle(ha, hb, ordered=False):
    """C = A @ B as a new device handle (reference: _matmul_mkl); ``ordered`` fuses the
    reference's separate mkl_sparse_order pass into the product."""
    ref = _ct.c_void_p()
    fn = "sdb_spgemm_ordered" if ordered else "sdb_spgemm"
    check(getattr(SDB.lib, fn)(_lib.OP_N, ha.ref, hb.ref, _ct.byref(ref)), fn)
    return _h.Handle(ref, ha.dtype)


def _spgemm_dense(ha, hb, shape, dtype, out=None):
    """Dense row-major C = A @ B, overwriting (reference: _matmul_mkl_dense)."""
    result = _v.output_array(shape, dtype, out=out, zero=False)
    check(SDB.lib.sdb_spgemm_dense(_lib.OP_N, ha.ref, hb.ref, _lib.LAYOUT_C, _ptr(result), shape[1]),
          "sdb_spgemm_dense")
    return result


def dot_sparse_sparse(a, b, cast=False, reorder_output=False, dense=False, out=None):
    if not _v.is_supported_sparse(a) or not _v.is_supported_sparse(b):
        raise ValueError("Input matrices to dot_product_mkl must be CSR, CSC, or BSR; COO is not supported")
    if out is not None and not dense:
        raise ValueError(
            "out argument cannot be used with sparse (dot) sparse matrix multiplication unless dense=True"
        )
    ctor, out_type = _v.sparse_container(a)
    _v.check_shapes(a, b)
    shape = (a.shape[0], b.shape[1])
    if _v.product_is_empty(a, b):
        if dense:
            return _v.output_array(shape, a.dtype, out=out)
        return ctor(shape, dtype=a.dtype)
    a, b = _v.unify_dtypes(a, b, cast=cast)

    t = _timer()
    ha, a_dbl, a_cplx = _h.create(a)
    try:
        hb, b_dbl, _ = _h.create(b)
    except Exception:
        ha.destroy()
        raise
    t = _timer("Created device sparse handles", t)
    with ha, hb:
        if dense:
            result = _spgemm_dense(ha, hb, shape, _v.OUTPUT_DTYPES[(a_dbl or b_dbl, a_cplx)], out=out)
            _timer("Multiplied matrices", t)
            return result
        hc = _spgemm_handle(ha, hb, ordered=reorder_output)
    with hc:
        t = _timer("Multiplied matrices" + (" (ordered)" if reorder_output else ""), t)
        result = _h.export(hc, output_type=ctor.__name__)
    _timer("Created python handle", t)
    return result


# ---------------------------------------------------------------- gram matrix
def _gram_sparse(a, aat=False, reorder_output=False):
    """Upper triangle of A^T A (or A A^T) as a csr_matrix — always csr_matrix,
    whatever the input container (_gram_matrix.py:82-87)."""
    handle, _, _ = _h.create(a)
    with handle:
        ref = _ct.c_void_p()
        fn = "sdb_syrk_ordered" if reorder_output else "sdb_syrk"
        check(getattr(SDB.lib, fn)(_lib.OP_N if aat else _lib.OP_T, handle.ref, _ct.byref(ref)), fn)
        with _h.Handle(ref, a.dtype) as hc:
            return _h.export(hc, output_type="csr_matrix")


def _gram_sparse_to_dense(a, aat=False, scalar=1.0, out=None, out_scalar=None):
    """Dense row-major upper triangle: alpha * gram + beta * out.  With no
    ``out`` the strict lower triangle is zero (the reference allocates zeros and
    scrubs what syrkd may have written there, _gram_matrix.py:136-139,168-169);
    with ``out`` it is left exactly as the caller had it."""
    handle, dbl, cplx = _h.create(a)
    with handle:
        n = a.shape[0 if aat else 1]
        dtype = _v.OUTPUT_DTYPES[(dbl, cplx)]
        result = _v.output_array((n, n), dtype, order="C", out=out, zero=False)
        op = _lib.OP_N if aat else _lib.OP_T
        alpha = scalar_pair(1.0 if scalar is None else scalar)
        if out is None:
            check(SDB.lib.sdb_syrkd_new(op, handle.ref, alpha, _ptr(result), _lib.LAYOUT_C, n), "sdb_syrkd_new")
        else:
            beta = scalar_pair(1.0 if out_scalar is None else out_scalar)
            check(SDB.lib.sdb_syrkd(op, handle.ref, alpha, beta, _ptr(result), _lib.LAYOUT_C, n), "sdb_syrkd")
    return result


def _gram_dense_to_dense(a, aat=False, scalar=1.0, out=None, out_scalar=None):
    """cblas_?syrk on a dense array (_gram_matrix.py:196-249): upper triangle of A^T A (or A A^T) in A's memory
    order; a fresh result has zeros below the diagonal, a given ``out`` keeps what it had there."""
    n, k = a.shape if aat else a.shape[::-1]
    layout, lda = _v.dense_layout(a)
    dbl, cplx = _v.precision_flags(a)
    result = _v.output_array((n, n), _v.OUTPUT_DTYPES[(dbl, cplx)], "C" if layout == _lib.LAYOUT_C else "F", out=out)
    beta = 0.0 if out is None else (1.0 if out_scalar is None else out_scalar)
    check(
        SDB.lib.sdb_syrk_dense(layout, _CBLAS_UPPER, _CBLAS_NOTRANS if aat else _CBLAS_TRANS, n, k,
                               scalar_pair(1.0 if scalar is None else scalar), _ptr(a), lda, scalar_pair(beta),
                               _ptr(result), n, _h._DTYPE_CODE[np.dtype(a.dtype)]),
        "sdb_syrk_dense",
    )
    return result


def gram(matrix, transpose=False, cast=False, dense=False, reorder_output=False, out=None, out_scalar=None):
    if _v.product_is_empty(matrix, matrix):
        _v.debug_print("Skipping multiplication because AT (dot) A must yield an empty matrix")
        # shape rule exactly as the reference has it (_gram_matrix.py:288-294)
        n = matrix.shape[1] if transpose else matrix.shape[0]
        maker = sps.csr_matrix if sps.isspmatrix(matrix) else np.zeros
        return maker((n, n), dtype=matrix.dtype)
    if np.iscomplexobj(matrix):
        raise ValueError("gram_matrix_mkl does not support complex datatypes")
    matrix = _v.unify_dtypes(matrix, cast=cast)
    if sps.issparse(matrix) and not (_v.is_csr(matrix) or _v.is_csc(matrix)):
        raise ValueError("gram_matrix requires sparse matrix to be CSR or CSC format")
    if _v.is_csc(matrix) and not cast:
        raise ValueError("gram_matrix cannot use a CSC matrix unless cast=True")
    if not sps.issparse(matrix):
        return _gram_dense_to_dense(matrix, aat=transpose, out=out, out_scalar=out_scalar)
    if dense:
        return _gram_sparse_to_dense(matrix, aat=transpose, out=out, out_scalar=out_scalar)
    if out is not None:
        raise ValueError("out argument cannot be used with sparse (dot) sparse matrix multiplication")
    return _gram_sparse(matrix, aat=transpose, reorder_output=reorder_output)

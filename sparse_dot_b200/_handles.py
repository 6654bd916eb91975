"""
Device handles for scipy sparse matrices: upload (create), download (export),
order, convert, destroy.  The counterpart of the reference's handle helpers in
sparse_dot_mkl/_mkl_interface/_common.py (:245-384 create, :387-642 export,
:671-722 destroy / order / convert).

Differences that follow from living on a GPU:
  * create COPIES the host arrays into HBM (MKL borrows them), so the caller's
    matrix is never modified — the reference may re-type or re-order it in place;
  * index arrays may be int32 or int64 on the host; in HBM row offsets are
    int64 and column indices int32;
  * export writes straight into numpy-owned arrays (no second copy).
"""
import ctypes as _ct

import numpy as np

from . import _lib
from ._lib import SDB, check
from . import _validate as _v

_DTYPE_CODE = {
    np.dtype(np.float32): _lib.F32,
    np.dtype(np.float64): _lib.F64,
    np.dtype(np.complex64): _lib.C64,
    np.dtype(np.complex128): _lib.C128,
}
_CODE_DTYPE = {v: k for k, v in _DTYPE_CODE.items()}
_INT32_MAX = np.iinfo(np.int32).max


def _ptr(a):
    return a.ctypes.data_as(_ct.c_void_p)


class Handle:
    """Owns one ``sdb_mat*``.  Use as a context manager or call ``destroy()``."""

    __slots__ = ("ref", "dtype", "_keep")

    def __init__(self, ref, dtype):
        self.ref = ref
        self.dtype = np.dtype(dtype)
        self._keep = None

    def __bool__(self):
        return bool(self.ref)

    def destroy(self):
        ref, self.ref = self.ref, _ct.c_void_p()
        check(SDB.lib.sdb_destroy(ref), "sdb_destroy")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.ref:
            self.destroy()
        return False


def _index_arrays(matrix):
    """Contiguous indptr / indices of one common width (32 or 64 bits)."""
    indptr, indices = matrix.indptr, matrix.indices
    if max(matrix.shape) > _INT32_MAX:
        raise ValueError(
            f"libsdb200 keeps int32 column indices and cannot hold matrix {matrix!r}"
        )
    wide = indptr.dtype.itemsize > 4 or indices.dtype.itemsize > 4
    it = np.int64 if wide else np.int32
    indptr = np.ascontiguousarray(indptr, dtype=it)
    indices = np.ascontiguousarray(indices, dtype=it)
    return indptr, indices, 64 if wide else 32


def create(matrix):
    """scipy CSR / CSC / BSR -> Handle (device copy).  Returns
    (handle, double_precision, complex) like the reference's _create_mkl_sparse."""
    dbl, cplx = _v.precision_flags(matrix)
    code = _DTYPE_CODE[np.dtype(matrix.dtype)]
    ref = _ct.c_void_p()
    if _v.is_csr(matrix) or _v.is_csc(matrix):
        fn_name = "sdb_create_csr" if _v.is_csr(matrix) else "sdb_create_csc"
        major = matrix.shape[0] if _v.is_csr(matrix) else matrix.shape[1]
        indptr, indices, bits = _index_arrays(matrix)
        data = np.ascontiguousarray(matrix.data)
        if data.shape[0] != indices.shape[0] or indptr.shape[0] != major + 1:
            raise ValueError("Sparse matrix arrays are inconsistent with its shape")
        status = getattr(SDB.lib, fn_name)(
            _ct.byref(ref), matrix.shape[0], matrix.shape[1], _ptr(indptr), _ptr(indices), bits, _ptr(data), code
        )
        check(status, fn_name)
    elif _v.is_bsr(matrix):
        b = matrix.blocksize[0]
        if b != matrix.blocksize[1]:
            raise ValueError(f"BSR representation requires square blocks; {matrix.blocksize} blocks provided")
        if matrix.shape[0] % b or matrix.shape[1] % b:
            raise ValueError(f"BSR blocks {matrix.blocksize} do not align with dims {matrix.shape}")
        indptr, indices, bits = _index_arrays(matrix)
        data = matrix.data
        if data.ndim == 3 and data.flags.c_contiguous:
            block_layout = _lib.LAYOUT_C
        elif data.ndim == 3 and data.transpose(0, 2, 1).flags.c_contiguous:
            # every block stored column-major
            block_layout, data = _lib.LAYOUT_F, data.transpose(0, 2, 1)
        else:
            block_layout, data = _lib.LAYOUT_C, np.ascontiguousarray(data)
        status = SDB.lib.sdb_create_bsr(
            _ct.byref(ref), matrix.shape[0] // b, matrix.shape[1] // b, b, block_layout,
            _ptr(indptr), _ptr(indices), bits, _ptr(data), code,
        )
        check(status, "sdb_create_bsr")
    else:
        raise ValueError("Matrix is not CSC, CSR, or BSR")
    return Handle(ref, matrix.dtype), dbl, cplx


def info(handle):
    fmt, dt, bl = _ct.c_int(), _ct.c_int(), _ct.c_int()
    rows, cols, nnz, bs = _ct.c_int64(), _ct.c_int64(), _ct.c_int64(), _ct.c_int64()
    check(
        SDB.lib.sdb_get_info(handle.ref, _ct.byref(fmt), _ct.byref(dt), _ct.byref(rows), _ct.byref(cols),
                             _ct.byref(nnz), _ct.byref(bs), _ct.byref(bl)),
        "sdb_get_info",
    )
    return {
        "format": fmt.value, "dtype": _CODE_DTYPE[dt.value], "rows": rows.value, "cols": cols.value,
        "nnz": nnz.value, "block": bs.value, "block_layout": bl.value,
    }


def export(handle, output_type="csr_matrix"):
    """Handle -> scipy matrix of class ``output_type`` (``csr_matrix``,
    ``csc_array``, ``bsr_matrix`` ...).  Index arrays come back int32 when they
    fit (scipy's own rule), else int64."""
    import scipy.sparse as sps

    output_type = output_type.lower()
    ctor = getattr(sps, output_type, None)
    if ctor is None or output_type[:3] not in ("csr", "csc", "bsr"):
        raise ValueError("Only CSR, CSC, and BSR output types are supported")
    meta = info(handle)
    tag = {_lib.FMT_CSR: "csr", _lib.FMT_CSC: "csc", _lib.FMT_BSR: "bsr"}[meta["format"]]
    if tag != output_type[:3]:
        raise ValueError(f"Handle holds a {tag.upper()} matrix, not {output_type[:3].upper()}")
    b = meta["block"]
    shape = (meta["rows"] * b, meta["cols"] * b)
    dtype = meta["dtype"]
    kwargs = {"blocksize": (b, b)} if tag == "bsr" else {}
    if shape[0] == 0 or shape[1] == 0 or meta["nnz"] == 0:
        return ctor(shape, dtype=dtype, **kwargs)
    major = meta["cols"] if tag == "csc" else meta["rows"]
    nnz = meta["nnz"]
    wide = nnz > _INT32_MAX
    it, bits = (np.int64, 64) if wide else (np.int32, 32)
    indptr = np.empty(major + 1, dtype=it)
    indices = np.empty(nnz, dtype=it)
    data = np.empty((nnz, b, b) if tag == "bsr" else (nnz,), dtype=dtype)
    check(SDB.lib.sdb_export(handle.ref, _ptr(indptr), bits, _ptr(indices), bits, _ptr(data)), "sdb_export")
    if tag == "bsr" and meta["block_layout"] == _lib.LAYOUT_F:
        data = data.transpose(0, 2, 1)
    return ctor((data, indices, indptr), shape=shape, **kwargs)


def order(handle):
    check(SDB.lib.sdb_order(handle.ref), "sdb_order")


def convert_to_csr(handle, destroy_original=False):
    ref = _ct.c_void_p()
    check(SDB.lib.sdb_convert_csr(handle.ref, _lib.OP_N, _ct.byref(ref)), "sdb_convert_csr")
    if destroy_original:
        handle.destroy()
    return Handle(ref, handle.dtype)

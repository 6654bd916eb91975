"""
mkl_ref.py — the REAL Intel oneMKL tier of the checker.  TEST INFRASTRUCTURE ONLY.

The reference package cannot be imported in this image: it needs ``libmkl_rt``
(sparse_dot_mkl/_mkl_interface/_load_library.py:31-96) and none is installed.
But torch's ``libtorch_cpu.so`` statically embeds oneMKL 2024.2 (LP64) and
exports the sparse inspector-executor routines the hot path uses:
``mkl_sparse_{s,d,c,z}_{create_csr,create_bsr,export_csr,mm,mv,spmmd}``,
``mkl_sparse_spmm`` and ``mkl_sparse_destroy`` (not ``order``, ``convert_csr``,
``syrk``, ``syrkd`` or any ``*_csc``).  This module repeats, call for call, what
the reference does with them:

    create   _mkl_interface/_common.py:296-324   (4-array CSR, zero based, zero copy)
    mm       _sparse_dense.py:111-123            (descr = GENERAL by value)
    spmm     _sparse_sparse.py:21-44             (op = 10)
    spmmd    _sparse_sparse.py:56-106            (layout 101, ldc = n)
    export   _mkl_interface/_common.py:429-500   (insert rows_start[0]; copy out)
    destroy  _mkl_interface/_common.py:671-680

so its outputs are the reference's numerical results for SpMM / SpGEMM, and its
timings are "the reference's MKL path on this host's cores".  It is used to
generate tests/golden (oracle/gen_golden.py), to validate sdb_oracle.c, and as
the CPU baseline of bench.py.  Gram (syrk/syrkd) is not available here; the
stand-in is ``spmm(op=TRANSPOSE, A, A)`` + upper triangle, labelled as such.
"""
import ctypes
import importlib.util
import os

import numpy as np
import scipy.sparse as sp

_MKL = None
_INT = ctypes.c_int  # LP64
_NPINT = np.int32


class _Descr(ctypes.Structure):
    # struct matrix_descr {type, mode, diag}; reference default (20, 0, 0):
    # _mkl_interface/_structs.py:13-30
    _fields_ = [("type", ctypes.c_int), ("mode", ctypes.c_int), ("diag", ctypes.c_int)]


class _C8(ctypes.Structure):
    _fields_ = [("re", ctypes.c_float), ("im", ctypes.c_float)]


class _C16(ctypes.Structure):
    _fields_ = [("re", ctypes.c_double), ("im", ctypes.c_double)]


_LETTER = {
    np.dtype(np.float32): "s",
    np.dtype(np.float64): "d",
    np.dtype(np.complex64): "c",
    np.dtype(np.complex128): "z",
}


def _scalar(v, dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return ctypes.c_float(v)
    if dtype == np.float64:
        return ctypes.c_double(v)
    v = complex(v)
    return (_C8 if dtype == np.complex64 else _C16)(v.real, v.imag)


def _scalar_type(dtype):
    dtype = np.dtype(dtype)
    return {
        np.dtype(np.float32): ctypes.c_float,
        np.dtype(np.float64): ctypes.c_double,
        np.dtype(np.complex64): _C8,
        np.dtype(np.complex128): _C16,
    }[dtype]


def library_path():
    spec = importlib.util.find_spec("torch")
    if spec is None or not spec.submodule_search_locations:
        return None
    p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libtorch_cpu.so")
    return p if os.path.exists(p) else None


def available():
    try:
        return mkl() is not None
    except OSError:
        return False


def mkl():
    """dlopen libtorch_cpu.so once and declare argtypes for what we call."""
    global _MKL
    if _MKL is not None:
        return _MKL
    path = library_path()
    if path is None:
        return None
    L = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    if not hasattr(L, "mkl_sparse_s_mm"):
        return None
    vp, ip = ctypes.c_void_p, ctypes.POINTER(_INT)
    for dt, ch in _LETTER.items():
        getattr(L, f"mkl_sparse_{ch}_create_csr").argtypes = [
            ctypes.POINTER(vp), ctypes.c_int, _INT, _INT, vp, vp, vp, vp]
        getattr(L, f"mkl_sparse_{ch}_export_csr").argtypes = [
            vp, ctypes.POINTER(ctypes.c_int), ip, ip,
            ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
        st = _scalar_type(dt)
        getattr(L, f"mkl_sparse_{ch}_mm").argtypes = [
            ctypes.c_int, st, vp, _Descr, ctypes.c_int, vp, _INT, _INT, st, vp, _INT]
        getattr(L, f"mkl_sparse_{ch}_spmmd").argtypes = [
            ctypes.c_int, vp, vp, ctypes.c_int, vp, _INT]
    L.mkl_sparse_spmm.argtypes = [ctypes.c_int, vp, vp, ctypes.POINTER(vp)]
    L.mkl_sparse_destroy.argtypes = [vp]
    L.MKL_Get_Max_Threads.restype = ctypes.c_int
    L.MKL_Set_Num_Threads_Local.argtypes = [ctypes.c_int]
    L.MKL_Set_Num_Threads_Local.restype = ctypes.c_int
    _MKL = L
    return L


def version_string():
    buf = ctypes.create_string_buffer(256)
    mkl().MKL_Get_Version_String(buf, 256)
    return buf.value.decode().strip()


def max_threads():
    return int(mkl().MKL_Get_Max_Threads())


def set_threads(n):
    """Thread-local MKL thread count (0 restores the global default)."""
    return int(mkl().MKL_Set_Num_Threads_Local(int(n)))


def _ok(status, name):
    if status != 0:
        raise ValueError(f"{name} returned {status}")


class Handle:
    """One mkl_sparse_?_create_csr handle over a scipy CSR matrix (zero copy:
    the arrays are kept alive by this object, as the reference keeps the scipy
    matrix alive)."""

    def __init__(self, m):
        if not sp.issparse(m) or m.format != "csr":
            raise ValueError("CSR only")
        self.dtype = np.dtype(m.dtype)
        self.shape = m.shape
        self.indptr = np.ascontiguousarray(m.indptr, dtype=_NPINT)
        self.indices = np.ascontiguousarray(m.indices, dtype=_NPINT)
        self.data = np.ascontiguousarray(m.data)
        self.ref = ctypes.c_void_p()
        ch = _LETTER[self.dtype]
        _ok(
            getattr(mkl(), f"mkl_sparse_{ch}_create_csr")(
                ctypes.byref(self.ref), 0, m.shape[0], m.shape[1],
                self.indptr[:-1].ctypes.data, self.indptr[1:].ctypes.data,
                self.indices.ctypes.data, self.data.ctypes.data),
            f"mkl_sparse_{ch}_create_csr",
        )

    def destroy(self):
        if self.ref:
            _ok(mkl().mkl_sparse_destroy(self.ref), "mkl_sparse_destroy")
            self.ref = ctypes.c_void_p()


def spmm(a_csr, x, alpha=1.0, beta=None, y=None, op=10, handle=None):
    """Y = alpha*op(A)@X + beta*Y through mkl_sparse_?_mm, the way
    _sparse_dense_matmul drives it (beta None -> 1.0, _common.py:875)."""
    h = handle or Handle(a_csr)
    try:
        dt = h.dtype
        if x.dtype != dt:
            raise ValueError("dtype mismatch")
        if x.flags.c_contiguous:
            layout, ldx, order = 101, x.shape[1], "C"
        elif x.flags.f_contiguous:
            layout, ldx, order = 102, x.shape[0], "F"
        else:
            raise ValueError("Array is not contiguous")
        m_out = h.shape[1] if op != 10 else h.shape[0]
        if y is None:
            y = np.zeros((m_out, x.shape[1]), dtype=dt, order=order)
        ldy = y.shape[1] if layout == 101 else y.shape[0]
        ch = _LETTER[dt]
        _ok(
            getattr(mkl(), f"mkl_sparse_{ch}_mm")(
                op, _scalar(alpha, dt), h.ref, _Descr(20, 0, 0), layout,
                x.ctypes.data, x.shape[1], ldx,
                _scalar(1.0 if beta is None else beta, dt), y.ctypes.data, ldy),
            f"mkl_sparse_{ch}_mm",
        )
        return y
    finally:
        if handle is None:
            h.destroy()


def spgemm(a_csr, b_csr, op=10):
    """C = op(A) @ B through mkl_sparse_spmm + export_csr; returns a scipy CSR
    with MKL's raw (unsorted) column order and structural zeros kept."""
    ha, hb = Handle(a_csr), Handle(b_csr)
    c = ctypes.c_void_p()
    try:
        _ok(mkl().mkl_sparse_spmm(op, ha.ref, hb.ref, ctypes.byref(c)), "mkl_sparse_spmm")
        ch = _LETTER[ha.dtype]
        base = ctypes.c_int()
        rows, cols = _INT(), _INT()
        rs, re_, ci, va = (ctypes.c_void_p() for _ in range(4))
        _ok(
            getattr(mkl(), f"mkl_sparse_{ch}_export_csr")(
                c, ctypes.byref(base), ctypes.byref(rows), ctypes.byref(cols),
                ctypes.byref(rs), ctypes.byref(re_), ctypes.byref(ci), ctypes.byref(va)),
            f"mkl_sparse_{ch}_export_csr",
        )
        m, n = rows.value, cols.value
        out = sp.csr_matrix((m, n), dtype=ha.dtype)
        if m == 0 or n == 0 or not rs.value:
            return out
        start = np.ctypeslib.as_array(ctypes.cast(rs, ctypes.POINTER(_INT)), shape=(m,))
        end = np.ctypeslib.as_array(ctypes.cast(re_, ctypes.POINTER(_INT)), shape=(m,))
        indptr = np.insert(end, 0, start[0])
        nnz = int(indptr[-1] - indptr[0])
        if nnz == 0:
            return out
        width = 2 if ha.dtype.kind == "c" else 1
        real_t = ctypes.c_float if ha.dtype in (np.float32, np.complex64) else ctypes.c_double
        data = np.array(
            np.ctypeslib.as_array(ctypes.cast(va, ctypes.POINTER(real_t)), shape=(nnz * width,)),
            copy=True,
        )
        if width == 2:
            data = data.view(ha.dtype)
        indices = np.array(
            np.ctypeslib.as_array(ctypes.cast(ci, ctypes.POINTER(_INT)), shape=(nnz,)), copy=True
        )
        out.indptr, out.indices, out.data = indptr.astype(np.int32), indices.astype(np.int32), data
        return out
    finally:
        if c:
            mkl().mkl_sparse_destroy(c)
        ha.destroy()
        hb.destroy()


def spmmd(a_csr, b_csr, out=None):
    """Dense row-major C = A @ B through mkl_sparse_?_spmmd (overwrites)."""
    ha, hb = Handle(a_csr), Handle(b_csr)
    try:
        m, n = a_csr.shape[0], b_csr.shape[1]
        if out is None:
            out = np.zeros((m, n), dtype=ha.dtype)
        ch = _LETTER[ha.dtype]
        _ok(
            getattr(mkl(), f"mkl_sparse_{ch}_spmmd")(10, ha.ref, hb.ref, 101, out.ctypes.data, n),
            f"mkl_sparse_{ch}_spmmd",
        )
        return out
    finally:
        ha.destroy()
        hb.destroy()


def gram_standin(a_csr, aat=False):
    """Stand-in for mkl_sparse_syrk (absent from the embedded MKL):
    spmm(op=TRANSPOSE, A, A) = A^T A, then the upper triangle."""
    if aat:
        at = sp.csr_matrix(a_csr.T)
        full = spgemm(at, at, op=11)
    else:
        full = spgemm(a_csr, a_csr, op=11)
    return sp.triu(full, format="csr")

"""
oracle.py — Python face of the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; nothing under sparse_dot_b200/ does.

Two tiers live here (the third, real oneMKL, is oracle/mkl_ref.py):

* ``c_*``   — ctypes over oracle/libsdb_oracle.so, the plain-C restatement in
              sdb_oracle.c (OpenMP; fast enough to be a CPU baseline "port").
* ``np_*``  — scipy/numpy one-liners: the comparator every reference test uses
              (sparse_dot_mkl/tests/test_mkl.py:53-67 compares against
              ``np.dot`` / ``A.dot(B)`` on the same inputs).  Complex dtypes go
              through this tier only.

Both tiers take and return scipy/numpy objects with the layout conventions of
the reference call sites cited in sdb_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdb_oracle.so")

LAYOUT_ROW, LAYOUT_COL = 101, 102
OP_N, OP_T = 10, 11


def build(force=False):
    """Compile sdb_oracle.c -> libsdb_oracle.so (gcc, seconds)."""
    src = os.path.join(_HERE, "sdb_oracle.c")
    if (
        force
        or not os.path.exists(_SO)
        or os.path.getmtime(_SO) < os.path.getmtime(src)
    ):
        subprocess.run(["make", "-s", "-C", _HERE, "libsdb_oracle.so"], check=True)
    return _SO


def build_ref(force=False):
    """Compile mkl_fwd.c -> oracle/_ref/libmkl_fwd.so (glue for running the unmodified reference on the real
    oneMKL inside libtorch_cpu.so; see oracle/ref_pkg.py).  Returns the path, or None if it cannot be built."""
    so = os.path.join(_HERE, "_ref", "libmkl_fwd.so")
    src = os.path.join(_HERE, "mkl_fwd.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        r = subprocess.run(["make", "-s", "-C", _HERE, "_ref/libmkl_fwd.so"], capture_output=True, text=True)
        if r.returncode != 0:
            return None
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"C oracle handles float32/float64 only, not {dtype}")


def _csr_parts(m):
    """int64 indptr, int32 indices, contiguous values of a CSR matrix."""
    m = m.tocsr() if not sp.issparse(m) or m.format != "csr" else m
    return (
        np.ascontiguousarray(m.indptr, dtype=np.int64),
        np.ascontiguousarray(m.indices, dtype=np.int32),
        np.ascontiguousarray(m.data),
    )


def _mk_csr(shape, indptr, indices, data):
    """scipy CSR over raw arrays without the constructor's checks/pruning; the
    two index arrays share one dtype (int32 when everything fits — scipy's own
    rule — else int64), which scipy's C++ kernels insist on."""
    big = max(int(indptr[-1]) if len(indptr) else 0, max(shape)) > np.iinfo(np.int32).max
    it = np.int64 if big else np.int32
    out = sp.csr_matrix(shape, dtype=data.dtype)
    out.indptr = indptr.astype(it, copy=False)
    out.indices = indices.astype(it, copy=False)
    out.data = data
    return out


def _check(status, name):
    if status != 0:
        raise RuntimeError(f"oracle {name} returned {status}")


def max_threads():
    return lib().orc_max_threads()


def set_threads(n):
    lib().orc_set_threads(ctypes.c_int(int(n)))


# --------------------------------------------------------------------- SpMM
def c_spmm(a_csr, x, alpha=1.0, beta=0.0, y=None, op=OP_N):
    """Y = alpha * op(A) @ X + beta * Y with A CSR; layout follows X's order
    (_sparse_dense.py:93); returns Y (allocated like the reference's
    _out_matrix, _common.py:885-955, when ``y`` is None)."""
    indptr, indices, data = _csr_parts(a_csr)
    suf = _suffix(data.dtype)
    if x.dtype != data.dtype:
        raise TypeError("dtype mismatch")
    m_out = a_csr.shape[0] if op == OP_N else a_csr.shape[1]
    n = x.shape[1]
    if x.flags.c_contiguous and not (x.flags.f_contiguous and y is not None and not y.flags.c_contiguous):
        layout, ldx, order = LAYOUT_ROW, x.shape[1], "C"
    elif x.flags.f_contiguous:
        layout, ldx, order = LAYOUT_COL, x.shape[0], "F"
    else:
        raise ValueError("X must be contiguous")
    if y is None:
        y = np.zeros((m_out, n), dtype=data.dtype, order=order)
        beta = 0.0
    ldy = y.shape[1] if layout == LAYOUT_ROW else y.shape[0]
    st = getattr(lib(), f"orc_spmm_{suf}")(
        ctypes.c_int(op), ctypes.c_double(alpha),
        ctypes.c_int64(a_csr.shape[0]), ctypes.c_int64(a_csr.shape[1]),
        _p(indptr), _p(indices), _p(data), ctypes.c_int(layout),
        _p(x), ctypes.c_int64(n), ctypes.c_int64(ldx), ctypes.c_double(beta),
        _p(y), ctypes.c_int64(ldy),
    )
    _check(st, "spmm")
    return y


def np_spmm(a, x, alpha=1.0, beta=0.0, y=None, op=OP_N):
    prod = (a.T if op == OP_T else a) @ x
    prod = np.asarray(prod)
    if y is None or beta == 0:
        return alpha * prod
    return alpha * prod + beta * y


# ------------------------------------------------------------------- SpGEMM
def c_spgemm(a_csr, b_csr, upper=False, sort=False):
    """C = A @ B (CSR x CSR -> CSR), structural nonzeros kept (MKL
    convention), first-touch column order unless ``sort``."""
    a_ptr, a_idx, a_val = _csr_parts(a_csr)
    b_ptr, b_idx, b_val = _csr_parts(b_csr)
    suf = _suffix(a_val.dtype)
    if b_val.dtype != a_val.dtype:
        raise TypeError("dtype mismatch")
    m, n = a_csr.shape[0], b_csr.shape[1]
    c_ptr = np.zeros(m + 1, dtype=np.int64)
    L = lib()
    _check(
        L.orc_spgemm_count(
            ctypes.c_int64(m), ctypes.c_int64(n), _p(a_ptr), _p(a_idx),
            _p(b_ptr), _p(b_idx), ctypes.c_int(int(upper)), _p(c_ptr)),
        "spgemm_count",
    )
    nnz = int(c_ptr[-1])
    c_idx = np.empty(nnz, dtype=np.int32)
    c_val = np.empty(nnz, dtype=a_val.dtype)
    _check(
        getattr(L, f"orc_spgemm_fill_{suf}")(
            ctypes.c_int64(m), ctypes.c_int64(n), _p(a_ptr), _p(a_idx), _p(a_val),
            _p(b_ptr), _p(b_idx), _p(b_val), ctypes.c_int(int(upper)),
            _p(c_ptr), _p(c_idx), _p(c_val)),
        "spgemm_fill",
    )
    if sort:
        _check(
            getattr(L, f"orc_order_{suf}")(ctypes.c_int64(m), _p(c_ptr), _p(c_idx), _p(c_val)),
            "order",
        )
    return _mk_csr((m, n), c_ptr, c_idx, c_val)


def np_spgemm(a, b, sort=True):
    c = (a @ b).tocsr()
    if sort:
        c.sort_indices()
    return c


def c_spmmd(a_csr, b_csr, out=None):
    """Dense row-major C = A @ B, overwriting ``out`` (_sparse_sparse.py:94-101)."""
    a_ptr, a_idx, a_val = _csr_parts(a_csr)
    b_ptr, b_idx, b_val = _csr_parts(b_csr)
    suf = _suffix(a_val.dtype)
    m, n = a_csr.shape[0], b_csr.shape[1]
    if out is None:
        out = np.empty((m, n), dtype=a_val.dtype)
    layout = LAYOUT_ROW if out.flags.c_contiguous else LAYOUT_COL
    ldc = n if layout == LAYOUT_ROW else m
    _check(
        getattr(lib(), f"orc_spmmd_{suf}")(
            ctypes.c_int64(m), ctypes.c_int64(n), _p(a_ptr), _p(a_idx), _p(a_val),
            _p(b_ptr), _p(b_idx), _p(b_val), ctypes.c_int(layout), _p(out),
            ctypes.c_int64(ldc)),
        "spmmd",
    )
    return out


# ---------------------------------------------------------- order / convert
def c_order(m_csr):
    """In-place mkl_sparse_order (_common.py:683-692)."""
    ptr = np.ascontiguousarray(m_csr.indptr, dtype=np.int64)
    idx = np.ascontiguousarray(m_csr.indices, dtype=np.int32)
    val = np.ascontiguousarray(m_csr.data)
    _check(
        getattr(lib(), f"orc_order_{_suffix(val.dtype)}")(
            ctypes.c_int64(m_csr.shape[0]), _p(ptr), _p(idx), _p(val)),
        "order",
    )
    m_csr.indices, m_csr.data = idx.astype(m_csr.indptr.dtype, copy=False), val
    return m_csr


def c_transpose(m_csr):
    """CSR(A) -> CSR(A^T): what mkl_sparse_convert_csr does to a CSC handle."""
    ptr, idx, val = _csr_parts(m_csr)
    rows, cols = m_csr.shape
    t_ptr = np.empty(cols + 1, dtype=np.int64)
    t_idx = np.empty(idx.shape[0], dtype=np.int32)
    t_val = np.empty_like(val)
    _check(
        getattr(lib(), f"orc_transpose_{_suffix(val.dtype)}")(
            ctypes.c_int64(rows), ctypes.c_int64(cols), _p(ptr), _p(idx), _p(val),
            _p(t_ptr), _p(t_idx), _p(t_val)),
        "transpose",
    )
    return _mk_csr((cols, rows), t_ptr, t_idx, t_val)


def c_bsr_to_csr(m_bsr):
    """Expand every stored block (test_mkl.py:251-268)."""
    b = m_bsr.blocksize[0]
    assert m_bsr.blocksize[0] == m_bsr.blocksize[1]
    mb = m_bsr.shape[0] // b
    bptr = np.ascontiguousarray(m_bsr.indptr, dtype=np.int64)
    bidx = np.ascontiguousarray(m_bsr.indices, dtype=np.int32)
    if m_bsr.data.flags.c_contiguous:
        bval, blayout = m_bsr.data, LAYOUT_ROW
    else:
        bval, blayout = np.ascontiguousarray(m_bsr.data), LAYOUT_ROW
    nblk = bidx.shape[0]
    indptr = np.empty(mb * b + 1, dtype=np.int64)
    indices = np.empty(nblk * b * b, dtype=np.int32)
    values = np.empty(nblk * b * b, dtype=bval.dtype)
    _check(
        getattr(lib(), f"orc_bsr_to_csr_{_suffix(bval.dtype)}")(
            ctypes.c_int64(mb), ctypes.c_int64(b), ctypes.c_int(blayout),
            _p(bptr), _p(bidx), _p(bval), _p(indptr), _p(indices), _p(values)),
        "bsr_to_csr",
    )
    return _mk_csr(m_bsr.shape, indptr, indices, values)


# --------------------------------------------------------------------- SYRK
def c_syrk(a_csr, aat=False, sort=False):
    """Upper triangle of A^T A (default) or A A^T as CSR
    (_gram_matrix.py:43-92; op mapping :35-40)."""
    at = c_transpose(a_csr)
    left, right = (a_csr, at) if aat else (at, a_csr)
    return c_spgemm(left, right, upper=True, sort=sort)


def c_syrkd(a_csr, aat=False, alpha=1.0, beta=0.0, out=None):
    """Dense upper triangle C = alpha * G + beta * C (_gram_matrix.py:104-171);
    the strict lower triangle of ``out`` is left as it was."""
    at = c_transpose(a_csr)
    left, right = (a_csr, at) if aat else (at, a_csr)
    l_ptr, l_idx, l_val = _csr_parts(left)
    r_ptr, r_idx, r_val = _csr_parts(right)
    n = left.shape[0]
    if out is None:
        out = np.zeros((n, n), dtype=l_val.dtype)
        beta = 0.0
    layout = LAYOUT_ROW if out.flags.c_contiguous else LAYOUT_COL
    _check(
        getattr(lib(), f"orc_syrkd_{_suffix(l_val.dtype)}")(
            ctypes.c_int64(n), _p(l_ptr), _p(l_idx), _p(l_val), _p(r_ptr), _p(r_idx),
            _p(r_val), ctypes.c_double(alpha), ctypes.c_double(beta),
            ctypes.c_int(layout), _p(out), ctypes.c_int64(n)),
        "syrkd",
    )
    return out


def np_gram_upper(a, aat=False):
    """What every reference gram test compares with
    (tests/test_gram_matrix.py:26-32)."""
    d = a.toarray() if sp.issparse(a) else np.asarray(a)
    g = d @ d.T if aat else d.T @ d
    return np.triu(g)


# ---------------------------------------------------------------- comparing
def canonical(m):
    """Sorted-index CSR copy: the form in which indptr / indices are compared
    bit-exactly (np.array_equal, so the integer width does not matter).
    Stored entries are kept even when their value is 0.0."""
    c = sp.csr_matrix(m, copy=True)
    c.has_sorted_indices = False
    c.sort_indices()
    return c


def value_bound(abs_a, abs_b):
    """|A|@|B|: the magnitude every entry's rounding error is relative to.
    Tolerances in tests/ are  |ours - oracle| <= tol * value_bound  with
    tol = 1e-5 (fp32) / 1e-12 (fp64) as north_star states; for the strictly
    positive BASELINE inputs this equals a plain relative tolerance."""
    prod = abs_a @ abs_b
    return prod.toarray() if sp.issparse(prod) else np.asarray(prod)

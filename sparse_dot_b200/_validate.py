"""
Input / output validation with the reference's semantics — this is the API
contract of sparse_dot_mkl, restated (not copied) from
sparse_dot_mkl/_mkl_interface/_common.py:

    check_shapes        <- _sanity_check          (:725-752)
    product_is_empty    <- _empty_output_check    (:1003-1024)
    unify_dtypes        <- _type_check            (:773-866)
    dense_layout        <- _get_numpy_layout      (:181-213)
    output_array        <- _out_matrix            (:885-955)
    precision_flags     <- _is_double             (:963-986)

All failures are ValueError, as in the reference.
"""
import numpy as np
import scipy.sparse as sps

from ._lib import LAYOUT_C, LAYOUT_F, SDB

REAL_DTYPES = (np.dtype(np.float32), np.dtype(np.float64))
COMPLEX_DTYPES = (np.dtype(np.complex64), np.dtype(np.complex128))
ALL_DTYPES = REAL_DTYPES + COMPLEX_DTYPES

# (double precision?, complex?) -> numpy dtype   (_common.py:60-66)
OUTPUT_DTYPES = {
    (False, False): np.float32,
    (True, False): np.float64,
    (False, True): np.complex64,
    (True, True): np.complex128,
}

_CSR_TYPES = tuple(t for t in (getattr(sps, "csr_matrix", None), getattr(sps, "csr_array", None)) if t)
_CSC_TYPES = tuple(t for t in (getattr(sps, "csc_matrix", None), getattr(sps, "csc_array", None)) if t)
_BSR_TYPES = tuple(t for t in (getattr(sps, "bsr_matrix", None), getattr(sps, "bsr_array", None)) if t)


def debug_print(msg):
    if SDB.DEBUG:
        print(msg)


def is_csr(x):
    return isinstance(x, _CSR_TYPES)


def is_csc(x):
    return isinstance(x, _CSC_TYPES)


def is_bsr(x):
    return isinstance(x, _BSR_TYPES)


def is_supported_sparse(x):
    """Dense inputs pass; sparse ones must be CSR, CSC or BSR (COO etc. do not)."""
    return (not sps.issparse(x)) or is_csr(x) or is_csc(x) or is_bsr(x)


def sparse_container(x):
    """(constructor, format tag) of the scipy class of ``x``: results come back
    in the container of the left operand (_common.py:228-242)."""
    for tag, group in (("csr", _CSR_TYPES), ("csc", _CSC_TYPES), ("bsr", _BSR_TYPES)):
        for cls in group:
            if isinstance(x, cls):
                return cls, tag
    raise ValueError("Input matrices to dot_product_mkl must be CSR, CSC, or BSR; COO is not supported")


def is_dense_vector(x):
    if sps.issparse(x):
        return False
    return x.ndim == 1 or (x.ndim == 2 and min(x.shape) == 1)


def check_shapes(a, b, allow_vector=False):
    """2-D (or vector, when allowed) operands whose inner dimensions agree."""
    a2, b2 = a.ndim == 2, b.ndim == 2
    if not allow_vector and not (a2 and b2):
        raise ValueError(f"Matrices must be 2d: {a.shape} * {b.shape} is not valid")
    dims_ok = (a2 or is_dense_vector(a)) and (b2 or is_dense_vector(b))
    inner_a = a.shape[0] if a.ndim == 1 else a.shape[1]
    if not dims_ok or inner_a != b.shape[0]:
        raise ValueError(f"Matrix alignment error: {a.shape} * {b.shape} is not valid")


def product_is_empty(a, b):
    """True when the product is trivially all-zero / zero-sized."""
    if min(tuple(a.shape) + tuple(b.shape)) == 0:
        return True
    for m in (a, b):
        if sps.issparse(m) and min(m.data.size, m.indices.size) == 0:
            return True
    return False


def _as(m, dtype):
    return m if m.dtype == dtype else m.astype(dtype)


def unify_dtypes(a, b=None, cast=False, allow_complex=True):
    """Both operands end up float32, float64, complex64 or complex128 and equal.
    Without ``cast`` anything else is an error; with it, reals go to float64,
    complex to complex128, and a real operand follows a valid complex partner."""
    n_cplx = int(np.iscomplexobj(a)) + int(np.iscomplexobj(b))
    if not allow_complex and n_cplx:
        raise ValueError("Complex datatypes are not supported")

    if b is None:
        if a.dtype in ALL_DTYPES:
            return a
        if not cast:
            raise ValueError(
                f"Matrix data type must be float32, float64, csingle, or cdouble; {a.dtype} provided"
            )
        return _as(a, np.complex128 if n_cplx else np.float64)

    if a.dtype in ALL_DTYPES and a.dtype == b.dtype:
        return a, b
    if not cast:
        raise ValueError(
            "Matrix data type must be float32, float64, csingle, or cdouble, and must be the same "
            f"if cast=False; {a.dtype} & {b.dtype} provided"
        )
    if n_cplx == 0:
        target = np.float64
    elif n_cplx == 2:
        target = np.complex128
    elif a.dtype in COMPLEX_DTYPES:
        target = a.dtype
    elif b.dtype in COMPLEX_DTYPES:
        target = b.dtype
    else:
        target = np.complex128
    debug_print(f"Recasting matrix data types {a.dtype} and {b.dtype} to {np.dtype(target)}")
    return _as(a, target), _as(b, target)


def precision_flags(m):
    """(is double precision, is complex) of a supported dtype."""
    dt = np.dtype(m.dtype)
    if dt not in ALL_DTYPES:
        raise ValueError("Only float32, float64, csingle, and cdouble dtypes are supported")
    return dt in (np.dtype(np.float64), np.dtype(np.complex128)), dt in COMPLEX_DTYPES


def dense_layout(arr, other=None):
    """(layout code, leading dimension) of a contiguous 2-D array.  An array
    that is both C and F contiguous (one row / one column) follows ``other``."""
    c, f = arr.flags.c_contiguous, arr.flags.f_contiguous
    if c and f and other is not None:
        if other.flags.c_contiguous:
            return LAYOUT_C, arr.shape[1]
        if other.flags.f_contiguous:
            return LAYOUT_F, arr.shape[0]
    if c:
        return LAYOUT_C, arr.shape[1]
    if f:
        return LAYOUT_F, arr.shape[0]
    raise ValueError("Array is not contiguous")


def output_array(shape, dtype, order="C", out=None, out_t=False, zero=True):
    """A fresh output array, or the caller's ``out`` after checking that its
    shape, dtype, memory order and contiguity are exactly what the product
    needs.  ``out_t`` only flips how a mismatch is reported (the caller passed
    out.T).  ``zero=False`` skips the zero fill when the kernel overwrites."""
    if out is None:
        return np.zeros(shape, dtype=dtype, order=order) if zero else np.empty(shape, dtype=dtype, order=order)

    order_ok = out.flags["C_CONTIGUOUS"] if order == "C" else out.flags["F_CONTIGUOUS"]
    if tuple(shape) == out.shape and np.dtype(dtype) == out.dtype and order_ok and out.data.contiguous:
        return out

    c, f = out.flags["C_CONTIGUOUS"], out.flags["F_CONTIGUOUS"]
    if out_t and out.ndim != 1:
        got_shape, need_shape = out.shape[::-1], tuple(shape)[::-1]
        got_order = "F" if (c and not f) else "C"
        need_order = "F" if order == "C" else "C"
    else:
        got_shape, need_shape = out.shape, tuple(shape)
        got_order = "C" if c else "F"
        need_order = order
    contig = "CONTIGUOUS" if out.data.contiguous else "NONCONTIGUOUS"
    raise ValueError(
        f"Provided out array is {got_shape} {out.dtype} [{got_order}_{contig}] and product requires "
        f"{need_shape} {np.dtype(dtype).name} [{need_order}_CONTIGUOUS]"
    )

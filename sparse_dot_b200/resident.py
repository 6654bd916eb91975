"""
Device-resident operands: keep a sparse matrix in HBM across calls and
multiply it by host arrays (only the dense panels cross PCIe) or by arrays that already live on
the GPU (torch CUDA tensors or anything exposing ``__cuda_array_interface__`` — nothing is copied).

The reference has no equivalent: MKL handles borrow host arrays and are rebuilt on every
``dot_product_mkl`` call (_sparse_dense.py:75-132).  On a GPU the upload of A dominates a single
call, so iterative callers (repeated SpMM with a fixed A, gram -> solve) want this.
"""
import ctypes as _ct

import numpy as np

from . import _handles as _h
from . import _lib
from . import _validate as _v
from ._lib import SDB, check, scalar_pair


def _device_pointer(obj):
    """(pointer, shape, dtype, row-major?) of a GPU array: torch CUDA tensor or __cuda_array_interface__."""
    if hasattr(obj, "data_ptr") and getattr(obj, "is_cuda", False):
        if not obj.is_contiguous():
            raise ValueError("Array is not contiguous")
        dt = np.dtype(str(obj.dtype).replace("torch.", ""))
        return int(obj.data_ptr()), tuple(obj.shape), dt
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            raise ValueError("Array is not contiguous")
        return int(cai["data"][0]), tuple(cai["shape"]), np.dtype(cai["typestr"])
    raise TypeError("expected a CUDA tensor / array (data_ptr() or __cuda_array_interface__)")


class ResidentCSR:
    """A scipy CSR / CSC / BSR matrix uploaded once.  Use as a context manager or call close()."""

    def __init__(self, matrix, _handle=None):
        if _handle is not None:
            self.handle = _handle
            meta = _h.info(_handle)
            b = meta["block"]
            self.shape = (meta["rows"] * b, meta["cols"] * b)
            self.dtype = meta["dtype"]
            return
        if not _v.is_supported_sparse(matrix) or not hasattr(matrix, "indptr"):
            raise ValueError("Matrix is not CSC, CSR, or BSR")
        matrix = _v.unify_dtypes(matrix)
        self.handle, _, _ = _h.create(matrix)
        self.shape = tuple(matrix.shape)
        self.dtype = np.dtype(matrix.dtype)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self.handle:
            self.handle.destroy()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    @property
    def nnz(self):
        meta = _h.info(self.handle)
        return meta["nnz"] * meta["block"] ** 2

    # ------------------------------------------------------------------ products with dense panels
    def dot(self, x, out=None, out_scalar=None, transpose=False):
        """op(A) @ x for a host array ``x`` (C or F ordered); same ``out`` / ``out_scalar`` rules as
        dot_product_mkl.  Only x (and ``out`` when it is accumulated into) is uploaded."""
        x = np.asarray(x)
        if x.ndim == 1:
            # a vector (what a CG / FGMRES loop passes): one column, the result comes back 1-D as well; from the
            # second product on a large matrix runs the shared-memory SpMV (csrc/spmv_tile.cu)
            if out is not None and out.ndim != 1:
                raise ValueError("out must be a vector when x is one")
            res = self.dot(x.reshape(-1, 1), out=None if out is None else out.reshape(-1, 1), out_scalar=out_scalar,
                           transpose=transpose)
            return out if out is not None else res.ravel()
        if x.ndim != 2 or x.shape[0] != self.shape[0 if transpose else 1]:
            raise ValueError(f"Matrix alignment error: {self.shape} * {x.shape} is not valid")
        if x.dtype != self.dtype:
            raise ValueError(f"Matrix data types must match: {self.dtype} & {x.dtype} provided")
        m = self.shape[1 if transpose else 0]
        layout, ldx = _v.dense_layout(x, other=out)
        order = "C" if layout == _lib.LAYOUT_C else "F"
        result = _v.output_array((m, x.shape[1]), self.dtype, order, out=out, zero=False)
        _, ldy = _v.dense_layout(result, other=x)
        beta = 0.0 if out is None else (1.0 if out_scalar is None else out_scalar)
        check(
            SDB.lib.sdb_spmm(_lib.OP_T if transpose else _lib.OP_N, scalar_pair(1.0), self.handle.ref, layout,
                             x.ctypes.data_as(_ct.c_void_p), x.shape[1], ldx, scalar_pair(beta),
                             result.ctypes.data_as(_ct.c_void_p), ldy),
            "sdb_spmm",
        )
        return result

    def dot_device(self, x, out, alpha=1.0, beta=0.0, transpose=False, stream=None):
        """out = alpha * op(A) @ x + beta * out with ``x`` and ``out`` row-major arrays already on the
        GPU.  Stream-ordered on ``stream`` (a cudaStream_t as int, or a torch stream); does not
        synchronise.  Returns ``out``."""
        xp, xs, xd = _device_pointer(x)
        yp, ys, yd = _device_pointer(out)
        m = self.shape[1 if transpose else 0]
        k = self.shape[0 if transpose else 1]
        if len(xs) != 2 or len(ys) != 2 or xs[0] != k or ys != (m, xs[1]):
            raise ValueError(f"Matrix alignment error: {self.shape} * {xs} -> {ys} is not valid")
        if xd != self.dtype or yd != self.dtype:
            raise ValueError(f"Matrix data types must match: {self.dtype}, {xd}, {yd} provided")
        sp = getattr(stream, "cuda_stream", stream)
        check(
            SDB.lib.sdb_spmm_dev(_lib.OP_T if transpose else _lib.OP_N, scalar_pair(alpha), self.handle.ref,
                                 _lib.LAYOUT_C, _ct.c_void_p(xp), xs[1], xs[1], scalar_pair(beta), _ct.c_void_p(yp),
                                 xs[1], _ct.c_void_p(sp) if sp else None),
            "sdb_spmm_dev",
        )
        return out

    # ------------------------------------------------------------------ sparse results stay resident
    def matmat(self, other, reorder_output=False):
        """A @ B as a new ResidentCSR (nothing leaves the GPU)."""
        if not isinstance(other, ResidentCSR):
            raise TypeError("matmat takes another ResidentCSR")
        ref = _ct.c_void_p()
        fn = "sdb_spgemm_ordered" if reorder_output else "sdb_spgemm"
        check(getattr(SDB.lib, fn)(_lib.OP_N, self.handle.ref, other.handle.ref, _ct.byref(ref)), fn)
        return ResidentCSR(None, _handle=_h.Handle(ref, self.dtype))

    def gram(self, transpose=False, reorder_output=False):
        """Upper triangle of A^T A (or A A^T) as a new ResidentCSR."""
        ref = _ct.c_void_p()
        fn = "sdb_syrk_ordered" if reorder_output else "sdb_syrk"
        check(getattr(SDB.lib, fn)(_lib.OP_N if transpose else _lib.OP_T, self.handle.ref, _ct.byref(ref)), fn)
        return ResidentCSR(None, _handle=_h.Handle(ref, self.dtype))

    def to_scipy(self):
        meta = _h.info(self.handle)
        tag = {_lib.FMT_CSR: "csr", _lib.FMT_CSC: "csc", _lib.FMT_BSR: "bsr"}[meta["format"]]
        return _h.export(self.handle, output_type=f"{tag}_matrix")

#!/usr/bin/env python
"""
spmm_rmat_probe.py — the row-gather SpMM (K1) on a power-law matrix: R-MAT scale 20, edge factor 8, fp32, times a
dense panel of 128 columns, operands in HBM; beside it a matrix of the same size with uniform rows.  Shows what the
longest rows cost a kernel that gives every row to one warp.  Diagnostic; feeds DESIGN.md.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sparse_dot_b200 as sdb  # noqa: E402
from sparse_dot_b200 import _handles as H  # noqa: E402
from sparse_dot_b200 import _lib  # noqa: E402
from tests import _cases as cs  # noqa: E402

lib = _lib.SDB.lib


def run(a, n, label):
    rng = np.random.default_rng(1)
    x = rng.random((a.shape[1], n), dtype=np.float32)
    ha, _, _ = H.create(a)
    with ha:
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.sdb_dev_alloc(C.byref(d_x), x.nbytes), "alloc")
        _lib.check(lib.sdb_dev_alloc(C.byref(d_y), a.shape[0] * n * 4), "alloc")
        try:
            _lib.check(lib.sdb_memcpy(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes, 1), "memcpy")
            one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)
            call = lambda: _lib.check(lib.sdb_spmm_dev(_lib.OP_N, one, ha.ref, _lib.LAYOUT_C, d_x, n, n, zero, d_y, n,
                                                       None), "sdb_spmm_dev")
            times = []
            for _ in range(6):
                _lib.check(lib.sdb_device_synchronize(), "sync")
                t0 = time.perf_counter()
                call()
                _lib.check(lib.sdb_device_synchronize(), "sync")
                times.append((time.perf_counter() - t0) * 1e3)
            lens = np.diff(a.indptr)
            r = int(np.argmax(lens))
            got = np.empty(n, dtype=np.float32)
            _lib.check(lib.sdb_memcpy(got.ctypes.data_as(C.c_void_p), C.c_void_p(d_y.value + r * n * 4), n * 4, 2), "memcpy")
            want = (a[r].astype(np.float64) @ x.astype(np.float64)).ravel()
            print(json.dumps({"matrix": label, "rows": a.shape[0], "nnz": int(a.nnz), "longest_row": int(lens.max()),
                              "n_dense": n, "ms_by_call": [round(t, 3) for t in times], "kernel": sdb.last_spmm_kernel(),
                              "longest_row_rel_err": float(np.abs(got - want).max() / np.abs(want).max())}), flush=True)
        finally:
            lib.sdb_dev_free(d_x)
            lib.sdb_dev_free(d_y)


def main():
    g = cs.rmat_csr(20, 8, np.float32, seed=1)
    run(g, 128, "R-MAT scale 20 ef 8")
    u = cs.uniform_rows_csr(g.shape[0], g.shape[1], max(1, g.nnz // g.shape[0]), np.float32, seed=2)
    run(u, 128, "uniform rows, same size")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""K2 decided by measurement: BSR x dense with the register-tiled FMA kernel vs the tensor-core
kernel (DMMA for fp64, 3xTF32 for fp32) on the configs[4] BSR-16 variant and on b = 32 / N = 512 — milliseconds,
block-model GB/s and max relative error against a float64 host recomputation of sampled rows.
    python scripts/bsr_mma_table.py > profiles/r2_bsr_mma_table.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_dot_b200 as sdb  # noqa: E402
from sparse_dot_b200 import _handles as H  # noqa: E402
from sparse_dot_b200 import _lib  # noqa: E402

lib = _lib.SDB.lib


def blocked(block_rows, per_row, b, dtype, seed=5):
    rng = np.random.default_rng(seed)
    cols = (rng.integers(0, block_rows, size=(block_rows, per_row)) + np.arange(per_row)[None, :] * 7919) % block_rows
    cols.sort(axis=1)
    for _ in range(4):
        dup = np.zeros(cols.shape, dtype=bool)
        dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
        if not dup.any():
            break
        cols[dup] = rng.integers(0, block_rows, size=int(dup.sum()))
        cols.sort(axis=1)
    indptr = np.arange(0, block_rows * per_row + 1, per_row, dtype=np.int32)
    data = (rng.random((block_rows * per_row, b, b), dtype=np.float32) + 0.5).astype(dtype)
    a = sp.bsr_matrix((data, cols.ravel().astype(np.int32), indptr), shape=(block_rows * b, block_rows * b))
    a.sum_duplicates()
    return a


def run(block_rows, per_row, b, n, dtype):
    a = blocked(block_rows, per_row, b, dtype)
    rng = np.random.default_rng(1)
    x = rng.random((a.shape[1], n), dtype=np.float32).astype(dtype)
    es = np.dtype(dtype).itemsize
    nblk = a.indices.shape[0]
    model = nblk * (b * b * es + 4 + b * n * es) + a.shape[0] * n * es
    out = {"block": b, "n": n, "dtype": np.dtype(dtype).name, "rows": a.shape[0], "blocks": int(nblk),
           "block_model_gbytes": model / 1e9, "gflop": 2.0 * nblk * b * b * n / 1e9}
    rows = np.sort(rng.choice(a.shape[0], size=24, replace=False))
    want = a.tocsr()[rows].astype(np.float64) @ x.astype(np.float64)
    ha, _, _ = H.create(a)
    with ha:
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.sdb_dev_alloc(C.byref(d_x), x.nbytes), "alloc")
        _lib.check(lib.sdb_dev_alloc(C.byref(d_y), a.shape[0] * n * es), "alloc")
        try:
            _lib.check(lib.sdb_memcpy(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes, 1), "memcpy")
            one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)
            for label, mma in (("fma", 0), ("mma", 1)):
                _lib.set_option("bsr_mma", mma)

                def call():
                    _lib.check(lib.sdb_spmm_dev(_lib.OP_N, one, ha.ref, _lib.LAYOUT_C, d_x, n, n, zero, d_y, n, None),
                               "sdb_spmm_dev")
                call()
                _lib.check(lib.sdb_device_synchronize(), "sync")
                t0 = time.perf_counter()
                for _ in range(10):
                    call()
                _lib.check(lib.sdb_device_synchronize(), "sync")
                ms = (time.perf_counter() - t0) / 10 * 1e3
                worst = 0.0
                for j, r in enumerate(rows):
                    got = np.empty(n, dtype=dtype)
                    _lib.check(lib.sdb_memcpy(got.ctypes.data_as(C.c_void_p), C.c_void_p(d_y.value + int(r) * n * es),
                                              n * es, 2), "memcpy")
                    worst = max(worst, float(np.max(np.abs(got - want[j]) / np.abs(want[j]))))
                out[label] = {"ms": ms, "kernel": sdb.last_spmm_kernel(), "block_model_gbs": model / (ms * 1e-3) / 1e9,
                              "tflops": out["gflop"] / ms, "max_rel_err": worst}
        finally:
            _lib.set_option("bsr_mma", -1)
            lib.sdb_dev_free(d_x)
            lib.sdb_dev_free(d_y)
    return out


if __name__ == "__main__":
    table = []
    for args in [(62_500, 4, 16, 256, np.float32), (62_500, 4, 16, 256, np.float64), (31_250, 4, 32, 512, np.float32),
                 (31_250, 4, 32, 512, np.float64), (31_250, 16, 32, 256, np.float32), (125_000, 4, 8, 128, np.float32)]:
        r = run(*args)
        table.append(r)
        print(json.dumps(r), flush=True)

#!/usr/bin/env python
"""
small_latency.py — wall-clock latency of the public call on small problems (BASELINE configs[0] and neighbours),
where launch / allocation / copy overheads, not bandwidth, decide: median of 30 calls of dot_product_mkl
(scipy CSR, numpy array in pageable memory) beside scipy's own `@` on the box's host, with the library's phase
timers (upload, kernel, download).  Not a bench line; feeds DESIGN.md.
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sparse_dot_b200 as sdb  # noqa: E402


def median_ms(fn, reps=30):
    fn()
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    # (scipy.sparse.random draws without replacement from n*n positions: fine up to 1e8, hopeless at 1e10 — the
    # larger sizes belong to bench.py / run_configs.py, which build their matrices row by row)
    cases = [(1_000, 1e-2, 16, np.float64), (10_000, 1e-3, 64, np.float64), (10_000, 1e-3, 64, np.float32)]
    for n, d, k, dt in cases:
        a = sp.random(n, n, density=d, format="csr", dtype=dt, random_state=86)
        x = np.random.default_rng(88).random((n, k)).astype(dt)
        out = np.zeros((n, k), dtype=dt)
        rec = {"rows": n, "nnz": int(a.nnz), "n_dense": k, "dtype": np.dtype(dt).name}
        rec["call_ms"], rec["call_min_ms"] = median_ms(lambda: sdb.dot_product_mkl(a, x))
        rec["phases_ms"] = [round(v, 4) for v in sdb.last_timing_ms()]
        rec["call_out_ms"], _ = median_ms(lambda: sdb.dot_product_mkl(a, x, out=out, out_scalar=0.0))
        rec["scipy_ms"], _ = median_ms(lambda: a @ x, reps=10)
        with sdb.optimize(a) as h:
            rec["resident_call_ms"], _ = median_ms(lambda: sdb.dot_product_mkl(h, x))
        a2 = sp.random(n, n, density=d, format="csr", dtype=dt, random_state=87)
        rec["spgemm_call_ms"], _ = median_ms(lambda: sdb.dot_product_mkl(a, a2), reps=10)
        rec["spgemm_scipy_ms"], _ = median_ms(lambda: a @ a2, reps=5)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()

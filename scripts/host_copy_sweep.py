#!/usr/bin/env python
"""Host side of the pageable-memory pipeline on THIS box: copy-pool bandwidth (sdb_probe_bandwidth kind 3) and
the end-to-end pageable dot_product_mkl call for a few thread counts / slot sizes (one subprocess each: the
settings are read once per process).  python scripts/host_copy_sweep.py [--e2e]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
sys.path.insert(0, %r)
import numpy as np
from sparse_dot_b200 import _lib
import sparse_dot_b200 as sdb
out = {"threads": os.environ.get("SDB_COPY_THREADS"), "slot_mb": os.environ.get("SDB_STAGE_SLOT_MB"),
       "piece_kb": os.environ.get("SDB_COPY_PIECE_KB"), "nt_stores": os.environ.get("SDB_COPY_NT"), "host_copy_gbs": _lib.probe_bandwidth(3, 1 << 30, 3)}
if "--e2e" in sys.argv:
    from tests import _cases as cs
    a, x, y0 = cs.c2_workload(1_000_000, 1_000_000, 50, 128, seed=0)
    y = y0.copy()
    sdb.dot_product_mkl(a, x, out=y, out_scalar=0.5)
    t0 = time.perf_counter()
    for _ in range(4):
        sdb.dot_product_mkl(a, x, out=y, out_scalar=0.5)
    out["pageable_e2e_ms"] = (time.perf_counter() - t0) / 4 * 1e3
    out["spans_ms"] = sdb.last_timing_ms()
print(json.dumps(out))
''' % ROOT

if __name__ == "__main__":
    e2e = ["--e2e"] if "--e2e" in sys.argv else []
    print(json.dumps({"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}))
    ncpu = len(os.sched_getaffinity(0))
    for threads, slot, piece, nt in [(8, 8, 1024, 1), (12, 8, 1024, 1), (ncpu, 8, 1024, 1), (ncpu, 8, 1024, 0),
                                     (ncpu, 16, 1024, 1), (ncpu, 4, 512, 1), (ncpu, 8, 512, 1), (ncpu, 8, 2048, 1)]:
        env = dict(os.environ, SDB_COPY_THREADS=str(threads), SDB_STAGE_SLOT_MB=str(slot), SDB_COPY_PIECE_KB=str(piece),
                   SDB_COPY_NT=str(nt))
        r = subprocess.run([sys.executable, "-c", CHILD] + e2e, env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)

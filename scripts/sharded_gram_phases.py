#!/usr/bin/env python
"""
sharded_gram_phases.py — where the wall clock of sharded.gram_dense_sharded / spgemm_sharded_device goes (run under
torchrun, NCCL): the same steps as the library functions with a timer after each.  Diagnostic only.
"""
import ctypes as ct
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sparse_dot_b200 import _handles as _h  # noqa: E402
from sparse_dot_b200 import _lib, sharded  # noqa: E402
from sparse_dot_b200._lib import SDB, check, scalar_pair  # noqa: E402
from tests import _cases as cs  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    check(SDB.lib.sdb_set_device(int(os.environ["LOCAL_RANK"])), "sdb_set_device")
    dist.init_process_group("nccl")
    m = cs.uniform_rows_csr(400_000, 20_000, 100, np.float32, seed=4)
    n = m.shape[1]
    for rep in range(3):
        marks = []

        def mark(name):
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

        dist.barrier()
        mark("start")
        bounds = sharded.partition_rows(m.indptr, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        block = sharded.row_block(m, lo, hi)
        owners = sharded.triangle_row_bounds(n, world)
        mark("partition")
        panel = torch.zeros((n, n), dtype=torch.float32, device="cuda")
        mark("zeros")
        handle, _, _ = _h.create(block)
        mark("upload")
        with handle:
            check(SDB.lib.sdb_syrkd_dev(_lib.OP_T, handle.ref, scalar_pair(1.0), scalar_pair(0.0),
                                        ct.c_void_p(panel.data_ptr()), _lib.LAYOUT_C, n, None), "syrkd")
            check(SDB.lib.sdb_device_synchronize(), "sync")
            mark("syrkd")
        mark("destroy")
        for q in range(world):
            r0, r1 = owners[q], owners[q + 1]
            if r1 > r0:
                dist.reduce(panel[r0:r1], dst=q)
        mark("reduce")
        out = sharded._download_upper(panel, owners[rank], owners[rank + 1])
        mark("download")
        del panel
        mark("free")
        rec = {"rank": rank, "rep": rep, "rows": int(out.shape[0])}
        for (a, ta), (b, tb) in zip(marks[:-1], marks[1:]):
            rec[b] = round((tb - ta) * 1e3, 2)
        rec["total"] = round((marks[-1][1] - marks[0][1]) * 1e3, 2)
        print(json.dumps(rec), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
import sparse_dot_b200 as sdb
from sparse_dot_b200 import sharded, _lib
from tests import _cases as cs
a, x, y0 = cs.c2_workload(1_000_000, 1_000_000, 50, 64, seed=0)
for slab in ("auto",):
    with sharded.RowShardedSpMM(a, 64) as plan:
        plan.set_x(x); plan.set_local_y(y0)
        for _ in range(4):
            plan.run(alpha=1.0, beta=0.5)
        plan.synchronize()
        t0=time.perf_counter()
        for _ in range(20):
            plan.run(alpha=1.0, beta=0.5)
        plan.synchronize()
        ms=(time.perf_counter()-t0)/20*1e3
        g = a.nnz*(8+256) + a.shape[0]*256*2
        print(sdb.last_spmm_kernel(), f"{ms:.3f} ms  {g/ms/1e6:.0f} GB/s effective")

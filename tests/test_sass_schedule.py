"""
Guard for the streaming SpMM's hot loop (csrc/spmm_slab.cu).  At 32 registers ptxas's schedule of the
rolling gather ring is sensitive to unrelated source changes: in one A/B on the same B200 an edit to the
epilogue made ptxas pair the refills (two gathers back to back, then two consumes), halving the loads in
flight, and the same loop ran 3.91 ms instead of 2.56 ms on BASELINE configs[1].  The two schedules are easy
to tell apart in the SASS without a GPU: in the good one the 16-byte gathers (LDG.E.128) of the unrolled
32-entry chunk are evenly spaced.  CPU only (nvcc's cuobjdump reads the object the build left in-tree).
"""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "sparse_dot_b200", "csrc", "_obj", "spmm_slab.o")


def _kernel_sass(pattern):
    out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    keep, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = pattern in line
        elif on and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            keep.append(re.sub(r"/\*[0-9a-fx]*\*/", "", line).strip())
    return keep


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(OBJ),
                    reason="needs cuobjdump and the in-tree object file")
@pytest.mark.parametrize("dtype_code", ["f", "d"])
@pytest.mark.parametrize("keep", [0, 1])
def test_default_streaming_kernel_keeps_its_gather_ring_rolling(dtype_code, keep):
    # <T, RPW=6, WARPS=32, U=2, CTAS=2, KEEP>: the plain gathers and the ones with an L2 evict_last policy
    sass = _kernel_sass(f"spmm_stream_kernelI{dtype_code}Li6ELi32ELi2ELi2ELb{keep}E")
    assert sass, "default streaming kernel not found in spmm_slab.o"
    at = [i for i, ins in enumerate(sass) if "LDG.E.128" in ins]
    gaps = collections.Counter(b - a for a, b in zip(at, at[1:]))
    spacing, count = gaps.most_common(1)[0]
    # 32 gathers per chunk: at least 24 of the 31 gaps identical and short (one entry's worth of instructions)
    assert count >= 24 and spacing <= 16, dict(gaps)
    assert not any("LDL" in ins or "STL" in ins for ins in sass[at[0]:at[31]]), "spill inside the unrolled chunk"

#!/usr/bin/env python
"""Schedule check for the streaming SpMM (csrc/spmm_slab.cu) without a GPU: for every spmm_stream_kernel
instantiation in the built object, the histogram of instruction distances between consecutive 16-byte gathers
(LDG.E.128) and the number of spill instructions.  A healthy rolling ring shows ~31 equal, short gaps per
unrolled chunk (see tests/test_sass_schedule.py and profiles/README.md, round 1e).

    python scripts/sass_gaps.py [path/to/spmm_slab.o]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sparse_dot_b200", "csrc", "_obj", "spmm_slab.o")
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
kernels, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1) if "spmm_stream_kernel" in m.group(1) else None
        if name:
            kernels[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        kernels[name].append(re.sub(r"/\*[0-9a-fx]*\*/", "", line).strip())
for name, sass in kernels.items():
    short = re.sub(r".*spmm_stream_kernelI(\w+?)EEv.*", r"\1", name)
    at = [i for i, ins in enumerate(sass) if "LDG.E.128" in ins]
    gaps = collections.Counter(b - a for a, b in zip(at, at[1:]))
    spills = sum("LDL" in i or "STL" in i for i in sass)
    (spacing, count) = gaps.most_common(1)[0] if gaps else (0, 0)
    verdict = "rolling" if count >= 24 and spacing <= 16 else "IRREGULAR"
    print(f"{short:28s} {len(sass):5d} instr, {spills:3d} spill ops, gathers {len(at):3d}, "
          f"mode gap {spacing} x{count}  -> {verdict}   {dict(sorted(gaps.items()))}")

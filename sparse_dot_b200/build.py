"""
Build libsdb200.so (the sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

    python -m sparse_dot_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so lands next to this file so it travels
with the source tree.  There is no other backend and no CPU fallback: if the
library is missing, importing sparse_dot_b200 raises ImportError (the analogue
of the reference's "mkl_rt not found", _mkl_interface/_load_library.py:83-94).
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libsdb200.so")
SHIM_SRC = os.path.join(CSRC, "mkl_shim.cpp")
SHIM_LIB = os.path.join(HERE, "libsdb200_mkl.so")  # oneMKL symbol names over libsdb200 (see csrc/mkl_shim.cpp)

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-Xcompiler", "-Wno-format-truncation",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    inc = os.path.join(os.path.dirname(HERE), "include", "sdb200.h")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [inc]
    return max(os.path.getmtime(h) for h in hs)


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libsdb200.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr = _headers_mtime()
    stale = []
    for s in srcs:
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(
            os.path.getmtime(os.path.join(CSRC, s)), hdr
        ):
            stale.append(s)
    if stale:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(stale))) as ex:
            list(ex.map(lambda s: _compile(nvcc, s, verbose), stale))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if stale or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-Xcompiler", "-fPIC", "-cudart", "static", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_shim(force)
    return LIB


def build_shim(force=False):
    """libsdb200_mkl.so: plain C++ (g++), links libsdb200.so from its own directory ($ORIGIN)."""
    if not (force or not os.path.exists(SHIM_LIB)
            or os.path.getmtime(SHIM_LIB) < max(os.path.getmtime(SHIM_SRC), os.path.getmtime(LIB))):
        return SHIM_LIB
    gxx = shutil.which("g++") or "/usr/bin/g++"
    cmd = [gxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-o", SHIM_LIB, SHIM_SRC,
           "-L" + HERE, "-l:libsdb200.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"shim build failed:\n{r.stdout}\n{r.stderr}")
    return SHIM_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

#!/usr/bin/env python
"""
bench.py — headline benchmark of the sparse-matmul hot path (BASELINE.json):

    SpMM effective HBM GB/s (and GFLOP/s) at 1/2/4/8 B200 vs MKL CPU

Workload (BASELINE.json configs[1]):  CSR(1M x 1M, 50 nnz/row, fp32) x dense(1M x 128),
`out=` accumulate path (beta = 0.5).  One step = one pass Y := A @ X + 0.5 * Y.

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a CUDA)
    python bench.py --impl reference --steps K --warmup W    # the reference's MKL path on host cores

N > 1 is launched by torch.distributed.run, one rank per GPU: every rank owns one 1M-row block of a
(N x 1M)-row product (weak scaling), X is replicated, and one library call per step leaves the rank's rows
in every rank's full output panel over NVLink peer mappings (the fused all-gather of SURVEY.md §8e; the
exchange strategy is picked by timing the candidates during warm-up); `--allgather nccl` runs kernel +
ncclAllGather instead, `--allgather none` skips the exchange.  After the timed region every rank checks rows
of EVERY peer's block in its own copy of the panel against a float64 recomputation.

ONE JSON line (rank 0):
  value       whole-job gather-model GB/s, operands resident in HBM, CUDA events on the launching stream;
  e2e         the public call dot_product_mkl(csr, ndarray, out=, out_scalar=) on page-locked host arrays, H2D and
              D2H inside the timed region; e2e.pageable = the same call on ordinary numpy arrays;
  roofline    of the dominant kernel: `traffic` = ncu DRAM bytes per launch (profiles/kernel_traffic.json; only
              when the SASS of that kernel is the one that was profiled), `frac` = traffic / time / measured HBM
              peak, `effective_*` = the same with the algorithmic (gather-model) bytes of SURVEY.md §8d,
              `l2_to_sm` = the binding resource of the L2-tiled kernel against an L2 read peak probed in this run;
  inspector   one-time cost of the handle's slab-ordered copy (ms, extra HBM bytes);
  legs        N = 1 only: BASELINE configs[2] (SpGEMM R-MAT scale 22, edge factors 1 and 4), configs[3] (dense gram
              2M x 100k) and the configs[4] BSR-16 variant, each timed with its own roofline;
  configs4    N = 8 (or --c5): BASELINE configs[4], CSR(8M x 1M, 64 nnz/row) x dense(1M x 256) row-sharded with the
              all-gather, peer rows verified;
  cpu_baseline  the unmodified reference (sparse_dot_mkl on real oneMKL) on this host's cores.
"""
import argparse
import ctypes
import hashlib
import importlib.util
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import _cases as cs  # noqa: E402  (the one generator of the BASELINE recipes, shared with the tests)

M_ROWS = 1_000_000
K_COLS = 1_000_000
NNZ_PER_ROW = 50
N_DENSE = 128
BETA = 0.5
DTYPE = np.float32
T_START = time.perf_counter()


# ----------------------------------------------------------------------------- workload
def make_workload(rows, cols, per_row, n_dense, seed):
    """BASELINE C2 / C5 recipe: tests/_cases.c2_workload."""
    return cs.c2_workload(rows, cols, per_row, n_dense, seed)


def algorithmic_bytes(rows, nnz, n_dense, beta_nonzero=True, si=4, sv=4):
    """Gather-model bytes of one SpMM launch:
    (si+sv) + N*sv per stored entry, N*sv*(1+[beta!=0]) per output row, indptr once."""
    return nnz * (si + sv) + (rows + 1) * 8 + nnz * n_dense * sv + rows * n_dense * sv * (2 if beta_nonzero else 1)


def unique_bytes(rows, cols, nnz, n_dense, beta_nonzero=True, si=4, sv=4):
    return nnz * (si + sv) + (rows + 1) * 8 + cols * n_dense * sv + rows * n_dense * sv * (2 if beta_nonzero else 1)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the query runs every 20 ms from before
    the warm-up steps (nvidia-smi needs ~0.1 s to deliver its first line) and carries nvidia-smi's own timestamp;
    stop(t0, t1) keeps the samples taken between the two wall-clock marks of the timed region."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, t0=None, t1=None):
        """t0 / t1: time.time() just before / after the timed region."""
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                stamp = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((stamp, float(f[1]), float(f[2]), [n for n, flag in zip(names, f[4:8])
                                                               if flag.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t0 - 0.005 <= r[0] <= t1 + 0.005]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: the loaded stretch around it (warm-up + timed steps)
            inside = [r for r in rows if t0 is not None and t0 - 0.25 <= r[0] <= t1 + 0.05] or rows
            window = "warm-up + timed region (no sample fell inside the timed region itself)"
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "samples": len(sm), "window": window, "period_ms": 20, "reasons": reasons}


# ----------------------------------------------------------------------------- ncu traffic, tied to the SASS
_SASS_FUNCS = None  # [(mangled name, instruction text)] of every function of libsdb200's objects, read once


def _sass_functions():
    global _SASS_FUNCS
    if _SASS_FUNCS is not None:
        return _SASS_FUNCS or None
    obj_dir = os.path.join(ROOT, "sparse_dot_b200", "csrc", "_obj")
    funcs = []
    try:
        objs = sorted(f for f in os.listdir(obj_dir) if f.endswith(".o"))
        for o in objs:
            out = subprocess.run(["cuobjdump", "-sass", os.path.join(obj_dir, o)], capture_output=True, text=True,
                                 timeout=300).stdout
            name, text = None, []
            for line in out.splitlines():
                m = re.match(r"\s*Function : (\S+)", line)
                if m:
                    if name is not None:
                        funcs.append((name, "".join(text)))
                    name, text = m.group(1), []
                    continue
                if name is not None:
                    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
                    if m:
                        text.append(m.group(1))
            if name is not None:
                funcs.append((name, "".join(text)))
    except (OSError, subprocess.TimeoutExpired):
        funcs = []
    _SASS_FUNCS = funcs
    return funcs or None


def sass_sha(kernel_substr):
    """sha1 of the instruction text of every function of libsdb200's objects whose mangled name contains
    `kernel_substr` (addresses, comments and the names themselves left out; objects in name order, functions in file
    order), or None when cuobjdump / the objects are not there or nothing matches.  The ncu DRAM bytes in
    profiles/kernel_traffic.json are only quoted for the SASS they were captured from."""
    funcs = _sass_functions()
    if funcs is None:
        return None
    h, found = hashlib.sha1(), False
    for name, text in funcs:
        if kernel_substr in name:
            found = True
            h.update(text.encode())
    return h.hexdigest() if found else None


def mangled_hint(kernel_name):
    """'spmm_stream_kernel<float,6,32,2,2>' -> the Itanium-mangled fragment of that instantiation."""
    m = re.match(r"(\w+)<(.*)>", kernel_name)
    if not m:
        return kernel_name
    base, args = m.group(1), m.group(2).split(",")
    enc = {"float": "f", "double": "d"}
    parts = []
    for a in args:
        a = a.strip()
        parts.append(enc[a] if a in enc else f"Li{a}E")
    return f"{len(base)}{base}I" + "".join(parts) + "E"


def captured_traffic(kernel_name):
    """(entry, note): the ncu capture of `kernel_name` if profiles/kernel_traffic.json has one for this SASS."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        entry = json.load(open(path)).get("kernels", {}).get(kernel_name)
    except (OSError, ValueError):
        return None, "profiles/kernel_traffic.json missing"
    if not entry:
        return None, f"no ncu capture of {kernel_name}"
    want = entry.get("sass_sha1")
    have = sass_sha(entry.get("sass_match", mangled_hint(kernel_name)))
    if want and have and want == have:
        return entry, "ncu capture matches the SASS of this build"
    if have is None:
        return None, "cannot hash the SASS here (cuobjdump or objects missing): traffic withheld"
    return None, f"SASS differs from the profiled build ({have[:12]} vs {str(want)[:12]}): traffic withheld"


# ----------------------------------------------------------------------------- CPU arm
def cpu_spmm_runner(a, x, y0, threads=None):
    """The reference's CPU path for this workload, best available first: (1) the UNMODIFIED reference
    package's own dot_product_mkl on real oneMKL (oracle/ref_pkg.py), (2) real oneMKL through the
    reference's call sequence (oracle/mkl_ref.py), (3) the C restatement (oracle/sdb_oracle.c)."""
    from oracle import mkl_ref
    from oracle import oracle as orc

    ref_pkg, ref_status = try_reference_package()
    if ref_pkg is not None and not threads:
        from oracle import ref_pkg as _rp

        y = y0.copy()

        def step():
            ref_pkg.dot_product_mkl(a, x, out=y, out_scalar=BETA)
            return y

        return step, "reference", _rp.max_threads(), \
            "UNMODIFIED sparse_dot_mkl.dot_product_mkl(csr, ndarray, out=, out_scalar=) from baseline/_ref; " + ref_status
    if mkl_ref.available():
        if threads:
            mkl_ref.set_threads(threads)
        cores = threads or mkl_ref.max_threads()
        h = mkl_ref.Handle(a)
        y = y0.copy()

        def step():
            mkl_ref.spmm(a, x, beta=BETA, y=y, handle=h)
            return y

        kind, what = "reference", f"mkl_sparse_s_mm, {mkl_ref.version_string()} (embedded in libtorch_cpu), " \
                                  "reference call sequence _common.py:310-319 -> _sparse_dense.py:111-123"
        return step, kind, cores, what
    orc.build()
    cores = threads or orc.max_threads()
    orc.set_threads(cores)
    y = y0.copy()

    def step():
        orc.c_spmm(a, x, beta=BETA, y=y)
        return y

    return step, "port", cores, "oracle/sdb_oracle.c orc_spmm_f32 (OpenMP)"


def try_reference_package():
    """The UNMODIFIED reference (baseline/_ref) on real oneMKL; see oracle/ref_pkg.py."""
    from oracle import ref_pkg

    return ref_pkg.load()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all the host threads it can use: torch.distributed.run exports OMP_NUM_THREADS=1 to every rank it spawns,
    # which would silently make MKL single-threaded under the N > 1 launch (measured: 2167 ms instead of 206 ms)
    host_threads = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(host_threads)
    a, x, y0 = make_workload(M_ROWS, K_COLS, NNZ_PER_ROW, N_DENSE, seed=0)
    step, kind, cores, what = cpu_spmm_runner(a, x, y0)
    ref_status = what
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    g = algorithmic_bytes(M_ROWS, a.nnz, N_DENSE)
    value = g / dt / 1e9
    line = {
        "impl": "reference", "metric": "spmm_effective_hbm_gbs", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gflops": 2.0 * a.nnz * N_DENSE / dt / 1e9,
        "config": workload_config(1, "none"),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"full workload, every step ({what})"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_package": ref_status,
    }
    print(json.dumps(line))


def workload_config(n_gpus, allgather):
    return {
        "workload": f"CSR({M_ROWS}x{K_COLS}, {NNZ_PER_ROW} nnz/row, fp32) x dense({K_COLS}x{N_DENSE}) per GPU, "
                    f"out= accumulate (beta={BETA}); BASELINE.json configs[1]",
        "rows_per_gpu": M_ROWS, "cols": K_COLS, "nnz_per_row": NNZ_PER_ROW, "n_dense": N_DENSE, "beta": BETA,
        "parallelism": f"row-sharded x{n_gpus}, X replicated, allgather={allgather}",
        "cache": "inputs larger than L2 (A 0.4 GB + X 0.5 GB + Y 0.5 GB vs 126 MB L2); no flush needed",
    }


# ----------------------------------------------------------------------------- our arm
def pinned_like(arr, lib):
    """Copy `arr` into page-locked host memory (numpy view over sdb_host_alloc)."""
    p = ctypes.c_void_p()
    st = lib.sdb_host_alloc(ctypes.byref(p), arr.nbytes)
    if st != 0:
        raise RuntimeError("sdb_host_alloc failed")
    buf = (ctypes.c_char * arr.nbytes).from_address(p.value)
    out = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out


def expected_rows(a, x, y0, rows, steps, beta):
    """float64 recomputation of panel rows after `steps` passes of y <- A x + beta y:
    y_k = (1 - beta^k)/(1 - beta) * A x + beta^k * y0."""
    out = []
    for r in rows:
        s, e = a.indptr[r], a.indptr[r + 1]
        ax = (a.data[s:e].astype(np.float64)[:, None] * x[a.indices[s:e]].astype(np.float64)).sum(axis=0)
        out.append((1.0 - beta ** steps) / (1.0 - beta) * ax + beta ** steps * y0[r].astype(np.float64))
    return out


def panel_check(plan, a, x, y0, steps, beta, world, rank, dist, n_rows=8, n_peer_rows=4):
    """Parity of what was just timed.  Local: `n_rows` rows of this rank's block.  N > 1: every rank also
    publishes `n_peer_rows` recomputed rows of ITS block and every rank verifies the rows of ALL blocks in
    its own copy of the panel — a broken exchange cannot report ok."""
    def worst_of(first_row, want_rows, picks):
        worst = 0.0
        for r, want in zip(picks, want_rows):
            got = plan.read_rows(first_row + int(r), 1)[0].astype(np.float64)
            worst = max(worst, float(np.max(np.abs(got - want) / np.abs(want))))
        return worst

    rows = np.linspace(0, a.shape[0] - 1, n_rows).astype(np.int64)
    local = worst_of(plan.panel_row0, expected_rows(a, x, y0, rows, steps, beta), rows)
    res = {"rows_checked": int(n_rows), "max_rel_err": local, "ok": bool(local < 1e-5)}
    if world > 1 and plan.mode != "none":
        picks = np.linspace(0, a.shape[0] - 1, n_peer_rows + 2).astype(np.int64)[1:-1]
        mine = (plan.row0, [int(p) for p in picks], expected_rows(a, x, y0, picks, steps, beta))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        worst, checked = 0.0, 0
        for q, (row0, prows, want_rows) in enumerate(everyone):
            if q == rank:
                continue
            worst = max(worst, worst_of(row0, want_rows, prows))
            checked += len(prows)
        import torch

        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.update({"peer_rows_checked_per_rank": checked, "peer_blocks": world - 1,
                    "peer_max_rel_err_over_ranks": float(t.item()),
                    "ok": bool(res["ok"] and float(t.item()) < 1e-5)})
    return res


def bind_to_gpu_numa(gpu_index):
    """Pin this rank to the CPU cores local to its GPU (NVML) BEFORE any host buffer is allocated, so
    the page-locked arrays of the e2e leg are first-touched on the GPU's own NUMA node.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = set(os.sched_getaffinity(0))
        cpus &= allowed
        if cpus and cpus != allowed:  # a real restriction: this GPU has its own NUMA-local cores
            os.sched_setaffinity(0, cpus)
            return len(cpus)
        return 0  # NVML names every allowed core for this GPU (one NUMA node, e.g. a VM): nothing to bind to
    except Exception:
        return 0


def timed_steps(torch, stream, step, steps, sync_all):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for i in range(steps):
        step()
        ev[i + 1].record(stream)
    sync_all()
    return ev[0].elapsed_time(ev[-1]), [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]


def max_over_ranks(torch, dist, world, value):
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def e2e_leg(sdb, lib, torch, dist, world, a, x, y0, steps, sync_all, pinned):
    """dot_product_mkl(csr, ndarray, out=, out_scalar=) on host arrays, wall clock, max over ranks."""
    if pinned:
        xa, ya = pinned_like(x, lib), pinned_like(y0, lib)
        ap = sp.csr_matrix((pinned_like(a.data, lib), pinned_like(a.indices, lib), pinned_like(a.indptr, lib)),
                           shape=a.shape)
    else:
        xa, ya, ap = x, y0.copy(), a
    sdb.dot_product_mkl(ap, xa, out=ya, out_scalar=BETA)  # warm-up (allocator pools, staging rings, copy threads)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        sdb.dot_product_mkl(ap, xa, out=ya, out_scalar=BETA)
    torch.cuda.synchronize()
    dt = max_over_ranks(torch, dist, world, time.perf_counter() - t0) / steps
    phases = sdb.last_timing_ms()
    # what came back, against the closed form after steps + 1 calls (first and last row of the block)
    rows = np.array([0, a.shape[0] - 1])
    want = expected_rows(a, x, y0, rows, steps + 1, BETA)
    err = max(float(np.max(np.abs(ya[r].astype(np.float64) - w) / np.abs(w))) for r, w in zip(rows, want))
    return dt, phases, err


def run_ours(args):
    import torch
    import torch.distributed as dist

    import sparse_dot_b200 as sdb
    from sparse_dot_b200 import _lib
    from sparse_dot_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    local_cpus = bind_to_gpu_numa(local) if world > 1 else 0
    torch.cuda.set_device(local)
    lib = _lib.SDB.lib
    _lib.check(lib.sdb_set_device(local), "sdb_set_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = args.allgather if world > 1 else "none"

    rows, cols, n = args.rows, K_COLS, N_DENSE
    a, x, y0 = make_workload(rows, cols, NNZ_PER_ROW, n, seed=rank * 10)
    x = make_workload(1, cols, 1, n, seed=0)[1] if world > 1 else x  # X is replicated: same on every rank
    nnz = a.nnz

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- operands resident in HBM
    plan = sharded.RowShardedSpMM(a, n, world_size=world, rank=rank, allgather=mode,
                                  group=dist.group.WORLD if world > 1 else None)
    plan.set_x(x)
    plan.set_local_y(y0)
    stream = torch.cuda.Stream()  # a real (non-default) stream: kernels, events and NCCL all order on it

    def step():
        with torch.cuda.stream(stream):
            plan.run(alpha=1.0, beta=BETA, stream_ptr=stream.cuda_stream)

    # the first three steps, one by one: row-gather kernel, inspector + streaming kernel, streaming kernel
    _, first_ms = timed_steps(torch, stream, step, 3, sync_all)
    steps_done = 3
    inspector_ms = max(0.0, first_ms[1] - first_ms[2])
    tuned = None
    if world > 1 and mode == "fused" and not args.no_autotune:
        tuned = plan.autotune(beta=BETA, stream=stream)
        steps_done += tuned["steps_run"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # nvidia-smi's first sample
    for _ in range(max(args.warmup, 3)):
        step()
    steps_done += max(args.warmup, 3)
    sync_all()
    launches0 = sdb.kernel_launches()
    wall0 = time.time()
    total_ms, kernel_ms = timed_steps(torch, stream, step, args.steps, sync_all)
    wall1 = time.time()
    steps_done += args.steps
    launches = sdb.kernel_launches() - launches0
    kernel_name = sdb.last_spmm_kernel()  # what the timed steps launched (same thread)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_per_step = max_over_ranks(torch, dist, world, total_ms) / args.steps

    # parity of what was just timed: local rows and (N > 1) rows of every peer's block in this rank's panel
    check = panel_check(plan, a, x, y0, steps_done, BETA, world, rank, dist)

    # ---- end to end through the public API on host arrays, rank-local shard
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        dt, phases, err = e2e_leg(sdb, lib, torch, dist, world, a, x, y0, e2e_steps, sync_all, pinned=True)
        g1 = algorithmic_bytes(rows, nnz, n)
        e2e = {
            "value": g1 * world / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3, "steps": e2e_steps,
            "h2d_bytes_per_step": int(a.data.nbytes + a.indices.nbytes + a.indptr.nbytes + x.nbytes + y0.nbytes),
            "d2h_bytes_per_step": int(y0.nbytes),
            "host_memory": "pinned (sdb_host_alloc)" + (f", rank bound to {local_cpus} GPU-local cores" if local_cpus else
                                                        (", all ranks share one NUMA node" if world > 1 else "")),
            "device_spans_ms": {"start_to_last_upload": phases[0], "kernel_sum": phases[1], "whole_call": phases[2]},
            "max_rel_err": err,
            "api": "sparse_dot_b200.dot_product_mkl(csr, ndarray, out=, out_scalar=) -> sdb_spmm_csr_host "
                   "(3-stream row-chunk pipeline: upload / kernel / download overlap)",
        }
        dtp, phases_p, err_p = e2e_leg(sdb, lib, torch, dist, world, a, x, y0, e2e_steps, sync_all, pinned=False)
        e2e["pageable"] = {
            "value": g1 * world / dtp / 1e9, "unit": "GB/s", "ms_per_step": dtp * 1e3, "steps": e2e_steps,
            "host_memory": "ordinary (pageable) numpy / scipy arrays, staged through page-locked rings by the "
                           "library's copy threads", "vs_pinned": dtp / dt, "max_rel_err": err_p,
            "device_spans_ms": {"start_to_last_upload": phases_p[0], "kernel_sum": phases_p[1],
                                "whole_call": phases_p[2]},
        }

    # ---- BASELINE configs[4] across the ranks (N = 8 by default)
    configs4 = None
    if world > 1 and mode == "fused" and (args.c5 or world == 8):
        plan.close()
        plan = None
        configs4 = run_configs4(args, torch, dist, sharded, sdb, world, rank, sync_all)

    # ---- the sparse x sparse and gram products sharded over the same ranks (device-level exchange)
    sharded_products = None
    if world > 1 and not args.no_sharded_products:
        if plan is not None:
            plan.close()
            plan = None
        sharded_products = run_sharded_products(torch, dist, sharded, world, rank, sync_all)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    g = algorithmic_bytes(rows, nnz, n)
    u = unique_bytes(rows, cols, nnz, n)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    k_ms = float(np.mean(kernel_ms))
    effective = g / (k_ms * 1e-3) / 1e9
    captured, traffic_note = captured_traffic(kernel_name)
    traffic = captured.get("dram_bytes_per_launch") if captured else None
    # probes of this box, this run: HBM read, L2 -> SM read and the 512-byte-row gather out of L2
    probes = None
    if not args.no_probes:
        try:
            probes = {"hbm_read_gbs": _lib.probe_bandwidth(0, 4 << 30, 3),
                      "l2_read_gbs": _lib.probe_bandwidth(1, 48 << 20, 3),
                      "l2_gather512_gbs": _lib.probe_bandwidth(2, 48 << 20, 3),
                      "how": "sdb_probe_bandwidth (csrc/probe.cu): 16-byte loads, CUDA events; HBM over 4 GiB, L2 "
                             "over an L2-resident 48 MiB (64 passes per launch, L1 bypassed), gather = whole 512 B rows "
                             "at random positions of the 48 MiB"}
        except ValueError as e:
            probes = {"error": str(e)[:200]}
    l2_to_sm = None
    if captured and captured.get("l2_to_sm_bytes_per_launch") and probes and probes.get("l2_gather512_gbs"):
        l2_bytes = captured["l2_to_sm_bytes_per_launch"]
        ach_l2 = l2_bytes / (k_ms * 1e-3) / 1e9
        l2_peak = max(probes["l2_gather512_gbs"], probes["l2_read_gbs"])
        l2_to_sm = {"bytes_per_launch": l2_bytes, "achieved": ach_l2, "peak": l2_peak, "unit": "GB/s",
                    "frac": ach_l2 / l2_peak, "peak_source": "probed in this run (the faster of the two L2 probes)",
                    "note": "binding resource of the L2-tiled kernel: every stored entry moves one 512 B X row from "
                            "L2 into an SM (ncu l1tex__m_xbar2l1tex_read_bytes); half of those rows never reach HBM"}
    if traffic:
        achieved, basis = traffic / (k_ms * 1e-3) / 1e9, "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch"
    else:
        achieved, basis = effective, "algorithmic (gather-model) bytes: no ncu capture of this SASS (" + traffic_note + ")"
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "basis": basis, "traffic_note": traffic_note,
        "effective_achieved": effective, "effective_frac": effective / peak,
        "effective_frac_of_8TBs_nominal": effective / 8000.0,
        "l2_to_sm": l2_to_sm, "probes": probes, "peak_source": peak_src, "kernel": kernel_name, "kernel_ms": k_ms,
        "algorithmic_bytes": g, "unique_bytes": u,
        "model": "gather model: (4+4)+128*4 B per nnz, 128*4*2 B per row, 8 B per indptr entry",
    }
    inspector = {
        "ms": inspector_ms, "first_three_steps_ms": first_ms,
        "extra_hbm_bytes": int(nnz * 8) if "stream" in kernel_name else 0,
        "what": "one-time slab-ordered copy of A built on the handle's second multiplication "
                "(strict_rows_kernel + slab_permute_kernel), outside the timed region; step 1 runs the row-gather "
                "kernel, step 2 the inspector + streaming kernel, step 3 onwards the streaming kernel",
    }

    # ---- CPU baseline beside it (rank 0, N = 1 only): a bounded sample = 3 full-workload steps
    cpu = None
    if world == 1 and not args.no_cpu:
        cstep, kind, cores, what = cpu_spmm_runner(a, x, y0)
        cstep()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            cstep()
        cdt = (time.perf_counter() - t0) / reps
        cpu = {"value": g / cdt / 1e9, "unit": "GB/s", "cores": cores, "kind": kind, "ms_per_step": cdt * 1e3,
               "gflops": 2.0 * nnz * n / cdt / 1e9,
               "sample": f"{reps} full-workload steps after 1 warm-up ({what})"}

    # ---- the other BASELINE configs on this GPU (N = 1 only)
    legs = None
    if world == 1 and not args.no_legs:
        if plan is not None:
            plan.close()
            plan = None
        del a, x, y0
        legs = run_legs(args, peak)

    line = {
        "metric": "spmm_effective_hbm_gbs", "value": g * world / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gflops": 2.0 * nnz * n * world / (ms_per_step * 1e-3) / 1e9,
        "config": workload_config(world, mode),
        "roofline": roofline,
        "inspector": inspector,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_spot_check": check,
        "exchange": tuned,
        "configs4": configs4,
        "sharded_products": sharded_products,
        "legs": legs,
        "version": sdb.get_version_string(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_configs4(args, torch, dist, sharded, sdb, world, rank, sync_all):
    """BASELINE configs[4]: CSR(world x 1M rows x 1M, 64 nnz/row, fp32) x dense(1M x 256), one 1M-row shard per
    rank, the all-gather of the output panel fused into the step; every rank verifies rows of every block."""
    rows, cols, per_row, n = 1_000_000, 1_000_000, 64, 256
    a, _, y0 = make_workload(rows, cols, per_row, n, seed=5 + 10 * rank)
    x = np.random.default_rng(6).random((cols, n), dtype=np.float32)
    stream = torch.cuda.Stream()
    res = {"workload": f"CSR({world * rows}x{cols}, {per_row} nnz/row, fp32) x dense({cols}x{n}), row-sharded x{world}, "
                       "fused all-gather; BASELINE.json configs[4]", "beta": BETA}
    # the shard's kernel alone (no exchange)
    solo = sharded.RowShardedSpMM(a, n, world_size=world, rank=rank, allgather="none", group=dist.group.WORLD)
    solo.set_x(x)
    solo.set_local_y(y0)

    def solo_step():
        with torch.cuda.stream(stream):
            solo.run(alpha=1.0, beta=BETA, stream_ptr=stream.cuda_stream)

    timed_steps(torch, stream, solo_step, 3, sync_all)
    t_ms, _ = timed_steps(torch, stream, solo_step, 5, sync_all)
    res["kernel_only_ms"] = max_over_ranks(torch, dist, world, t_ms) / 5
    res["kernel"] = sdb.last_spmm_kernel()
    solo.close()

    plan = sharded.RowShardedSpMM(a, n, world_size=world, rank=rank, allgather="fused", group=dist.group.WORLD)
    plan.set_x(x)
    plan.set_local_y(y0)

    def step():
        with torch.cuda.stream(stream):
            plan.run(alpha=1.0, beta=BETA, stream_ptr=stream.cuda_stream)

    timed_steps(torch, stream, step, 3, sync_all)
    done = 3
    if not args.no_autotune:
        tuned = plan.autotune(beta=BETA, stream=stream)
        done += tuned["steps_run"]
        res["exchange"] = tuned
    steps = 10
    t_ms, _ = timed_steps(torch, stream, step, steps, sync_all)
    done += steps
    ms = max_over_ranks(torch, dist, world, t_ms) / steps
    ingress = (world - 1) * rows * n * 4
    g = algorithmic_bytes(rows, a.nnz, n)
    res.update({
        "ms_per_step": ms, "steps": steps, "value_gbs": g * world / (ms * 1e-3) / 1e9,
        "gflops": 2.0 * a.nnz * n * world / (ms * 1e-3) / 1e9,
        "ingress_bytes_per_rank": ingress, "ingress_gbs_achieved": ingress / (ms * 1e-3) / 1e9,
        "step_over_kernel": ms / res["kernel_only_ms"],
        "parity_spot_check": panel_check(plan, a, x, y0, done, BETA, world, rank, dist),
    })
    plan.close()
    return res


def run_sharded_products(torch, dist, sharded, world, rank, sync_all):
    """SURVEY.md §8e last row at this rank count: C = A @ B (R-MAT scale 20, edge factor 4, sorted output) with the
    rows of A split across the ranks and the row blocks exchanged device to device, and the dense gram of
    CSR(400k x 20k, 100 nnz/row) reduce-scattered onto panel owners.  Wall clock per call (max over ranks): uploads
    of the operands, products, exchange.  Checked with the checksum of checksums / sampled entries."""
    out = {}
    try:
        a = cs.rmat_csr(20, 4, np.float32, seed=1)
        b = cs.rmat_csr(20, 4, np.float32, seed=2)
        with sharded.spgemm_sharded_device(a, b, world, rank, reorder_output=True):
            pass  # warm-up: memory pools, NCCL channels
        sync_all()
        t0 = time.perf_counter()
        c = sharded.spgemm_sharded_device(a, b, world, rank, reorder_output=True)
        dt = max_over_ranks(torch, dist, world, time.perf_counter() - t0)
        with c:
            total = float(c.values.sum(dtype=torch.float64).item())
            colsum_a = np.bincount(a.indices, weights=a.data.astype(np.float64), minlength=a.shape[1])
            rowsum_b = np.asarray(b.astype(np.float64).sum(axis=1)).ravel()
            want = float(np.dot(colsum_a, rowsum_b))
            srt = bool((torch.diff(c.indices.to(torch.int64)) > 0).sum().item() >= c.nnz - c.shape[0])
            out["spgemm"] = {"workload": "R-MAT scale 20, edge factor 4, fp32, sorted; rows of A sharded, B replicated",
                             "ms": dt * 1e3, "nnz_c": int(c.nnz), "total_rel_err": abs(total - want) / want,
                             "rows_sorted": srt, "exchange": "NCCL broadcasts of the row blocks out of HBM"}
    except Exception as e:
        out["spgemm"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    try:
        m = cs.uniform_rows_csr(400_000, 20_000, 100, np.float32, seed=4)
        sharded.gram_dense_sharded(m, world, rank, gather=False)
        sync_all()
        t0 = time.perf_counter()
        row0, panel = sharded.gram_dense_sharded(m, world, rank, gather=False)
        dt = max_over_ranks(torch, dist, world, time.perf_counter() - t0)
        # this rank's panel against a host recomputation of two of its rows
        worst = 0.0
        for r in (row0, row0 + panel.shape[0] - 1):
            hits = np.flatnonzero(m.indices == r)
            src = np.searchsorted(m.indptr, hits, side="right") - 1
            want = np.asarray(m[src].astype(np.float64).T @ m.data[hits].astype(np.float64)).ravel()
            got = panel[r - row0].astype(np.float64)
            worst = max(worst, float(np.abs(got[r:] - want[r:]).max() / max(1e-30, np.abs(want[r:]).max())))
            worst = max(worst, float(np.abs(got[:r]).max()) if r else 0.0)  # strict lower triangle is zero
        out["gram"] = {"workload": "A^T A of CSR(400k x 20k, 100 nnz/row, fp32), dense upper; rows of A sharded",
                       "ms": dt * 1e3, "panel_rows_of_this_rank": int(panel.shape[0]),
                       "max_rel_err_over_ranks": max_over_ranks(torch, dist, world, worst),
                       "exchange": "NCCL reduce of equal-area row panels onto their owners (reduce-scatter)"}
    except Exception as e:
        out["gram"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def load_run_configs():
    spec = importlib.util.spec_from_file_location("_run_configs", os.path.join(ROOT, "scripts", "run_configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_legs(args, peak):
    """The other BASELINE configs on one GPU, each with its own roofline (algorithmic bytes per SURVEY.md §8d over
    the measured time; `traffic` from profiles/kernel_traffic.json when the SASS matches).  Checks are the
    size-independent properties of scripts/run_configs.py (checksums, sampled rows, triangle rules)."""
    rc = load_run_configs()
    legs, budget_s = {}, args.legs_budget

    def roof(name, gbytes, ms, kernel_key):
        entry, note = captured_traffic(kernel_key)
        traffic = entry.get("dram_bytes_per_launch") if entry else None
        ach = gbytes / (ms * 1e-3)
        r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "basis": "algorithmic bytes", "traffic": traffic, "traffic_note": note}
        if traffic:
            r["traffic_gbs"] = traffic / (ms * 1e-3) / 1e9
            r["traffic_frac"] = r["traffic_gbs"] / peak
        return r

    wanted = [w.strip() for w in args.legs.split(",") if w.strip()]
    for name in wanted:
        if time.perf_counter() - T_START > budget_s:
            legs[name] = {"skipped": f"time budget of {budget_s} s for the whole bench reached"}
            continue
        t0 = time.perf_counter()
        try:
            if name in ("spgemm_ef1", "spgemm_ef4"):
                ef = 1 if name.endswith("1") else 4
                r = rc.run_c3_resident(22, ef)
                ms = r["spgemm_ordered_ms"]
                # numeric (4+4) + symbolic 4 bytes per product, A once per pass, C written once
                gb = (r["products"] * 12 + 2 * r["nnz_a"] * 8 + r["nnz_c"] * 8) / 1e9
                r["g_products_per_s_ordered"] = r["products"] / (ms * 1e-3) / 1e9
                # the ncu capture (sum over the kernels of one ordered call) exists for edge factor 1 only
                r["roofline"] = roof(name, gb, ms, "spgemm_ordered" if ef == 1 else "spgemm_ordered_ef4")
                r["config"] = f"BASELINE configs[2]: R-MAT scale 22, edge factor {ef}, fp32, sorted sparse output " \
                              "(sdb_spgemm_ordered = reorder_output=True), result kept in HBM"
            elif name == "gram":
                r = rc.run_c4(2_000_000, 100_000, 100)
                ms = r["syrkd_ms"]
                gb = (r["nnz"] * 8 + r["n"] * (r["n"] + 1) / 2 * 4) / 1e9
                r["roofline"] = roof(name, gb, ms, "spgemm_dense_red_kernel<float>")
                r["config"] = "BASELINE configs[3]: A^T A of CSR(2M x 100k, 100 nnz/row, fp32), dense upper triangle, " \
                              "device resident"
            elif name == "bsr16":
                r = rc.run_c5bsr(62_500, 4, 16, 256)
                ms = r["spmm_ms"]
                gb = (r["nnz"] * 4 + r["nnz"] / 256 * (4 + 16 * 256 * 4) + r["rows"] * 256 * 4) / 1e9
                r["roofline"] = roof(name, gb, ms, r.get("kernel", "spmm_bsr_kernel"))
                r["config"] = "BASELINE configs[4] BSR variant: 1M x 1M, 62 500 block rows x 4 blocks of 16x16 fp32, " \
                              "x dense(1M x 256)"
            elif name == "spmv":
                r = rc.run_spmv(1_000_000, 1_000_000, 50)
                r["roofline"] = roof(name, r["algorithmic_bytes"] / 1e9, r["spmv_ms"], r.get("kernel", "spmv_wide_kernel<float,16,0>"))
                r["config"] = "SURVEY 8f rank 2 (mkl_sparse_?_mv): the configs[1] matrix x one dense column, fp32, " \
                              "beta = 0.5, operands in HBM"
            else:
                r = {"skipped": "unknown leg"}
        except Exception as e:  # a leg must never take the headline down with it
            r = {"error": f"{type(e).__name__}: {e}"[:300]}
        r["wall_s"] = time.perf_counter() - t0
        legs[name] = r
    return legs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--allgather", choices=["fused", "nccl", "none"], default="fused")
    ap.add_argument("--rows", type=int, default=M_ROWS, help="rows per GPU (default: the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-probes", action="store_true")
    ap.add_argument("--no-autotune", action="store_true", help="N > 1: keep the library's default exchange strategy")
    ap.add_argument("--no-legs", action="store_true", help="N = 1: skip the configs[2] / [3] / BSR legs")
    ap.add_argument("--legs", default="bsr16,spmv,spgemm_ef1,gram,spgemm_ef4")
    ap.add_argument("--legs-budget", type=float, default=400.0, help="seconds after which remaining legs are skipped")
    ap.add_argument("--c5", action="store_true", help="N > 1: also run BASELINE configs[4] (default at N = 8)")
    ap.add_argument("--no-sharded-products", action="store_true", help="N > 1: skip the sharded SpGEMM / gram record")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""
ncu_traffic.py — keeps profiles/kernel_traffic.json, the table bench.py reads `roofline.traffic` from.

    python scripts/ncu_traffic.py --hashes                       (on the GPU box: SASS sha1 per profiled kernel family)
    python scripts/ncu_traffic.py --dump <rep>                   (on the GPU box: per-launch numbers as JSON)
    python scripts/ncu_traffic.py --update <tag> <rep|json> ..   (here: fold captures / dumps + that run's hashes in)

An entry is only quoted by bench.py while the SASS of the kernel is the one that was profiled: the hash written
by `--hashes` in the same gpurun call travels back in gpurun_out/<tag>_sass_hashes.json and is stored with the bytes.
Entries are keyed by the kernel's demangled name without spaces ("spmm_stream_kernel<float,6,32,2,2>"); a family
entry (key = what `--family` names, e.g. "spgemm_ordered") sums every launch of the capture whose name matches.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TABLE = os.path.join(ROOT, "profiles", "kernel_traffic.json")

# kernel families whose SASS is hashed on the box: substring of the mangled name
FAMILIES = {
    "spmm_stream_kernel": "spmm_stream_kernel",
    "spmm_stream_half_kernel": "spmm_stream_half_kernel",
    "spmm_rowmajor_kernel": "spmm_rowmajor_kernel",
    "spmm_bsr_kernel": "spmm_bsr_kernel",
    "spmv_wide_kernel": "spmv_wide_kernel",
    "spmv_tile_kernel": "spmv_tile_kernel",
    "spmm_bsr_mma_kernel": "spmm_bsr_mma_kernel",
    "spgemm_dense_red_kernel": "spgemm_dense_red_kernel",
    "spgemm_": "spgemm_",
}
# per-instantiation matches for the kernels bench.py quotes (mangled-name fragments)
MATCH = {
    "spmm_stream_kernel<float,6,32,2,2>": "spmm_stream_kernelIfLi6ELi32ELi2ELi2ELb0",
    "spmm_bsr_kernel<float,16,256,0,2>": "spmm_bsr_kernelIfLi16ELi256ELb0ELi2E",
    "spgemm_dense_red_kernel<float>": "spgemm_dense_red_kernelIf",
    "spmm_rowmajor_kernel<float,4,32,2,8>": "spmm_rowmajor_kernelIfLi4ELi32ELi2ELi8ELb0E",
    "spmm_bsr_mma_kernel<float,16,256,0,2>": "spmm_bsr_mma_kernelIfLi16ELi256ELb0ELi2E",
    "spmm_stream_half_kernel<float,6,16,4>": "spmm_stream_half_kernelIfLi6ELi16ELi4E",
    "spmv_wide_kernel<float,16,0>": "spmv_wide_kernelIfLi16ELb0E",
    "spmv_tile_kernel<float>": "spmv_tile_kernelIfE",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TSCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6,
          "nsecond": 1e-9}


def short_name(kernel_name):
    m = re.search(r"(\w+)\s*<([^()]*)>\s*\(", kernel_name)
    if m:
        return m.group(1) + "<" + m.group(2).replace(" ", "") + ">"
    m = re.search(r"(\w+)\s*\(", kernel_name)
    return m.group(1) if m else kernel_name


def family_of(short):
    base = short.split("<")[0]
    return base if base in FAMILIES else ("spgemm_" if base.startswith("spgemm_") else None)


def hashes():
    import bench

    print(json.dumps({fam: bench.sass_sha(sub) for fam, sub in FAMILIES.items()}, indent=1))


def read_rep(path):
    if path.endswith(".json"):  # a dump made on the GPU box (--dump): the reports themselves are too big to bring back
        return json.load(open(path))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def col(r, key, table):
        if key not in hdr:
            return None
        i = hdr.index(key)
        try:
            return float(r[i].replace(",", "")) * table.get(units[i], 1.0)
        except ValueError:
            return None

    out = []
    for r in rows[2:]:
        out.append({
            "name": short_name(r[hdr.index("Kernel Name")]),
            "dram_read": col(r, "dram__bytes_read.sum", SCALE),
            "dram_write": col(r, "dram__bytes_write.sum", SCALE),
            "l2_to_sm": col(r, "l1tex__m_xbar2l1tex_read_bytes.sum", SCALE),
            "seconds": col(r, "gpu__time_duration.sum", TSCALE),
            "l2_hit_pct": col(r, "lts__t_sector_hit_rate.pct", {}),
            "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active", {}),
        })
    return out


def update(tag, reps, family_key=None):
    table = {"note": "", "kernels": {}}
    if os.path.exists(TABLE):
        table = json.load(open(TABLE))
    table["note"] = ("DRAM bytes per launch from ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum), "
                     "each stored with the sha1 of the SASS it was captured from (scripts/ncu_traffic.py); bench.py quotes "
                     "an entry only while the SASS of the build matches")
    import bench

    # The hash is taken from the objects HERE: run --update right after the gpurun call, before touching a kernel
    # (the box builds nothing: it runs the objects this tree shipped).  The family hashes written on the box by
    # --hashes (gpurun_out/<tag>_sass_hashes.json) are kept beside them as a cross-check.
    hpath = os.path.join(ROOT, "gpurun_out", f"{tag}_sass_hashes.json")
    box = json.load(open(hpath)) if os.path.exists(hpath) else {}
    here = {fam: bench.sass_sha(sub) for fam, sub in FAMILIES.items()}
    for fam, sha in box.items():
        if sha and here.get(fam) and sha != here[fam]:
            raise SystemExit(f"{fam}: the objects here differ from the ones the box profiled ({here[fam]} vs {sha})")
    for rep in reps:
        launches = read_rep(rep)
        if family_key:
            tot = {"dram_bytes_read_per_launch": 0.0, "dram_bytes_write_per_launch": 0.0, "seconds": 0.0, "launches": 0}
            for k in launches:
                tot["dram_bytes_read_per_launch"] += k["dram_read"] or 0.0
                tot["dram_bytes_write_per_launch"] += k["dram_write"] or 0.0
                tot["seconds"] += k["seconds"] or 0.0
                tot["launches"] += 1
            fam = "spgemm_" if family_key.startswith("spgemm") else family_key
            table["kernels"][family_key] = {
                "source": f"{os.path.basename(rep)} (sum over the {tot['launches']} launches of the capture)",
                "dram_bytes_read_per_launch": tot["dram_bytes_read_per_launch"],
                "dram_bytes_write_per_launch": tot["dram_bytes_write_per_launch"],
                "dram_bytes_per_launch": tot["dram_bytes_read_per_launch"] + tot["dram_bytes_write_per_launch"],
                "duration_ms_under_ncu": tot["seconds"] * 1e3,
                "sass_match": FAMILIES.get(fam, fam), "sass_sha1": bench.sass_sha(FAMILIES.get(fam, fam)),
            }
            continue
        for k in launches:
            fam = family_of(k["name"])
            table["kernels"][k["name"]] = {
                "source": os.path.basename(rep),
                "dram_bytes_read_per_launch": k["dram_read"], "dram_bytes_write_per_launch": k["dram_write"],
                "dram_bytes_per_launch": (k["dram_read"] or 0.0) + (k["dram_write"] or 0.0),
                "l2_to_sm_bytes_per_launch": k["l2_to_sm"],
                "duration_ms_under_ncu": (k["seconds"] or 0.0) * 1e3, "l2_hit_rate_pct": k["l2_hit_pct"],
                "warps_active_pct": k["warps_active_pct"],
                "sass_match": MATCH.get(k["name"], FAMILIES.get(fam, k["name"].split("<")[0])),
                "sass_sha1": bench.sass_sha(MATCH.get(k["name"], FAMILIES.get(fam, k["name"].split("<")[0]))),
            }
    json.dump(table, open(TABLE, "w"), indent=1)
    print(json.dumps({k: (v["dram_bytes_per_launch"], v.get("sass_sha1")) for k, v in table["kernels"].items()}, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--hashes", action="store_true")
    ap.add_argument("--dump", metavar="REP", help="per-launch numbers of an .ncu-rep as JSON on stdout")
    ap.add_argument("--update", nargs="+", metavar=("TAG", "REP"))
    ap.add_argument("--family", default=None, help="store the capture as ONE entry under this key (sum of its launches)")
    a = ap.parse_args()
    if a.dump:
        print(json.dumps(read_rep(a.dump)))
    elif a.hashes:
        hashes()
    elif a.update:
        update(a.update[0], a.update[1:], a.family)

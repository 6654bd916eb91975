#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one block of key metrics per profiled launch.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x_summary.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"== {name[:110]}")
        rd = wr = dur = None
        for k in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                print(f"   {k:75s} {v} {u}")
                if k == "dram__bytes_read.sum":
                    rd = (float(v.replace(",", "")), u)
                if k == "dram__bytes_write.sum":
                    wr = (float(v.replace(",", "")), u)
                if k == "gpu__time_duration.sum":
                    dur = (float(v.replace(",", "")), u)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        if rd and wr and dur:
            b = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
            t = dur[0] * tscale[dur[1]]
            print(f"   -> DRAM traffic {b / 1e9:.3f} GB in {t * 1e3:.3f} ms = {b / t / 1e9:.0f} GB/s")


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text table for profiles/."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_thr_%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__cycles_active.avg", "smsp_cycles"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none summary of", path)
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print("kernel:", name[:110])
        for k, short in KEYS:
            if k in idx:
                print("   {:<16} {:>16} {}".format(short, r[idx[k]], units[idx[k]]))


if __name__ == "__main__":
    main(sys.argv[1])

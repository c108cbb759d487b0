#!/usr/bin/env python
"""One CSV row per profiled launch with the metrics the rooflines are argued from (read here, no GPU needed):
    python profiles/ncu_table.py gpurun_out/x.ncu-rep > profiles/x_ncu_table.csv"""
import csv
import subprocess
import sys

COLS = [
    ("ID", "id"), ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pct"), ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    have = [(k, n) for k, n in COLS if k in hdr]
    w.writerow([f"{n} [{units[hdr.index(k)]}]" if units[hdr.index(k)] else n for k, n in have])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        w.writerow([d[k][:90] if n == "kernel" else d[k] for k, n in have])


if __name__ == "__main__":
    main(sys.argv[1])

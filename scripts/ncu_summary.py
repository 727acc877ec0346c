#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small CSV of the
metrics the roofline numbers in bench.py / DESIGN.md are taken from.

    python scripts/ncu_summary.py gpurun_out/prof_chain.ncu-rep profiles/r01_chain_fused.csv
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["launch", "kernel", "metric", "value", "unit"])
        for li, r in enumerate(data):
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    w.writerow([li, r[name_i], k, r[i], units[i]])
    print("wrote", out, "(%d launches)" % len(data))


if __name__ == "__main__":
    main()

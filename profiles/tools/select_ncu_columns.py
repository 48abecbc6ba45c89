#!/usr/bin/env python
"""Reduce the raw page of an `ncu --set full` capture to the columns the tables under profiles/ use.

    ncu -i gpurun_out/<tag>_prof.ncu-rep --page raw --csv > /tmp/raw.csv        (build container, no GPU needed)
    python profiles/tools/select_ncu_columns.py /tmp/raw.csv > profiles/<round>_ncu_full_volume_2048.csv
"""
import csv
import sys

COLS = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__cycles_active.avg']


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
    hdr = rows[0]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    w = csv.writer(sys.stdout)
    for r in rows:
        if len(r) >= len(hdr):
            w.writerow([r[i] for i in idx])


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Key metrics of the first kernel in an ncu report (raw page)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'launch__shared_mem_per_block_dynamic',
        'smsp__cycles_active.avg', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_active.avg']


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:3 if len(sys.argv) < 3 else None]:
        for i, h in enumerate(hdr):
            if h in KEYS or (len(sys.argv) > 2 and sys.argv[2] in h):
                print('%-70s %-12s %s' % (h, units[i], r[i]))


if __name__ == '__main__':
    main()

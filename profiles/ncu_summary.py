#!/usr/bin/env python
"""Prints the key metrics of an .ncu-rep (one block per captured launch).  Usage: ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__inst_executed_op_shared_ld.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stalls = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
    for row in data:
        print(row[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                print(f'   {w:75s} {row[hdr.index(w)]} {units[hdr.index(w)]}')
        top = sorted(((float(row[hdr.index(s)] or 0), s) for s in stalls), reverse=True)[:6]
        for v, s in top:
            print(f'   stall {s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):30s} {v:.2f}')


if __name__ == '__main__':
    main(sys.argv[1])

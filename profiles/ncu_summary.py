#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__inst_executed_pipe_', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__cycles_active.avg', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'smsp__average_warps_issue_stalled',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_active.avg']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('== kernel:', d.get('Kernel Name'), ' id', d.get('ID'))
        for h, u in zip(hdr, units):
            if any(h == w or h.startswith(w) for w in WANT):
                print('  %-90s %-12s %s' % (h, u, d[h]))


if __name__ == '__main__':
    main(sys.argv[1])

#!/usr/bin/env python
"""Summarise an .ncu-rep (needs `ncu` on PATH; works without a GPU): key metrics of every captured launch."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__maximum_warps_per_active_cycle_pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('== %s :: %s' % (rep, name[:80]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-72s %s %s' % (w, r[i], units[i]))
        # stall reasons, largest first
        st = [(float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp_latency_issue_stalled') or h.startswith('smsp__average_warps_issue_stalled')]
        for v, h in sorted(st, reverse=True)[:6]:
            print('  stall %-66s %.3f' % (h.replace('smsp__average_', ''), v))

#!/usr/bin/env python3
"""Print the headline metrics of an .ncu-rep (one kernel) - usage: ncu_summary.py <report>"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d['Kernel Name'][:60])
    for k in keys:
        if k in d:
            print('  %-62s %s %s' % (k, d[k], units[hdr.index(k)]))
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
            v = float(d[k] or 0)
            if v > 0.15:
                print('  %-62s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''), v))

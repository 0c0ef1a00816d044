#!/usr/bin/env python
"""Print selected metrics from an .ncu-rep (raw page) per kernel.  usage: ncu_metrics.py REPORT [regex ...]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pats = [re.compile(p) for p in sys.argv[2:]] or [re.compile(p) for p in (
    r'gpu__time_duration.sum', r'dram__bytes_(read|write).sum$', r'smsp__inst_executed.sum$', r'smsp__issue_active.avg.pct',
    r'wavefronts_mem_shared', r'bank_conflicts_pipe_lsu_mem_shared', r'sm__warps_active.avg.pct', r'registers_per_thread',
    r'inst_executed_pipe_(lsu|alu|fma|fmaheavy|fmalite|uniform|xu|adu)\.sum$', r'pipe_(alu|fma|fmaheavy)_cycles_active.avg.pct',
    r'inst_executed_op_(shared|global|local)_(ld|st)\.sum', r'l1tex__lsu_writeback', r'lts__t_sector_hit_rate.pct',
    r'smsp__average_warps_issue_stalled_.*_per_issue_active', r'l1tex__data_pipe_lsu_wavefronts.sum$')]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
ki = h.index('Kernel Name')
for row in r[2:]:
    print('==', row[0], row[ki][:90])
for i, name in enumerate(h):
    if any(p.search(name) for p in pats):
        print('%-95s %-10s %s' % (name, r[1][i], '  '.join(row[i] for row in r[2:])))

#!/usr/bin/env python
"""Opcode mix weighted by executed count from `ncu --page source --csv --print-source sass` output.
usage: ncu_opmix.py REPORT.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; mix = None; tot = 0; samples = None
def flush():
    if kern and mix and want in kern:
        print('==', kern[:100], 'warp-instr', tot)
        for op, n in mix.most_common(28):
            print('   %-22s %12d  %5.1f%%   stall-samples %6d' % (op, n, 100.0 * n / tot, samples[op]))
for r in rows:
    if r and r[0] == 'Kernel Name':
        flush(); kern = r[1]; mix = collections.Counter(); samples = collections.Counter(); tot = 0; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = r; ie = hdr.index('Instructions Executed'); si = hdr.index('# Samples'); continue
    if hdr and len(r) > ie:
        toks = r[1].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = '.'.join(op.split('.')[:3])
        try: n = int(r[ie])
        except ValueError: continue
        mix[op] += n; tot += n; samples[op] += int(r[si] or 0)
flush()

#!/usr/bin/env python
"""Top stalled SASS lines of one kernel in an .ncu-rep.  usage: ncu_hot.py REPORT kernel-substring [N]"""
import csv, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; lines = []; done = False
for r in rows:
    if r and r[0] == 'Kernel Name':
        if kern and want in kern and lines: break
        kern = r[1]; lines = []; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = r; si = hdr.index('# Samples'); ie = hdr.index('Instructions Executed'); continue
    if hdr and kern and want in kern and len(r) > ie:
        try: lines.append((int(r[si] or 0), int(r[ie] or 0), r[1].strip(), len(lines)))
        except ValueError: pass
tot = sum(l[0] for l in lines)
print(kern[:100], 'total samples', tot, 'sass lines', len(lines))
# stall reason columns
names = [h for h in hdr if h.startswith('stall_')] if hdr else []
for s, n, txt, idx in sorted(lines, reverse=True)[:N]:
    print('%5d  %5.2f%%  exec %9d  #%4d  %s' % (s, 100.0 * s / max(tot, 1), n, idx, txt))

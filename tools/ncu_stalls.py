#!/usr/bin/env python
"""Stall-reason samples per opcode for one kernel of an .ncu-rep.  usage: ncu_stalls.py REPORT kernel-substring"""
import csv, subprocess, sys, collections
rep, want = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; data = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        if kern and want in kern and data: break
        kern = r[1]; data = []; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr and kern and want in kern and len(r) == len(hdr): data.append(r)
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx = {h: hdr.index(h) for h in reasons}
ie = hdr.index('Instructions Executed')
by = collections.defaultdict(lambda: collections.Counter())
execs = collections.Counter()
tot = collections.Counter()
for r in data:
    t = r[1].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    execs[op] += int(r[ie] or 0)
    for h in reasons:
        v = int(r[idx[h]] or 0)
        by[op][h] += v; tot[h] += v
allsum = sum(tot.values())
print(kern[:90]); print('all samples', allsum)
print('reason totals:', ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / allsum) for h, v in tot.most_common(9)))
for op, c in sorted(by.items(), key=lambda kv: -sum(kv[1].values()))[:16]:
    s = sum(c.values())
    print('%-8s exec %10d samples %6d (%4.1f%%) : %s' % (op, execs[op], s, 100.0 * s / allsum, ', '.join('%s %d' % (h[6:], v) for h, v in c.most_common(5))))

#!/usr/bin/env python
"""Decode the scheduling control bits (stall count, yield, barriers) of a kernel's SASS.
usage: sass_ctrl.py OBJECT mangled-function-name [first last]   (line range of the listing, optional)"""
import re, subprocess, sys, collections
obj, fn = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump', '-sass', '-fun', fn, obj], capture_output=True, text=True).stdout
ins = []
lines = out.splitlines()
i = 0
pat = re.compile(r'^\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/')
pat2 = re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            lo, hi = int(m.group(3), 16), int(m2.group(1), 16)
            w = (hi << 64) | lo
            stall = (w >> 105) & 0xF
            yld = (w >> 109) & 1
            wbar = (w >> 110) & 7
            rbar = (w >> 113) & 7
            wait = (w >> 116) & 0x3F
            ins.append((m.group(2).strip(), stall, yld, wbar, rbar, wait))
            i += 2
            continue
    i += 1
a = int(sys.argv[3]) if len(sys.argv) > 3 else 0
b = int(sys.argv[4]) if len(sys.argv) > 4 else len(ins)
sel = ins[a:b]
tot = sum(s[1] for s in sel)
print('instructions', len(sel), 'sum of stall counts', tot, 'avg %.2f' % (tot / max(1, len(sel))))
by = collections.defaultdict(lambda: [0, 0])
for t, st, *_ in sel:
    op = t.split()[1] if t.startswith('@') else t.split()[0]
    op = '.'.join(op.split('.')[:2])
    by[op][0] += 1; by[op][1] += st
for op, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
    print('  %-18s n=%4d  stall-sum=%5d  avg=%.2f' % (op, n, s, s / n))
if '-l' in sys.argv:
    for k, (t, st, y, wb, rb, wt) in enumerate(sel):
        print('%4d  st=%2d y=%d wb=%d rb=%d wait=%02x  %s' % (a + k, st, y, wb, rb, wt, t))

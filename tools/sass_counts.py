#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the instructions that identify TMA (UTMALDG), mbarriers (SYNCS),
packed integer dot products (IDP.2A / IDP.4A), shared-memory atomics (ATOMS), byte permutes, saturating packs, and the
absence of tensor-core instructions, from `cuobjdump -sass` of the shipped libvtseg.so.
usage: python tools/sass_counts.py [video_transformer_b200/libvtseg.so] > profiles/rNN_sass_counts.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "video_transformer_b200/libvtseg.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTMALDG", "SYNCS", "IDP.2A", "IDP.4A", "ATOMS", "RED", "ATOMG", "PRMT", "I2IP", "VABSDIFF4", "SHF", "IMAD", "LDG",
        "STG", "LDS", "STS", "UTCMMA", "HMMA", "IMMA", "UTCBAR"]
kern = None
counts = collections.OrderedDict()
total = collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("vt::", "").replace("(anonymous namespace)::", "")
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for w in want:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
                total[w] += 1
print("# cuobjdump -sass %s : instruction counts per kernel (static)" % so)
print("# %-70s %7s %s" % ("kernel", "total", " ".join("%s" % w for w in want)))
for k, c in counts.items():
    if c["total"] == 0:
        continue
    print("%-72s %7d %s" % (k[:72], c["total"], " ".join("%*d" % (len(w), c[w]) for w in want)))
print("# whole library: " + ", ".join("%s %d" % (w, total[w]) for w in want))
print("# tensor-core opcodes (UTCMMA/HMMA/IMMA): %d -- nothing on this path is a contraction" %
      (total["UTCMMA"] + total["HMMA"] + total["IMMA"]))

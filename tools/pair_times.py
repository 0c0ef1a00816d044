#!/usr/bin/env python
"""Per-warp start/end times of the pair scaler (VT_PAIR_DEBUG_TIMES=file): how evenly do the warps of an SM finish?"""
import sys, numpy as np
a = np.fromfile(sys.argv[1], np.uint64).reshape(-1, 2)
n = 740 * 4
for name, t in (("luma", a[:n]), ("chroma", a[n:2 * n])):
    if len(t) < n: break
    t0 = t[:, 0].min()
    s, e = (t[:, 0] - t0).astype(np.float64) / 1e3, (t[:, 1] - t0).astype(np.float64) / 1e3
    busy = e > s
    print(name, "warps", n, "busy", int(busy.sum()), "kernel us %.1f" % e.max())
    q = np.percentile(e[busy], [0, 5, 25, 50, 75, 95, 100])
    print("  end-time percentiles (us):", np.round(q, 1))
    print("  mean residency: %.3f of the kernel" % ((e[busy] - s[busy]).sum() / (busy.sum() * e.max())))
    # by position of the warp's block in launch order (block id = gw // 4)
    blk = np.arange(n) // 4
    for lo in range(0, 740, 148):
        m = busy & (blk >= lo) & (blk < lo + 148)
        print("  blocks %3d-%3d: mean end %.1f us" % (lo, lo + 147, e[m].mean()))

#!/usr/bin/env python
"""Host<->device copy ceilings on the GPU box (pinned memory, 44 MB chunks like the ingest pipeline)."""
import time, torch
dev = torch.device("cuda:0")
n = 44 * 1024 * 1024
for chunks, label in ((1, "1 buffer"), (2, "2 buffers alternating"), (4, "4 buffers")):
    d = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(chunks)]
    h = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(chunks)]
    s = torch.cuda.Stream()
    for direction in ("d2h", "h2d"):
        with torch.cuda.stream(s):
            for i in range(8):
                (h[i % chunks].copy_(d[i % chunks], non_blocking=True) if direction == "d2h" else d[i % chunks].copy_(h[i % chunks], non_blocking=True))
            s.synchronize()
            t0 = time.perf_counter()
            reps = 64
            for i in range(reps):
                (h[i % chunks].copy_(d[i % chunks], non_blocking=True) if direction == "d2h" else d[i % chunks].copy_(h[i % chunks], non_blocking=True))
            s.synchronize()
            dt = time.perf_counter() - t0
        print("%-22s %s %.1f GB/s" % (label, direction, reps * n / dt / 1e9))
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
h = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(64):
    with torch.cuda.stream(s1): h[0].copy_(d[0], non_blocking=True)
    with torch.cuda.stream(s2): d[1].copy_(h[1], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("bidirectional: d2h %.1f GB/s + h2d %.1f GB/s" % (64 * n / dt / 1e9, 64 * n / dt / 1e9))

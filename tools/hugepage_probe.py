#!/usr/bin/env python
"""Does the D2H landing buffer's page size matter when several GPUs copy at once?  Compares, per rank and concurrently,
copies into (a) torch pinned memory (cudaHostAlloc) and (b) an anonymous mapping advised to transparent huge pages and
registered with cudaHostRegister.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/hugepage_probe.py
"""
import ctypes, mmap, os, time

import numpy as np
import torch
import torch.distributed as dist

N = 88 * 1024 * 1024


def thp_buffer(nbytes):
    mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    arr = np.frombuffer(mm, dtype=np.uint8)
    base = arr.ctypes.data
    off = (-base) % (2 << 20)
    libc = ctypes.CDLL("libc.so.6", use_errno=True)
    rc = libc.madvise(ctypes.c_void_p(base + off), ctypes.c_size_t(nbytes), 14)   # MADV_HUGEPAGE
    view = arr[off:off + nbytes]
    view[:] = 0                                                                    # touch: fault the pages in
    r = torch.cuda.cudart().cudaHostRegister(base + off, nbytes, 0)
    return mm, view, rc, r


def anon_huge_kb():
    tot = 0
    for line in open("/proc/self/smaps"):
        if line.startswith("AnonHugePages:"):
            tot += int(line.split()[1])
    return tot


def rate(dev, host_t, world, label):
    d = torch.empty(N, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(4):
            host_t.copy_(d, non_blocking=True)
        s.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = 48
        for _ in range(reps):
            host_t.copy_(d, non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
    print("rank %d %-26s d2h %.1f GB/s (pinned=%s)" % (int(os.environ.get("RANK", 0)), label, reps * N / dt / 1e9,
                                                      host_t.is_pinned()), flush=True)


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        for p in ("/sys/kernel/mm/transparent_hugepage/enabled", "/sys/kernel/mm/transparent_hugepage/defrag",
                  "/proc/sys/vm/nr_hugepages"):
            try:
                print(p, open(p).read().strip())
            except OSError as e:
                print(p, e)
    a = torch.empty(N, dtype=torch.uint8, pin_memory=True)
    rate(dev, a, world, "cudaHostAlloc")
    before = anon_huge_kb()
    mm, view, rc, r = thp_buffer(N)
    print("rank %d madvise rc=%d register=%s AnonHugePages +%d kB" % (rank, rc, r, anon_huge_kb() - before), flush=True)
    b = torch.from_numpy(view)
    rate(dev, b, world, "THP + cudaHostRegister")
    rate(dev, a, world, "cudaHostAlloc (again)")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

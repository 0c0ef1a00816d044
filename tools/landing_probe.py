#!/usr/bin/env python
"""How fast can device frames land in a FILE on the GPU box?  (K5 / extract_segment's `.frames` artefact.)

Measures, for a fresh file of SIZE bytes on /dev/shm (tmpfs) and on the root disk:
  a. page allocation: posix_fallocate with 1..T threads over disjoint ranges
  b. mmap(MAP_SHARED) + MADV_POPULATE_WRITE of the allocated file
  c. cudaHostRegister of the mapping (whole, and in 256 MiB pieces)
  d. D2H copies straight into the registered mapping (no host copy at all)
  e. pinned staging buffer -> os.pwrite with 1..T threads, cold file and rewritten (warm) file
  f. d + e at once (the copy engine and the writer threads share the host memory system)
Prints one JSON object.
"""
import ctypes
import json
import mmap
import os
import sys
import threading
import time

import numpy as np
import torch

SIZE = int(float(os.environ.get("PROBE_GB", "4")) * (1 << 30))
CHUNK = 44 << 20
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.posix_fallocate.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_long]
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
rt = torch.cuda.cudart()
out = {"size": SIZE}


def threads_do(n, fn):
    ts = [threading.Thread(target=fn, args=(i, n)) for i in range(n)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def falloc(path, n_threads):
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    os.ftruncate(fd, SIZE)
    step = (SIZE // n_threads + 4095) // 4096 * 4096

    def work(i, n):
        lo = i * step
        ln = min(step, SIZE - lo)
        if ln > 0:
            libc.posix_fallocate(fd, lo, ln)

    dt = threads_do(n_threads, work)
    return fd, SIZE / dt / 1e9


def probe_dir(d, tag):
    res = {}
    path = os.path.join(d, "vt_landing_probe.bin")
    for nt in (1, 4, 8, 16):
        fd, r = falloc(path, nt)
        res["fallocate_%dt_gbs" % nt] = r
        os.close(fd)
        os.unlink(path)
    fd, _ = falloc(path, 8)
    t0 = time.perf_counter()
    mm = mmap.mmap(fd, SIZE, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
    arr = np.frombuffer(mm, dtype=np.uint8)
    base = arr.ctypes.data
    step = (SIZE // 8 + 4095) // 4096 * 4096
    rcs = []

    def pop(i, n):
        lo = i * step
        ln = min(step, SIZE - lo)
        if ln > 0:
            rcs.append(libc.madvise(base + lo, ln, 23))   # MADV_POPULATE_WRITE

    dt = threads_do(8, pop)
    res["populate_8t_gbs"] = SIZE / dt / 1e9
    res["populate_rc"] = sorted(set(rcs))
    t0 = time.perf_counter()
    r = rt.cudaHostRegister(base, SIZE, 0)
    dt = time.perf_counter() - t0
    res["register_rc"] = int(r)
    res["register_gbs"] = SIZE / dt / 1e9
    if int(r) == 0:
        dev = torch.empty(CHUNK, dtype=torch.uint8, device="cuda")
        host = torch.from_numpy(arr)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            n = SIZE // CHUNK
            for rep in range(2):
                s.synchronize()
                t0 = time.perf_counter()
                for i in range(n):
                    host[i * CHUNK:(i + 1) * CHUNK].copy_(dev, non_blocking=True)
                s.synchronize()
                dt = time.perf_counter() - t0
                res["d2h_into_mapping_gbs_pass%d" % rep] = n * CHUNK / dt / 1e9
        t0 = time.perf_counter()
        rt.cudaHostUnregister(base)
        res["unregister_gbs"] = SIZE / (time.perf_counter() - t0) / 1e9
        # piecewise register (pipelinable with copies)
        piece = 256 << 20
        t0 = time.perf_counter()
        k = 0
        for lo in range(0, SIZE, piece):
            rr = rt.cudaHostRegister(base + lo, min(piece, SIZE - lo), 0)
            k += int(rr) == 0
        dt = time.perf_counter() - t0
        res["register_256m_pieces_gbs"] = SIZE / dt / 1e9
        res["register_256m_pieces_ok"] = k
        for lo in range(0, SIZE, piece):
            rt.cudaHostUnregister(base + lo)
        del host
    del arr
    try:
        mm.close()
    except BufferError:
        pass
    os.close(fd)
    os.unlink(path)

    # pinned staging -> pwrite threads
    stage = torch.empty(CHUNK * 4, dtype=torch.uint8, pin_memory=True)
    stage.fill_(7)
    mv = memoryview(stage.numpy())
    for nt in (1, 4, 8, 12, 16):
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        n_chunks = SIZE // CHUNK

        def wr(i, n):
            for c in range(i, n_chunks, n):
                os.pwrite(fd, mv[(c % 4) * CHUNK:(c % 4 + 1) * CHUNK], c * CHUNK)

        dt = threads_do(nt, wr)
        res["pwrite_cold_%dt_gbs" % nt] = n_chunks * CHUNK / dt / 1e9
        dt = threads_do(nt, wr)
        res["pwrite_warm_%dt_gbs" % nt] = n_chunks * CHUNK / dt / 1e9
        os.close(fd)
        os.unlink(path)
    # D2H into the pinned ring while 8 threads write it out (cold file)
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    dev = torch.empty(CHUNK, dtype=torch.uint8, device="cuda")
    n_chunks = SIZE // CHUNK
    s = torch.cuda.Stream()
    evs = [torch.cuda.Event() for _ in range(4)]
    done = [threading.Event() for _ in range(n_chunks)]
    freed = [threading.Semaphore(0) for _ in range(4)]
    for f in freed:
        f.release()
    pending = [0] * 4

    def wr2(i, n):
        for c in range(i, n_chunks, n):
            done[c].wait()
            os.pwrite(fd, mv[(c % 4) * CHUNK:(c % 4 + 1) * CHUNK], c * CHUNK)
            freed[c % 4].release()

    ts = [threading.Thread(target=wr2, args=(i, 4)) for i in range(4)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    with torch.cuda.stream(s):
        for c in range(n_chunks):
            freed[c % 4].acquire()
            stage[(c % 4) * CHUNK:(c % 4 + 1) * CHUNK].copy_(dev, non_blocking=True)
            evs[c % 4].record(s)
            evs[c % 4].synchronize()
            done[c].set()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    res["d2h_plus_pwrite_4slots_gbs"] = n_chunks * CHUNK / dt / 1e9
    os.close(fd)
    os.unlink(path)
    out[tag] = res


for d, tag in (("/dev/shm", "shm"), (os.environ.get("PROBE_DISK_DIR", "/tmp"), "disk")):
    try:
        probe_dir(d, tag)
    except Exception as e:  # noqa: BLE001
        out[tag + "_error"] = repr(e)
try:
    out["thp_shmem"] = open("/sys/kernel/mm/transparent_hugepage/shmem_enabled").read().strip()
except OSError as e:
    out["thp_shmem"] = repr(e)
print(json.dumps(out, indent=1))

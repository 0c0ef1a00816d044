"""What does the FIRST landing file of a process cost, and can it be made cheaper?  (A 600 s segment of 720p frames is
24.9 GB: tools/full_config1.py measured 8.9 s for allocate + page-lock, of a 15.7 s job.)

    gpurun -- python tools/cold_landing_probe.py [GB]
"""
import ctypes
import mmap
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_transformer_b200._lib import lib  # noqa: E402

GB = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
N = int(GB * (1 << 30)) & ~((2 << 20) - 1)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.posix_fallocate.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_long]
L = lib()
PATH = "/dev/shm/_cold_probe.bin"


def T(label, fn, nbytes=N):
    t = time.perf_counter()
    r = fn()
    dt = time.perf_counter() - t
    print("%-64s %8.1f ms  %6.2f GB/s" % (label, dt * 1e3, nbytes / dt / 1e9), flush=True)
    return r


def fresh():
    if os.path.exists(PATH):
        os.unlink(PATH)
    return os.open(PATH, os.O_RDWR | os.O_CREAT | os.O_EXCL, 0o644)


def in_threads(n, fn):
    step = (N // n) & ~((2 << 20) - 1)
    ths = [threading.Thread(target=fn, args=(i * step, step if i < n - 1 else N - i * step)) for i in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()


def reg(base, off, ln):
    rc = L.vt_host_register(ctypes.c_void_p(base + off), ln)
    assert rc == 0, rc


def unreg_all(base, offs):
    for o in offs:
        L.vt_host_unregister(ctypes.c_void_p(base + o))


def main():
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    print("file of %.1f GB on /dev/shm" % (N / (1 << 30)))
    # A: what the product does today
    fd = fresh()
    T("A  posix_fallocate, one call", lambda: libc.posix_fallocate(fd, 0, N))
    mm = mmap.mmap(fd, N)
    base = np.frombuffer(mm, dtype=np.uint8).ctypes.data
    T("A  cudaHostRegister, one call", lambda: reg(base, 0, N))
    d = torch.ones(64 << 20, dtype=torch.uint8, device="cuda")
    host = torch.from_numpy(np.frombuffer(mm, dtype=np.uint8))

    def d2h(at):
        host[at:at + d.numel()].copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    d2h(0)
    T("A  D2H 64 MB into it", lambda: d2h(128 << 20), d.numel())
    T("A  cudaHostUnregister", lambda: unreg_all(base, [0]))
    # again on the same (now existing, once-pinned) pages
    T("A' cudaHostRegister again (pages exist)", lambda: reg(base, 0, N))
    unreg_all(base, [0])
    # C: chunked registration of existing pages, sequential and threaded
    step = 512 << 20
    offs = list(range(0, N, step))
    T("C  register in 512 MB chunks, sequential (pages exist)", lambda: [reg(base, o, min(step, N - o)) for o in offs])
    # E: a copy that spans two registrations is refused (cudaErrorInvalidValue, measured): copies must be split there
    d2h(step - (64 << 20))
    T("E  D2H 64 MB ending at a registration boundary", lambda: d2h(step - (64 << 20)), d.numel())
    T("E  D2H 64 MB starting at a registration boundary", lambda: d2h(step), d.numel())
    unreg_all(base, offs)
    for nt in (2, 4, 8):
        got = []
        T("C  register in %d threads (pages exist)" % nt,
          lambda: in_threads(nt, lambda o, ln: (reg(base, o, ln), got.append(o))))
        unreg_all(base, got)
    del host
    mm.close()
    os.close(fd)
    # B: threaded fallocate on a fresh file
    for nt in (4,):
        fd = fresh()
        os.ftruncate(fd, N)
        T("B  posix_fallocate in %d threads (fresh file)" % nt,
          lambda: in_threads(nt, lambda o, ln: libc.posix_fallocate(fd, o, ln)))
        mm = mmap.mmap(fd, N)
        base = np.frombuffer(mm, dtype=np.uint8).ctypes.data
        got = []
        T("B  register in %d threads (fresh pages)" % nt,
          lambda: in_threads(nt, lambda o, ln: (reg(base, o, ln), got.append(o))))
        unreg_all(base, got)
        mm.close()
        os.close(fd)
    # D: no fallocate at all: the pinning faults the pages in
    fd = fresh()
    os.ftruncate(fd, N)
    mm = mmap.mmap(fd, N)
    base = np.frombuffer(mm, dtype=np.uint8).ctypes.data
    T("D  ftruncate only, cudaHostRegister faults the pages in", lambda: reg(base, 0, N))
    unreg_all(base, [0])
    mm.close()
    os.close(fd)
    # F: MAP_POPULATE instead of fallocate
    fd = fresh()
    os.ftruncate(fd, N)
    mm = T("F  ftruncate + mmap(MAP_POPULATE)", lambda: mmap.mmap(fd, N, flags=mmap.MAP_SHARED | mmap.MAP_POPULATE))
    base = np.frombuffer(mm, dtype=np.uint8).ctypes.data
    T("F  cudaHostRegister after populate", lambda: reg(base, 0, N))
    unreg_all(base, [0])
    mm.close()
    os.close(fd)
    # G: pinned ring + pwrite (the staged writer) for comparison, one thread
    fd = fresh()
    pin = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    mv = memoryview(pin.numpy())

    def staged():
        off = 0
        while off < N:
            n = min(len(mv), N - off)
            done = 0
            while done < n:
                done += os.pwrite(fd, mv[done:n], off + done)
            off += n
    T("G  pwrite from pinned memory into a fresh file", staged)
    os.close(fd)
    os.unlink(PATH)


if __name__ == "__main__":
    main()

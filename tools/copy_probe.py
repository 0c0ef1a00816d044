#!/usr/bin/env python
"""File -> file copy rates on the box's /dev/shm (the stream copy of extract_segment): sendfile, copy_file_range,
pread+pwrite, pwrite from an mmap of the source; into a fresh file (cold) and over an existing one (warm)."""
import json, mmap, os, time
SIZE = 440 << 20
d = "/dev/shm"
src = os.path.join(d, "vt_copy_src.bin")
dst = os.path.join(d, "vt_copy_dst.bin")
with open(src, "wb") as f:
    f.write(os.urandom(1 << 20) * (SIZE >> 20))
out = {}


def run(name, fn):
    for mode in ("cold", "warm"):
        if mode == "cold" and os.path.exists(dst):
            os.unlink(dst)
        fi = os.open(src, os.O_RDONLY)
        fo = os.open(dst, os.O_RDWR | os.O_CREAT, 0o644)
        t0 = time.perf_counter()
        fn(fi, fo)
        os.ftruncate(fo, SIZE)
        dt = time.perf_counter() - t0
        os.close(fi); os.close(fo)
        out["%s_%s_gbs" % (name, mode)] = SIZE / dt / 1e9


def sendfile(fi, fo):
    done = 0
    while done < SIZE:
        done += os.sendfile(fo, fi, done, SIZE - done)


def cfr(fi, fo):
    done = 0
    while done < SIZE:
        done += os.copy_file_range(fi, fo, SIZE - done, done, done)


def rw(fi, fo, chunk=8 << 20):
    done = 0
    while done < SIZE:
        b = os.pread(fi, chunk, done)
        os.pwrite(fo, b, done)
        done += len(b)


def mm_pwrite(fi, fo, chunk=32 << 20):
    m = mmap.mmap(fi, SIZE, prot=mmap.PROT_READ)
    mv = memoryview(m)
    done = 0
    while done < SIZE:
        done += os.pwrite(fo, mv[done:done + chunk], done)
    mv.release(); m.close()


def mm_mm(fi, fo):
    os.ftruncate(fo, SIZE)
    m = mmap.mmap(fi, SIZE, prot=mmap.PROT_READ)
    o = mmap.mmap(fo, SIZE)
    o[:] = m[:]
    m.close(); o.close()


for name, fn in (("sendfile", sendfile), ("copy_file_range", cfr), ("pread_pwrite_8m", rw), ("mmap_src_pwrite", mm_pwrite),
                 ("mmap_to_mmap", mm_mm)):
    try:
        run(name, fn)
    except Exception as e:  # noqa: BLE001
        out[name + "_error"] = repr(e)
os.unlink(src); os.unlink(dst)
print(json.dumps(out, indent=1))

# ---- several threads copying disjoint ranges of ONE output file (is the write path serialised per file?) ----
import threading
with open(src, "wb") as f:
    f.write(os.urandom(1 << 20) * (SIZE >> 20))
res2 = {}
for nt in (1, 2, 4):
    for mode in ("cold", "warm"):
        if mode == "cold" and os.path.exists(dst):
            os.unlink(dst)
        fi = os.open(src, os.O_RDONLY)
        fo = os.open(dst, os.O_RDWR | os.O_CREAT, 0o644)
        step = SIZE // nt

        def work(i):
            lo, hi = i * step, (i + 1) * step if i + 1 < nt else SIZE
            done = lo
            while done < hi:
                done += os.copy_file_range(fi, fo, hi - done, done, done)

        ts = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.perf_counter() - t0
        os.close(fi); os.close(fo)
        res2["copy_file_range_%dthreads_%s_gbs" % (nt, mode)] = SIZE / dt / 1e9
os.unlink(src); os.unlink(dst)
print(json.dumps(res2, indent=1))

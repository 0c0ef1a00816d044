"""K5: where a segment's frame buffers land -- directly in the `.frames` file, with no host copy.

The artefact of extract_segment (/root/reference/src/utils/video_segmenter.py:86-154 leaves ONE file per call) gets a
sibling `<segment>.frames`.  Instead of copying device frames into a pinned staging buffer and then `write()`-ing
them (two passes over host memory, the second one in the file system), the file itself is the copy target:

    create the file at its final size -> mmap(MAP_SHARED) -> cudaHostRegister(mapping) -> cudaMemcpyAsync D2H into it

Registering is the expensive part (measured on the B200 box, tmpfs: allocate 13.6 GB/s, register 4.3 GB/s, D2H into
the registered mapping 50-56 GB/s = the PCIe ceiling; tools/landing_probe.py), so registered files are RECYCLED:
each lives under a hidden arena directory beside the outputs and the `.frames` name is a hard link to it.  When the
consumer deletes (or this module replaces) the `.frames` name, the arena file's link count drops back to 1 and the
next segment of the same size lands in the already-registered pages.  File systems whose mappings cannot be
registered (anything but tmpfs here) get the classic path: pinned ring -> a writer thread -> pwrite.
"""
from __future__ import annotations

import atexit
import ctypes
import mmap
import os
import threading
from pathlib import Path

import numpy as np
import torch

from ._lib import lib

ARENA_DIR = ".vt_landing"
_libc = ctypes.CDLL("libc.so.6", use_errno=True)
_libc.posix_fallocate.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_long]
_lock = threading.Lock()
_slots: list["_Slot"] = []
_seq = 0
_arena_of_device: dict = {}      # st_dev -> arena directory: hard links work anywhere on one file system, so the
                                 # segments of every video written there share one arena


class _Slot:
    """One registered file mapping."""

    def __init__(self, path: Path, nbytes: int, register: bool = True):
        self.path, self.nbytes = path, nbytes
        self.pid = os.getpid()           # a forked child inherits this object but does not own the file or the pinning
        self.fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_EXCL, 0o644)
        try:
            if register:
                rc = _libc.posix_fallocate(self.fd, 0, nbytes)
                if rc != 0:
                    raise OSError(rc, "posix_fallocate(%d bytes) failed" % nbytes)
                self.mm = mmap.mmap(self.fd, nbytes, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
        except Exception:
            os.close(self.fd)
            os.unlink(path)
            raise
        self.plain = not register
        if self.plain:                 # an MP4 arena file: first filled through its descriptor, mapped afterwards
            self.mm = self.array = self.tensor = None
            self.base = 0
            self.registered = False
            self.stamp = 0
            return
        self.array = np.frombuffer(self.mm, dtype=np.uint8)
        self.base = self.array.ctypes.data
        self.registered = register and lib().vt_host_register(ctypes.c_void_p(self.base), nbytes) == 0
        self.tensor = torch.from_numpy(self.array) if register else None
        self.stamp = 0

    def free(self) -> bool:
        try:
            return os.fstat(self.fd).st_nlink <= 1 and os.path.exists(self.path)
        except OSError:
            return False

    def destroy(self) -> None:
        if self.pid != os.getpid():      # inherited through fork(): the parent unregisters and unlinks
            self.registered = False
            self.tensor = self.array = None
            return
        if self.registered:
            try:
                lib().vt_host_unregister(ctypes.c_void_p(self.base))
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self.registered = False
        self.tensor = None
        self.array = None
        try:
            if self.mm is not None:
                self.mm.close()
        except (BufferError, ValueError):
            pass
        try:
            os.close(self.fd)
        except OSError:
            pass
        try:
            os.unlink(self.path)
        except OSError:
            pass


class Landing:
    """A `.frames` file being filled.  `tensor` (uint8, [nbytes]) is host memory the copy engine writes into when
    `direct` is True; otherwise frames arrive through write_chunk() from a pinned staging buffer."""

    def __init__(self, path: Path, nbytes: int, slot: _Slot | None, recycled: bool):
        self.path, self.nbytes = path, nbytes
        self.slot = slot
        self.direct = slot is not None
        self.recycled = recycled
        self.tensor = slot.tensor[:nbytes] if slot is not None else None
        self._fd = None
        self._writer = None
        if slot is None:
            self._fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
            from concurrent.futures import ThreadPoolExecutor
            self._writer = ThreadPoolExecutor(max_workers=1)
            self._pending = []

    # -- staged path -----------------------------------------------------------------------------------------
    def write_chunk(self, chunk: torch.Tensor, byte_offset: int):
        """Queue `chunk` (pinned host memory, contiguous) for pwrite at byte_offset; returns a future the caller waits on
        before reusing the staging buffer."""
        mv = memoryview(chunk.numpy()).cast("B")
        fut = self._writer.submit(_pwrite_all, self._fd, mv, byte_offset)
        self._pending.append(fut)
        return fut

    def finish(self, nbytes_used: int | None = None) -> None:
        """All copies into the landing have completed (the caller synchronised its stream)."""
        if self._writer is not None:
            for f in self._pending:
                f.result()
            self._writer.shutdown()
            if nbytes_used is not None:
                os.ftruncate(self._fd, nbytes_used)
            os.close(self._fd)
            self._fd = None

    def abort(self) -> None:
        if self._writer is not None:
            self._writer.shutdown(wait=True)
            if self._fd is not None:
                os.close(self._fd)
        try:
            os.unlink(self.path)
        except OSError:
            pass


def _pwrite_all(fd: int, mv: memoryview, offset: int) -> None:
    done = 0
    while done < len(mv):
        done += os.pwrite(fd, mv[done:], offset + done)


def _arena_cap() -> int:
    """Page-locked landing memory this process keeps (registered files, in use or free).  At the cap the oldest FREE
    files are released (unregistered, unmapped, their arena name removed); files whose `.frames` name still exists are
    left alone and new outputs take the staged writer instead (_make_room).  (An unbounded arena made every later
    page-locking call of the process slower: 64 retained 415 MB files took engine set-up from 30 ms to 530 ms in the
    configs[3] batch.)"""
    return int(float(os.environ.get("VT_LANDING_CAP_GB", "12")) * (1 << 30))


def acquire(path: str | Path, nbytes: int, direct: bool | None = None) -> Landing:
    """Prepare `path` (a `.frames` file of exactly nbytes) as a D2H copy target.

    direct=None: try the registered-mapping path and fall back to the staged writer when the file system's mappings
    cannot be registered; True: registered mapping or OSError; False: staged writer."""
    global _seq
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    if nbytes <= 0:
        raise ValueError("landing size must be positive")
    if direct is None:
        direct = os.environ.get("VT_LANDING", "direct") != "staged"
    slot = None
    room = False
    with _lock:
        if path.exists() or path.is_symlink():
            path.unlink()                       # a replaced artefact frees its arena file for the search below
        if direct:
            arena = _arena_for(path.parent)
            for s in _slots:
                if s.nbytes == nbytes and not s.plain and s.path.parent == arena and s.free():
                    slot = s
                    break
            if slot is not None:
                try:
                    os.link(slot.path, path)
                except OSError:
                    slot.destroy()
                    _slots.remove(slot)
                    slot = None
                else:
                    _seq += 1
                    slot.stamp = _seq
                    return Landing(path, nbytes, slot, True)
            # Nothing to recycle.  A new registered file is worth its price (allocate + page-lock, ~3 GB/s) only while
            # the arena has room: FREE files are released to make it, files whose `.frames` name still exists are not
            # -- a consumer that keeps every output would otherwise make each segment pay for a registration AND an
            # un-registration (64 kept 415 MB clips: 1.1 s per clip), where the staged writer costs 0.1 s.
            room = _make_room(nbytes, False, _arena_cap())
            _seq += 1
            name = arena / ("landing_%d_%d.bin" % (os.getpid(), _seq))
    if direct and room:                         # the slow part runs outside the lock (the MP4 writer takes it too)
        try:
            arena.mkdir(exist_ok=True)
            slot = _Slot(name, nbytes)
            if not slot.registered:
                slot.destroy()
                slot = None
            else:
                os.link(slot.path, path)
        except OSError:
            if slot is not None:
                slot.destroy()
            slot = None
        if slot is not None:
            with _lock:
                _seq += 1
                slot.stamp = _seq
                _slots.append(slot)
            return Landing(path, nbytes, slot, False)
    if direct is True and os.environ.get("VT_LANDING") == "direct-only":
        raise OSError("cannot register a mapping of %s" % path)
    return Landing(path, nbytes, None, False)


def _make_room(nbytes: int, plain: bool, cap: int) -> bool:
    """Release the oldest FREE arena files of one kind until nbytes more fit under cap; False if they still do not
    (the rest are outputs somebody still holds: they stay as they are).  One file larger than the whole cap is allowed
    when it would be the only one: a 720 s segment of 720p frames is 30 GB, and recycling that file is worth far more
    than any number of small ones."""
    mine = [s for s in _slots if s.plain == plain]
    used = sum(s.nbytes for s in mine)
    for s in sorted((s for s in mine if s.free()), key=lambda s: s.stamp):
        if used + nbytes <= cap:
            break
        s.destroy()
        _slots.remove(s)
        used -= s.nbytes
    return used + nbytes <= cap or used == 0


def _arena_for(parent: Path) -> Path:
    try:
        dev_id = os.stat(parent).st_dev
    except OSError:
        dev_id = None
    arena = _arena_of_device.get(dev_id)
    if arena is None or not arena.is_dir():
        arena = parent / ARENA_DIR
        _arena_of_device[dev_id] = arena
        _sweep_stale(arena)
    return arena


_MEMORY_FS_MAGIC = {0x01021994, 0x858458F6}     # tmpfs, ramfs (statfs f_type)


_memory_fs_of_device: dict = {}


def on_memory_fs(path: Path) -> bool:
    """True when `path` lives on a RAM-backed file system (statfs f_type; /proc/self/mounts as the second opinion).
    The answer is kept per st_dev."""
    try:
        dev_id = os.stat(path).st_dev
    except OSError:
        return False
    hit = _memory_fs_of_device.get(dev_id)
    if hit is None:
        hit = _memory_fs_of_device[dev_id] = _probe_memory_fs(path)
    return hit


def _probe_memory_fs(path: Path) -> bool:
    buf = ctypes.create_string_buffer(256)
    try:
        if _libc.statfs(os.fsencode(str(path)), buf) == 0:
            if ctypes.c_long.from_buffer(buf).value & 0xFFFFFFFF in _MEMORY_FS_MAGIC:
                return True
    except (OSError, AttributeError):
        pass
    try:
        best, kind = "", ""
        real = os.path.realpath(path)
        with open("/proc/self/mounts") as f:
            for line in f:
                parts = line.split()
                if len(parts) >= 3 and (real == parts[1] or real.startswith(parts[1].rstrip("/") + "/")) \
                        and len(parts[1]) >= len(best):
                    best, kind = parts[1], parts[2]
        return kind in ("tmpfs", "ramfs")
    except OSError:
        return False


class MappedFile:
    """An output file (a segment's MP4) in the arena.  Recycled: `array[:nbytes]` IS the file (a long-lived shared
    mapping whose pages exist).  New: `array` is None -- fill the file through `fd` (the in-kernel copy allocates
    pages several times faster than first-touch faults through a mapping: 175 vs 279 ms for 220 MB in this container),
    then call populate() so that the NEXT segment finds a warm mapping."""

    def __init__(self, path: Path, nbytes: int, slot: _Slot, recycled: bool):
        self.path, self.nbytes, self.slot, self.recycled = path, nbytes, slot, recycled
        self.fd = slot.fd
        self.array = slot.array[:nbytes] if slot.array is not None else None

    def populate(self) -> None:
        """Map the (filled) file over the slot's capacity and pre-fault the pages it has."""
        s = self.slot
        if s.mm is not None:
            return
        with _lock:
            try:
                os.ftruncate(s.fd, s.nbytes)
                s.mm = mmap.mmap(s.fd, s.nbytes, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
                os.ftruncate(s.fd, self.nbytes)
                s.array = np.frombuffer(s.mm, dtype=np.uint8)
            except (OSError, ValueError):
                s.mm = s.array = None
                os.ftruncate(s.fd, self.nbytes)
                return
        try:
            s.mm.madvise(getattr(mmap, "MADV_POPULATE_WRITE", 23), 0, self.nbytes & ~(mmap.PAGESIZE - 1))
        except (OSError, ValueError, AttributeError):
            pass                                    # older kernels: the first recycled copy takes the faults instead

    def abort(self) -> None:
        try:
            os.unlink(self.path)
        except OSError:
            pass


def _mapped_cap() -> int:
    return int(float(os.environ.get("VT_MP4_ARENA_CAP_GB", "8")) * (1 << 30))


def acquire_mapped(path: str | Path, nbytes: int) -> MappedFile | None:
    """`path` as a file of exactly nbytes whose content is written by plain stores into a recycled mapping.

    The stream copy of a segment (K4) is a 100-400 MB file-to-file copy.  Through write()/copy_file_range() every
    call allocates the output's pages anew (3.6-5 GB/s on the B200 box's tmpfs, one thread, and the rate does not
    scale with threads); into a mapping whose pages already exist it is a memcpy (7-8 GB/s per thread).  The same
    recycling as the `.frames` landing makes the pages exist: the MP4 name is a hard link to an arena file, and once
    the consumer deletes the name the next segment is copied into the same pages.  The mapping covers the slot's
    capacity; the FILE is ftruncate()d to the exact size each time, so only the tail pages churn between segments of
    slightly different sizes.  Returns None when the file system is not RAM-backed (dirty shared mappings of disk files
    buy nothing) or the arena cannot be set up -- the caller then uses the in-kernel copy."""
    global _seq
    path = Path(path)
    if nbytes <= 0 or os.environ.get("VT_MP4_ARENA", "1") == "0":
        return None
    path.parent.mkdir(parents=True, exist_ok=True)
    if not on_memory_fs(path.parent):
        return None
    with _lock:
        if path.exists() or path.is_symlink():
            path.unlink()
        arena = _arena_for(path.parent)
        slot = None
        for s in _slots:
            if s.plain and s.array is not None and s.path.parent == arena \
                    and nbytes <= s.nbytes <= 2 * nbytes + (64 << 20) and s.free() \
                    and (slot is None or s.nbytes < slot.nbytes):
                slot = s
        recycled = slot is not None
        try:
            if slot is None:
                cap = -(-(nbytes + nbytes // 8) // (16 << 20)) * (16 << 20)
                if not _make_room(cap, True, _mapped_cap()):
                    return None                 # every arena file is an output somebody still holds: plain copy
                arena.mkdir(exist_ok=True)
                _seq += 1
                slot = _Slot(arena / ("landing_%d_%d.bin" % (os.getpid(), _seq)), cap, register=False)
            os.ftruncate(slot.fd, nbytes)
            os.link(slot.path, path)
        except OSError:
            if slot is not None:
                slot.destroy()
                if recycled:
                    _slots.remove(slot)
            return None
        _seq += 1
        slot.stamp = _seq
        if not recycled:
            _slots.append(slot)
        return MappedFile(path, nbytes, slot, recycled)


def _sweep_stale(arena: Path) -> None:
    """Remove arena names left behind by processes that no longer exist (a crash skips release_all).  The data of a
    `.frames` file that is still linked survives: only the extra name goes."""
    try:
        names = list(arena.iterdir())
    except OSError:
        return
    for p in names:
        parts = p.name.split("_")
        if len(parts) != 3 or parts[0] != "landing" or not parts[1].isdigit():
            continue
        pid = int(parts[1])
        if pid == os.getpid():
            continue
        try:
            os.kill(pid, 0)
        except ProcessLookupError:
            try:
                p.unlink()
            except OSError:
                pass
        except OSError:
            pass                                # exists but not ours to signal: leave it


def stats() -> dict:
    with _lock:
        return {"files": len(_slots), "bytes": sum(s.nbytes for s in _slots),
                "free": sum(1 for s in _slots if s.free()), "mapped_files": sum(1 for s in _slots if s.plain)}


def release_all() -> None:
    """Unregister and remove every arena file (the `.frames` hard links stay)."""
    with _lock:
        for s in _slots:
            s.destroy()
        _slots.clear()


atexit.register(release_all)

"""K5 landing, host logic that needs no GPU: the staged writer (file systems / machines whose mappings cannot be
page-locked) and the clean-up of arena names left by dead processes.  The direct path (registered mapping, recycling)
is a GPU test: tests/test_ingest.py::test_frames_land_directly_in_the_file_and_landing_files_are_recycled."""
import os

import numpy as np
import torch

from video_transformer_b200 import landing


def test_staged_landing_writes_chunks_at_their_offsets(tmp_path):
    path = tmp_path / "seg" / "segment_0000.frames"
    fb, n = 1000, 7
    land = landing.acquire(path, fb * n, direct=False)
    assert not land.direct and land.tensor is None
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, (n, fb), dtype=np.uint8)
    futs = []
    for a, b in ((4, 7), (0, 2), (2, 4)):                       # out of order on purpose: offsets decide, not call order
        futs.append(land.write_chunk(torch.from_numpy(data[a:b].copy()), a * fb))
    for f in futs:
        f.result()
    land.finish(fb * n)
    assert path.read_bytes() == data.tobytes()
    # a replaced artefact starts from an empty file
    land = landing.acquire(path, fb * 2, direct=False)
    land.write_chunk(torch.from_numpy(data[:2].copy()), 0).result()
    land.finish(fb * 2)
    assert path.stat().st_size == fb * 2
    land = landing.acquire(path, fb, direct=False)
    land.abort()
    assert not path.exists()


def test_without_a_gpu_the_direct_request_falls_back_to_the_staged_writer(tmp_path):
    if torch.cuda.is_available():
        return                                                   # covered by the GPU test
    land = landing.acquire(tmp_path / "x.frames", 4096)          # registration fails: no device
    assert not land.direct
    land.abort()
    assert landing.stats()["files"] == 0
    assert not any((tmp_path / landing.ARENA_DIR).glob("*")) if (tmp_path / landing.ARENA_DIR).exists() else True


def test_stale_arena_names_are_swept(tmp_path):
    arena = tmp_path / landing.ARENA_DIR
    arena.mkdir()
    dead = arena / "landing_999999_3.bin"
    mine = arena / ("landing_%d_1.bin" % os.getpid())
    other = arena / "notes.txt"
    for p in (dead, mine, other):
        p.write_bytes(b"x")
    frames = tmp_path / "segment_0001.frames"
    os.link(dead, frames)                                        # a crashed process left both names
    landing._sweep_stale(arena)
    assert not dead.exists() and mine.exists() and other.exists()
    assert frames.read_bytes() == b"x"                           # the artefact survives, only the extra name went


def _shm_dir():
    import tempfile
    if not os.path.isdir("/dev/shm") or not landing.on_memory_fs("/dev/shm"):
        import pytest
        pytest.skip("no RAM-backed file system here")
    return tempfile.mkdtemp(prefix="vt_test_", dir="/dev/shm")


def test_mapped_output_is_refused_off_ram(tmp_path):
    if landing.on_memory_fs(tmp_path):
        import pytest
        pytest.skip("tmp_path is RAM-backed here")
    assert landing.acquire_mapped(tmp_path / "a.mp4", 4096) is None
    assert not (tmp_path / "a.mp4").exists()


def test_mapped_output_recycles_pages_and_sizes_the_file_exactly():
    import shutil
    d = _shm_dir()
    try:
        from pathlib import Path
        a = landing.acquire_mapped(Path(d) / "a.mp4", 1_000_003)
        # a NEW arena file is filled through its descriptor (no mapping yet), then mapped for the segments after it
        assert a is not None and not a.recycled and a.array is None
        os.pwrite(a.fd, b"\x07" * 1_000_003, 0)
        a.populate()
        assert os.path.getsize(a.path) == 1_000_003 and a.path.read_bytes() == b"\x07" * 1_000_003
        b = landing.acquire_mapped(Path(d) / "b.mp4", 900_000)             # a.mp4 still exists: a second arena file
        assert b is not None and not b.recycled and b.slot is not a.slot
        os.pwrite(b.fd, b"\x01" * 900_000, 0)
        b.populate()
        os.unlink(a.path)                                                  # the consumer is done with a.mp4
        c = landing.acquire_mapped(Path(d) / "c.mp4", 1_100_000)           # a little larger: still fits the capacity
        assert c is not None and c.recycled and c.slot is a.slot
        assert c.array is not None and c.array.size == 1_100_000
        c.array[:] = 9
        assert os.path.getsize(c.path) == 1_100_000 and c.path.read_bytes() == b"\x09" * 1_100_000
        assert b.path.read_bytes() == b"\x01" * 900_000                    # the other file is untouched by the reuse
        # replacing an existing name frees its arena file for the same call
        c2 = landing.acquire_mapped(c.path, 1_050_000)
        assert c2 is not None and c2.recycled and c2.slot is c.slot and os.path.getsize(c2.path) == 1_050_000
        assert landing.stats()["mapped_files"] == 2
    finally:
        landing.release_all()
        shutil.rmtree(d, ignore_errors=True)


def test_arena_never_evicts_outputs_that_still_exist(monkeypatch):
    """At the cap only FREE arena files are released; while every file is an output somebody still holds, new outputs
    are written the plain way instead of churning mappings."""
    import shutil
    from pathlib import Path
    d = _shm_dir()
    try:
        monkeypatch.setenv("VT_MP4_ARENA_CAP_GB", str(50 / 1024))          # 50 MB: three 16 MB granules fit, four do not
        outs = []
        for k in range(3):
            m = landing.acquire_mapped(Path(d) / ("seg_%d.mp4" % k), 10 << 20)
            assert m is not None
            os.pwrite(m.fd, bytes([k + 1]) * (10 << 20), 0)
            m.populate()
            outs.append(m)
        assert landing.acquire_mapped(Path(d) / "seg_3.mp4", 10 << 20) is None      # all three files are still held
        assert all(o.path.read_bytes() == bytes([k + 1]) * (10 << 20) for k, o in enumerate(outs))
        os.unlink(outs[0].path)                                                      # the consumer lets one go
        m = landing.acquire_mapped(Path(d) / "seg_3.mp4", 10 << 20)
        assert m is not None and m.recycled and m.slot is outs[0].slot
        os.unlink(m.path)
        os.unlink(outs[1].path)
        big = landing.acquire_mapped(Path(d) / "big.mp4", 17 << 20)                  # needs 32 MB: both free files go
        assert big is not None and not big.recycled and landing.stats()["mapped_files"] == 2
        assert outs[2].path.read_bytes() == bytes([3]) * (10 << 20)
    finally:
        landing.release_all()
        shutil.rmtree(d, ignore_errors=True)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_registered_arena_falls_back_to_the_staged_writer_at_its_cap(cuda, monkeypatch):
    """A consumer that keeps every `.frames` must not make each new segment pay for un-registering an old file and
    registering a new one: at the cap, with nothing free, the landing is the staged writer."""
    import shutil
    from pathlib import Path
    d = _shm_dir()
    try:
        monkeypatch.setenv("VT_LANDING_CAP_GB", str(3 / 1024))
        a = landing.acquire(Path(d) / "a.frames", 2 << 20)
        if not a.direct:
            pytest.skip("mappings of this file system cannot be registered")
        a.finish()
        b = landing.acquire(Path(d) / "b.frames", 2 << 20)                  # a.frames still exists, 2 + 2 > 3 MB
        assert not b.direct
        b.abort()
        assert os.path.getsize(a.path) == 2 << 20
        os.unlink(a.path)
        c = landing.acquire(Path(d) / "c.frames", 2 << 20)
        assert c.direct and c.recycled and c.slot is a.slot
        c.finish()
    finally:
        landing.release_all()
        shutil.rmtree(d, ignore_errors=True)

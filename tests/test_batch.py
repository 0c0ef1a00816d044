"""Batch pre-ingest: per-video sharding, progress.json semantics, manifest reuse."""
import json

import numpy as np
import pytest

from video_transformer_b200 import batch, container, synth, video_segmenter


def _write_clip(path, w, h, n, gop):
    bs, _ = synth.make_testsrc_h264(w, h, n, fps=30, gop=gop, cuts=[])
    raw = path.with_suffix(".h264")
    raw.write_bytes(bs)
    container.annexb_to_mp4(raw, path)
    raw.unlink()
    return path


def test_plan_batch_is_deterministic_and_balanced(tmp_path):
    paths = [_write_clip(tmp_path / ("v%02d.mp4" % i), 64, 48, n, 5) for i, n in enumerate([40, 10, 30, 20, 10, 10])]
    for world in (1, 2, 4):
        a = batch.plan_batch(paths, world)
        assert a == batch.plan_batch(paths, world)
        assert sorted(i for r in a for i in r) == list(range(6))
    a2 = batch.plan_batch(paths, 2)
    loads = [sum([40, 10, 30, 20, 10, 10][i] for i in r) for r in a2]
    assert abs(loads[0] - loads[1]) <= 10


def test_ingest_batch_progress_and_failure_contract(tmp_path, monkeypatch):
    """Host logic without a GPU: extract_segment is patched exactly like the reference's tests patch it
    (tests/test_long_video_integration.py:168-171)."""
    vids = [_write_clip(tmp_path / ("clip%d.mp4" % i), 64, 48, 30 * (i + 1), 10) for i in range(3)]
    calls = []

    def fake_extract(*, input_path, start, end, output_path, stream_copy=True):
        calls.append((str(input_path), start, end))
        if "clip1" in str(input_path):
            return False
        output_path.parent.mkdir(parents=True, exist_ok=True)
        output_path.write_bytes(b"x")
        return True

    monkeypatch.setattr(video_segmenter, "extract_segment", fake_extract)
    tmp = tmp_path / "temp"
    reps = [batch.ingest_batch(vids, tmp, rank=r, world=2) for r in range(2)]
    assert sorted(v for rep in reps for v in rep.processed) == ["clip0", "clip2"]
    assert [list(rep.failed) for rep in reps if rep.failed] == [["clip1"]]
    merged = batch.merge_progress(tmp, 2)
    assert sorted(merged["processed"]) == ["clip0", "clip2"] and list(merged["failed"]) == ["clip1"]
    assert set(merged) == {"processed", "failed", "last_updated"}          # the reference's ProgressTracker schema
    man = video_segmenter.load_manifest(video_segmenter.get_manifest_path("clip2", tmp))
    assert man["segments"][0]["file_path"].endswith("segment_0000.mp4") and man["segments"][0]["status"] == "pending"
    n_calls = len(calls)
    again = batch.ingest_batch(vids, tmp, rank=0, world=1)                  # resume: processed videos are skipped
    assert sorted(again.processed) == ["clip0", "clip2"] and len(calls) > n_calls   # only the failed one is retried
    assert all("clip1" in c[0] for c in calls[n_calls:])


@pytest.mark.gpu
def test_ingest_batch_on_gpu(cuda, tmp_path):
    vids = [_write_clip(tmp_path / ("g%d.mp4" % i), 320, 240, 20 + 10 * i, 10) for i in range(2)]
    video_segmenter.configure(target_height=120, batch_frames=8)
    try:
        rep = batch.ingest_batch(vids, tmp_path / "temp", rank=0, world=1)
    finally:
        video_segmenter.configure(target_height=720, batch_frames=32)
    assert not rep.failed and rep.segments_done == 2 and rep.pictures == 20 + 30
    seg = tmp_path / "temp" / "segments" / "g1" / "segment_0000.mp4"
    side = json.loads(seg.with_suffix(".json").read_text())
    assert side["frames"] == 30 and side["frame_size"] == [160, 120]
    assert np.fromfile(seg.with_suffix(".frames"), np.uint8).size == 30 * 160 * 120 * 3 // 2


def test_dynamic_queue_claims_each_video_once_across_processes(tmp_path):
    """Two worker processes (one per GPU in production) pull from the same list: every video is ingested exactly once,
    and the merged progress.json has the reference's schema.  frame_buffers off: host logic only."""
    import multiprocessing as mp
    vids = [str(_write_clip(tmp_path / ("q%02d.mp4" % i), 64, 48, 20 + 10 * (i % 3), 10)) for i in range(7)]
    tmp = tmp_path / "temp"
    ctx = mp.get_context("spawn")
    with ctx.Pool(2) as pool:
        reps = pool.starmap(_dynamic_worker, [(vids, str(tmp), r) for r in range(2)])
    done = sorted(v for rep in reps for v in rep["processed"])
    assert done == sorted("q%02d" % i for i in range(7))
    assert all(not rep["failed"] for rep in reps)
    merged = batch.merge_progress(tmp, 2)
    assert sorted(merged["processed"]) == done
    for i in range(7):
        seg = video_segmenter.get_segment_dir("q%02d" % i, tmp) / "segment_0000.mp4"
        assert seg.exists() and container.probe(seg).n_frames == 20 + 10 * (i % 3)


def _dynamic_worker(vids, tmp, rank):
    from video_transformer_b200 import batch as b, video_segmenter as vs
    vs.configure(frame_buffers=False)
    rep = b.ingest_batch_dynamic(vids, tmp, rank=rank, world=2)
    return {"processed": rep.processed, "failed": rep.failed}


def test_scheduler_overlaps_ingest_with_analysis(tmp_path, monkeypatch):
    """process_batch_overlapped: while video i is 'with the remote model' (a stub that sleeps), video i+1 is being
    pre-ingested; the analysis of every video starts only after its own segments are in place."""
    import threading
    import time
    vids = [_write_clip(tmp_path / ("s%d.mp4" % i), 64, 48, 30, 10) for i in range(4)]
    events = []
    lock = threading.Lock()

    def slow_extract(*, input_path, start, end, output_path, stream_copy=True):
        with lock:
            events.append(("ingest_begin", input_path.stem, time.perf_counter()))
        time.sleep(0.15)
        output_path.parent.mkdir(parents=True, exist_ok=True)
        output_path.write_bytes(b"x")
        with lock:
            events.append(("ingest_end", input_path.stem, time.perf_counter()))
        return True

    monkeypatch.setattr(video_segmenter, "extract_segment", slow_extract)

    def analyze(p):
        with lock:
            events.append(("analyze_begin", p.stem, time.perf_counter()))
        time.sleep(0.15)
        seg = video_segmenter.get_segment_dir(p.stem, tmp_path / "temp") / "segment_0000.mp4"
        return seg.exists()

    t0 = time.perf_counter()
    res = batch.process_batch_overlapped(vids, tmp_path / "temp", analyze)
    elapsed = time.perf_counter() - t0
    assert res == [True] * 4
    t = {(k, v): ts for k, v, ts in events}
    for i in range(4):
        assert t[("ingest_end", "s%d" % i)] <= t[("analyze_begin", "s%d" % i)]
    assert t[("ingest_begin", "s1")] < t[("analyze_begin", "s0")] + 0.15       # s1 was being ingested during s0's analysis
    assert elapsed < 4 * 0.30 - 0.2                                              # overlapped: well under the serial time

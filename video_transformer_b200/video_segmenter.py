"""Drop-in for the reference's segmenter module (module-level functions, same names and signatures).

Mirrors /root/reference/src/utils/video_segmenter.py:
    SegmentInfo :12-18, SegmentEntry :21-30, SegmentManifest :33-39, plan_segments :42-83,
    extract_segment :86-154, snap_to_keyframe :157-159, get_segment_dir :162, get_manifest_path :166,
    create_manifest :170-205, load_manifest :208, save_manifest :213-218, load_or_create_manifest :221-238,
    pending_segments :241, update_segment_status :247-266.

Host functions (plan, manifest) are plain Python and must match the reference value-for-value (float64 bit
patterns included; tests/test_host_parity.py replays vectors generated from the reference's own code).
extract_segment keeps its contract -- never raises, returns bool, leaves a non-empty file at output_path --
but instead of spawning ffmpeg it indexes the container, stream-copies the selected samples into a
faststart MP4 and runs the GPU ingest pass (decode -> SAD/histogram -> swscale-exact downscale to
downloader.max_resolution) whose frame buffers and scene scores land beside the MP4:
    <output>.frames   raw planar YUV420P pictures at the target size
    <output>.json     picture range, sizes, per-picture SAD / score, detected cuts
There is no CPU pixel path: without the CUDA library or a GPU the pixel step is reported as failed.
"""
from __future__ import annotations

import json
import logging
import threading
from dataclasses import dataclass
from datetime import datetime, timezone
from pathlib import Path
from typing import TypedDict, cast

import numpy as np

log = logging.getLogger(__name__)


@dataclass(frozen=True)
class SegmentInfo:
    segment_id: int
    start: float
    end: float
    effective_start: float
    effective_end: float


class SegmentEntry(TypedDict):
    id: int
    start: float
    end: float
    effective_start: float
    effective_end: float
    file_path: str
    status: str
    attempts: int
    error: str | None


class SegmentManifest(TypedDict):
    version: int
    video_id: str
    created_at: str
    segment_seconds: float
    overlap_seconds: float
    segments: list[SegmentEntry]


# ---- ingest configuration (set by the host application from its config.yaml) ------------
_OPTIONS = {
    "target_height": 720,      # downloader.max_resolution
    "sws_flags": 4,            # SWS_BICUBIC
    "scene_threshold": 0.10,
    "batch_frames": 64,        # pictures per pipeline batch (32-picture launches run the scaler at half its rate)
    "frame_buffers": True,     # run the GPU pass and write .frames/.json beside the MP4
    "output": "yuv420p",       # or "rgb24" (the upload product: BASELINE.json configs[4])
    "rgb_size": None,          # (width, height) of the RGB output, e.g. (768, 768)
    "sample_every": 1,         # keep pictures whose index is a multiple of this (30 = 1 fps at 30 fps)
    "device": "cuda",
    # Boundary selection (K4), opt-in: > 0 moves a requested start/end to the detected scene cut nearest in time within
    # this many seconds (a score-only GPU pass over the two windows finds the cuts).  0 keeps the reference's times, so
    # plans and manifests stay value-identical by default.
    "snap_tolerance_s": 0.0,
    # write the segment's MP4 through a recycled mapping when the output directory is RAM-backed (landing.acquire_mapped);
    # anywhere else, or when False, the bytes move with copy_file_range
    "mapped_output": True,
}


def configure(**options) -> dict:
    """Update ingest options (unknown keys are rejected).  Returns the active options."""
    for k, v in options.items():
        if k not in _OPTIONS:
            raise KeyError("unknown ingest option %r" % k)
        _OPTIONS[k] = v
    return dict(_OPTIONS)


# ---- time plan ----------------------------------------------------------------------------------------------
def plan_segments(duration: float, segment_seconds: float, overlap_seconds: float) -> list[SegmentInfo]:
    """Tile [0, duration) with cores of segment_seconds; pad each extract window by the overlap.

    The cursor advances by repeated addition of segment_seconds (not id * segment_seconds): the float64
    results must equal the reference's bit for bit.
    """
    plan: list[SegmentInfo] = []
    if duration <= 0 or segment_seconds <= 0:
        return plan
    pad = overlap_seconds if overlap_seconds > 0.0 else 0.0
    lo = 0.0
    while lo < duration:
        hi = min(lo + segment_seconds, duration)
        first = 0.0 if lo == 0 else max(0.0, lo - pad)
        last = duration if hi >= duration else min(duration, hi + pad)
        if last <= first:
            break
        plan.append(SegmentInfo(len(plan), first, last, lo, hi))
        lo = hi
    return plan


def snap_to_keyframe(video_path: str | Path, timestamp: float) -> float:
    """Reference behaviour (a stub there): clamp to >= 0 and ignore the video.  GOP-aligned snapping is
    `keyframe_at_or_before`; it is kept separate so this function stays value-identical to the reference."""
    _ = video_path
    return max(0.0, float(timestamp))


def keyframe_at_or_before(video_path: str | Path, timestamp: float) -> float:
    """Presentation time of the last keyframe at or before `timestamp` (what `-ss T -i IN -c copy` starts at).
    This is the GOP-aligned answer for the reference's stub hook (src/utils/video_segmenter.py:157-159)."""
    t = max(0.0, float(timestamp))
    idx = _probe_cached(Path(video_path))
    if idx is None or not idx.n_frames:
        return t
    times = _picture_times(idx)
    k = _keyframe_index_at_or_before(idx, times, t)
    return float(times[k]) if k is not None else 0.0


def _picture_times(idx) -> np.ndarray:
    """Presentation time of every picture (decode order): k*den/num for plain constant-rate streams (the K4
    definition, one multiply then one divide), the container's own table otherwise."""
    from . import scene
    tt = idx.extra.get("times")
    if tt is not None:
        return np.asarray(tt, np.float64)
    if idx.fps_num <= 0 or idx.fps_den <= 0:          # a corrupt header: no usable time base, nothing gets selected
        return np.full(idx.n_frames, np.inf)
    return scene.pts(np.arange(idx.n_frames), idx.fps_num, idx.fps_den)


def _keyframe_index_at_or_before(idx, times: np.ndarray, t: float):
    ok = np.nonzero(idx.keyframe & (times <= t))[0]
    return int(ok[-1]) if ok.size else None


# ---- cutting ------------------------------------------------------------------------------------------------
def extract_segment(input_path: str | Path, start: float, end: float, output_path: str | Path,
                    stream_copy: bool = True) -> bool:
    duration = end - start
    if duration <= 0:
        return False
    src, dst = Path(input_path), Path(output_path)
    dst.parent.mkdir(parents=True, exist_ok=True)
    # The reference is single-threaded (content_analyzer.py:870, pipeline.py:376); the batch scheduler here ingests
    # video i+1 on a worker thread while the caller works on video i and may cut segments itself.  The engine, index
    # cache and timing record are per process, so calls take turns.
    with _CALL_LOCK:
        try:
            ok = _cut(src, start, end, dst, stream_copy)
        except Exception as exc:  # noqa: BLE001 - same contract as the reference: failures become False
            log.warning("event=segment_cut_failed input=%s start=%.3f end=%.3f error=%s", src, start, end, exc)
            ok = False
    if not ok:
        for p in (dst, _frames_path(dst), _sidecar_path(dst)):
            if p.exists():
                p.unlink()
        return False
    return dst.exists() and dst.stat().st_size > 0


def _frames_path(mp4: Path) -> Path:
    return mp4.with_suffix(".frames")


def _sidecar_path(mp4: Path) -> Path:
    return mp4.with_suffix(".json")


_CALL_LOCK = threading.RLock()
_INDEX_CACHE: dict = {}
_ENGINE_CACHE: dict = {}
LAST_TIMINGS: dict = {}          # seconds spent in the stages of the most recent extract_segment call (diagnostics)


def _file_key(path: Path):
    st = path.stat()
    return (str(path.resolve()), st.st_mtime_ns, st.st_size)


def _probe_cached(path: Path):
    """container.probe with a one-entry-per-file cache: the segments of one long video are cut by successive calls
    (src/analyzer/content_analyzer.py:745-758), and indexing a two-hour file once is enough."""
    from . import container
    try:
        key = _file_key(path)
    except OSError:
        return None
    hit = _INDEX_CACHE.get(key[0])
    if hit is not None and hit[0] == key:
        return hit[1]
    idx = container.probe(path)
    if len(_INDEX_CACHE) > 8:
        _INDEX_CACHE.clear()
    _INDEX_CACHE[key[0]] = (key, idx)
    return idx


def _cut(src: Path, start: float, end: float, dst: Path, stream_copy: bool) -> bool:
    import time
    from . import container, isobmff, scene
    t_begin = time.perf_counter()
    LAST_TIMINGS.clear()
    idx = _probe_cached(src)
    if idx is None or idx.n_frames == 0:
        return False
    times = _picture_times(idx)
    snapped = None
    if _OPTIONS["snap_tolerance_s"] > 0 and _OPTIONS["frame_buffers"]:
        snapped = _snap_window(idx, times, start, end, float(_OPTIONS["snap_tolerance_s"]))
        if snapped is not None:
            start, end = snapped["start"], snapped["end"]
    LAST_TIMINGS["snap"] = time.perf_counter() - t_begin
    if idx.kind == "h264":
        # raw Annex-B elementary stream (the synthetic clips): one slice NAL per picture, wrapped into a new MP4
        if idx.fps_num <= 0:
            return False
        keyframes = np.nonzero(idx.keyframe)[0]
        first, last = scene.frames_for_window(start, end, idx.n_frames, idx.fps_num, idx.fps_den, keyframes,
                                              stream_copy)
        if last <= first:
            return False
        data = np.memmap(src, dtype=np.uint8, mode="r")
        sps, pps = container.find_parameter_sets(data)
        samples = [data[int(o):int(o) + int(z)] for o, z in zip(idx.nal_offsets[first:last], idx.nal_sizes[first:last])]
        keys = [bool(k) for k in idx.keyframe[first:last]]
        if not keys[0]:
            k = _keyframe_index_at_or_before(idx, times, float(times[first]))
            if k is None or not idx.extra.get("pcm_intra_only"):
                return False                 # a mid-GOP start needs an encoder unless pictures merely repeat the IDR
            samples[0] = data[int(idx.nal_offsets[k]):int(idx.nal_offsets[k]) + int(idx.nal_sizes[k])]
            keys[0] = True
        gate = None
        container.write_mp4(dst, sps=sps, pps=pps, samples=samples, width=idx.width, height=idx.height,
                            fps_num=idx.fps_num, fps_den=idx.fps_den, keyframes=keys)
        copier = None
    else:
        movie = idx.extra["movie"]
        track = movie.video_track()
        sel = isobmff.select_reference_range(times, idx.keyframe, start, end, stream_copy)
        if sel is None:
            return False
        first, last, first_acc = sel
        kwargs = {"stream_copy": stream_copy, "selection": sel}
        if not stream_copy and not idx.keyframe[first]:
            # frame-accurate cut that starts inside a GOP (the reference re-encodes with libx264 here,
            # src/utils/video_segmenter.py:138-154; this image has no encoder).  PCM-intra streams let the first
            # picture be re-expressed exactly: it repeats its reference IDR, whose sample becomes the first sample.
            # Every other stream keeps the keyframe lead-in in the file and starts the PRESENTATION at the exact
            # picture through the edit list.
            k = _keyframe_index_at_or_before(idx, times, float(times[first]))
            if k is not None and container.classify_pcm(idx):
                with open(src, "rb") as f:
                    f.seek(int(track.offsets[k]))
                    sample = f.read(int(track.sizes[k]))
                kwargs["first_sample"] = sample
            else:
                first = k if k is not None else first
                kwargs = {"stream_copy": True, "accurate_presentation": True, "selection": (first, last, first_acc)}
        # the stream copy (file -> file, inside the kernel) runs beside the GPU pass
        kwargs["mapped"] = bool(_OPTIONS["mapped_output"])
        gate = _copy_gate()
        copier = _Background(isobmff.cut_movie, movie, start, end, dst, gate=gate, **kwargs)
    LAST_TIMINGS["index"] = time.perf_counter() - t_begin
    try:
        if _OPTIONS["frame_buffers"]:
            reason = None
            if idx.kind == "mp4" and not idx.extra.get("decodable"):
                reason = ("video track %r is not single-slice H.264: the decode front end of this build (K0) handles "
                          "PCM-intra H.264 only and NVDEC is not available on this host" % idx.extra.get("codec"))
            elif idx.kind == "mp4" and not container.classify_pcm(idx):
                reason = ("H.264 stream uses coding tools outside the PCM-intra subset K0 decodes; NVDEC is not "
                          "available on this host")
            if reason is None:
                _ingest_to_files(idx, first, last, dst, snapped, gate.set if gate is not None else None)
            else:                               # the stream copy needs no decode: the cut stands, the pixel pass is skipped
                _write_sidecar(dst, idx, first, last, None, reason)
                log.info("event=segment_pixel_pass_skipped reason=%s", reason)
    finally:
        t_pix = time.perf_counter()
        if gate is not None:
            gate.set()
        res = copier.result() if copier is not None else True
        LAST_TIMINGS["pixel_pass"] = t_pix - t_begin - LAST_TIMINGS["index"]
        LAST_TIMINGS["wait_for_stream_copy"] = time.perf_counter() - t_pix
        if copier is not None:
            LAST_TIMINGS["stream_copy"] = copier.seconds
        LAST_TIMINGS["total"] = time.perf_counter() - t_begin
    return res is not None and res is not False


def _copy_gate():
    """Event the stream-copy thread waits for before it starts (None: start at once).  Its first milliseconds are Python
    (sample tables, moov) and hold the GIL exactly while the calling thread enqueues the first batches of the GPU pass:
    measured with tools/trace_unit.py, the first H2D of a call started 1.4-2.6 ms late and the D2H engine idled that
    long.  The ingest sets the event once its pipeline is full (SegmentIngestor.run(on_primed=...)); from then on the
    calling thread sleeps on CUDA events most of the time."""
    import os
    import threading
    if not _OPTIONS["frame_buffers"] or os.environ.get("VT_COPY_GATE", "1") == "0":
        return None
    return threading.Event()


class _Background:
    """Runs one call on a helper thread; result() re-raises its exception."""

    def __init__(self, fn, *args, gate=None, **kwargs):
        import threading
        self._out = None
        self._exc = None

        self.seconds = 0.0

        def body():
            import time
            if gate is not None:
                gate.wait(0.05)                 # never longer than 50 ms: the copy must not depend on the pixel pass
            t0 = time.perf_counter()
            try:
                self._out = fn(*args, **kwargs)
            except BaseException as e:  # noqa: BLE001
                self._exc = e
            self.seconds = time.perf_counter() - t0

        self._t = threading.Thread(target=body, daemon=True)
        self._t.start()

    def result(self):
        self._t.join()
        if self._exc is not None:
            raise self._exc
        return self._out


def _engine_for(idx, keep_frames: bool = True):
    """SegmentIngestor for (file, options), kept across calls: the segments of one video share plans, pinned staging
    and device buffers.  keep_frames=False is the score-only engine of the boundary selection."""
    from . import ingest
    key = (_file_key(idx.path), tuple(sorted((k, str(v)) for k, v in _OPTIONS.items())))
    slot = "engine" if keep_frames else "engine_scores"
    eng = _ENGINE_CACHE.get(slot)
    if eng is not None and eng[0] == key:
        return eng[1]
    old = _ENGINE_CACHE.pop(slot, None)
    if old is not None:
        old[1].release()                 # its buffers serve the next engine of the same geometry
    opts = ingest.IngestOptions(target_height=_OPTIONS["target_height"], sws_flags=_OPTIONS["sws_flags"],
                                batch_frames=_OPTIONS["batch_frames"], scene_threshold=_OPTIONS["scene_threshold"],
                                output=_OPTIONS["output"], rgb_size=_OPTIONS["rgb_size"],
                                sample_every=_OPTIONS["sample_every"], device=_OPTIONS["device"],
                                keep_frames=keep_frames)
    engine = ingest.SegmentIngestor(idx, opts)
    _ENGINE_CACHE[slot] = (key, engine)
    return engine


def _snap_window(idx, times: np.ndarray, start: float, end: float, tol: float):
    """Boundary selection for one extract window: score the pictures within +-tol of each requested boundary (K3 on the
    GPU, frames not kept), select cuts (K4) and move the boundary to the nearest one.  Boundaries at the very start or
    end of the stream stay.  Returns {"start", "end", "start_snapped", "end_snapped", ...} or None when the stream is
    not decodable here (then nothing moves)."""
    from . import container, scene
    if idx.kind == "mp4" and not (idx.extra.get("decodable") and container.classify_pcm(idx)):
        return None
    if idx.kind == "h264" and not idx.extra.get("pcm_intra_only"):
        return None
    if idx.fps_num <= 0 or idx.extra.get("times") is not None:
        return None                                   # the snap rule is defined on the constant-rate picture grid
    eng = _engine_for(idx, keep_frames=False)
    n = idx.n_frames
    out = {"requested_start": float(start), "requested_end": float(end), "start": float(start), "end": float(end),
           "start_snapped": False, "end_snapped": False, "tolerance_s": tol}
    last_t = float(times[-1])
    for name, t in (("start", float(start)), ("end", float(end))):
        if t <= 0.0 or t > last_t:
            continue
        a = int(np.searchsorted(times, t - tol, side="left"))
        b = int(np.searchsorted(times, t + tol, side="right"))
        a, b = max(a, 0), min(b, n)
        if b <= a:
            continue
        res = eng.run(a, b, None)
        rec = scene.snap_boundaries([t], res.cuts, n, idx.fps_num, idx.fps_den, tol)[0]
        if rec["snapped"]:
            out[name] = rec["time"]
            out[name + "_snapped"] = True
            out[name + "_frame"] = rec["frame"]
    if out["end"] <= out["start"]:
        return None
    return out


def _ingest_to_files(idx, first: int, last: int, dst: Path, snapped=None, on_primed=None) -> None:
    """GPU pass for pictures [first,last): writes <dst>.frames and <dst>.json.  Raises on any failure."""
    import time
    from . import landing
    t0 = time.perf_counter()
    eng = _engine_for(idx)
    n_out = eng.kept_pictures(first, last)
    t1 = time.perf_counter()
    land = landing.acquire(_frames_path(dst), max(n_out, 1) * eng.frame_bytes)
    t2 = time.perf_counter()
    LAST_TIMINGS["engine"] = t1 - t0
    LAST_TIMINGS["landing_acquire"] = t2 - t1
    try:
        res = eng.run(first, last, landing=land, on_primed=on_primed)
        land.finish(res.stats["landed_frames"] * eng.frame_bytes)
        LAST_TIMINGS["gpu_pass"] = time.perf_counter() - t2
    except BaseException:
        land.abort()
        raise
    _write_sidecar(dst, idx, first, last, res, None, eng.opts, land, snapped)
    log.info("event=segment_ingest frames=%d size=%dx%d cuts=%d landing=%s", res.stats["landed_frames"],
             res.out_width, res.out_height, len(res.cuts), res.stats["landing"])


def _write_sidecar(dst: Path, idx, first: int, last: int, res, reason, opts=None, land=None, snapped=None) -> None:
    side = {"source": str(idx.path), "first_picture": first, "last_picture": last,
            "fps": [idx.fps_num, idx.fps_den], "source_size": [idx.width, idx.height],
            "codec": idx.extra.get("codec", "h264")}
    if res is None:
        side.update({"frames": None, "reason": reason})
    else:
        side.update({
            "frame_size": [res.out_width, res.out_height], "pixel_format": opts.output,
            "sample_every": opts.sample_every, "frame_bytes": res.frame_bytes,
            "frames": res.stats["landed_frames"],
            "landing": res.stats["landing"], "landing_recycled": bool(land is not None and land.recycled),
            "scene_threshold": opts.scene_threshold, "cuts": res.cuts.tolist(),
            "sad": res.sad.tolist(), "score": res.scores.tolist(),
        })
    if snapped is not None:
        side["boundaries"] = snapped
    _sidecar_path(dst).write_text(json.dumps(side), encoding="utf-8")


# ---- manifest lifecycle -------------------------------------------------------------------------------------
def get_segment_dir(video_id: str, temp_dir: str | Path) -> Path:
    return Path(temp_dir) / "segments" / video_id


def get_manifest_path(video_id: str, temp_dir: str | Path) -> Path:
    return get_segment_dir(video_id, temp_dir) / "manifest.json"


def create_manifest(*, video_id: str, duration: float, segment_seconds: float, overlap_seconds: float,
                    temp_dir: str | Path) -> SegmentManifest:
    folder = get_segment_dir(video_id, temp_dir)
    folder.mkdir(parents=True, exist_ok=True)
    entries: list[SegmentEntry] = []
    for seg in plan_segments(duration, segment_seconds, overlap_seconds):
        entries.append({
            "id": seg.segment_id,
            "start": seg.start,
            "end": seg.end,
            "effective_start": seg.effective_start,
            "effective_end": seg.effective_end,
            "file_path": str(folder / ("segment_%04d.mp4" % seg.segment_id)),
            "status": "pending",
            "attempts": 0,
            "error": None,
        })
    manifest: SegmentManifest = {
        "version": 1,
        "video_id": video_id,
        "created_at": datetime.now(timezone.utc).isoformat(),
        "segment_seconds": segment_seconds,
        "overlap_seconds": overlap_seconds,
        "segments": entries,
    }
    save_manifest(get_manifest_path(video_id, temp_dir), manifest)
    return manifest


def load_manifest(manifest_path: str | Path) -> SegmentManifest:
    return cast(SegmentManifest, json.loads(Path(manifest_path).read_text(encoding="utf-8")))


def save_manifest(manifest_path: str | Path, manifest: SegmentManifest) -> None:
    target = Path(manifest_path)
    target.parent.mkdir(parents=True, exist_ok=True)
    target.write_text(json.dumps(manifest, indent=2, ensure_ascii=True), encoding="utf-8")


def load_or_create_manifest(*, video_id: str, duration: float, segment_seconds: float, overlap_seconds: float,
                            temp_dir: str | Path) -> SegmentManifest:
    existing = get_manifest_path(video_id, temp_dir)
    if existing.exists():  # resume: the file on disk wins, even if the plan parameters changed
        return load_manifest(existing)
    return create_manifest(video_id=video_id, duration=duration, segment_seconds=segment_seconds,
                           overlap_seconds=overlap_seconds, temp_dir=temp_dir)


def pending_segments(manifest: SegmentManifest) -> list[SegmentEntry]:
    return [entry for entry in manifest["segments"] if entry["status"] != "completed"]


def update_segment_status(manifest: SegmentManifest, segment_id: int, status: str, *, error: str | None = None,
                          increment_attempts: bool = False) -> None:
    match = next((entry for entry in manifest["segments"] if entry["id"] == segment_id), None)
    if match is None:
        log.warning("Segment id %s not found in manifest", segment_id)
        return
    match["status"] = status
    if error is not None:
        match["error"] = error
    if increment_attempts:
        match["attempts"] = match["attempts"] + 1

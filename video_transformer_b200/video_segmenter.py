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
    "batch_frames": 32,
    "frame_buffers": True,     # run the GPU pass and write .frames/.json beside the MP4
    "output": "yuv420p",       # or "rgb24" (the upload product: BASELINE.json configs[4])
    "rgb_size": None,          # (width, height) of the RGB output, e.g. (768, 768)
    "sample_every": 1,         # keep pictures whose index is a multiple of this (30 = 1 fps at 30 fps)
    "device": "cuda",
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
    """Presentation time of the last keyframe at or before `timestamp` (what `-ss T -i IN -c copy` starts at)."""
    from . import container
    t = max(0.0, float(timestamp))
    idx = container.probe(Path(video_path))
    if idx is None or not idx.n_frames:
        return t
    k = np.nonzero(idx.keyframe)[0]
    times = k.astype(np.float64) * float(idx.fps_den) / float(idx.fps_num)
    ok = times[times <= t]
    return float(ok[-1]) if ok.size else 0.0


# ---- cutting ------------------------------------------------------------------------------------------------
def extract_segment(input_path: str | Path, start: float, end: float, output_path: str | Path,
                    stream_copy: bool = True) -> bool:
    duration = end - start
    if duration <= 0:
        return False
    src, dst = Path(input_path), Path(output_path)
    dst.parent.mkdir(parents=True, exist_ok=True)
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


def _cut(src: Path, start: float, end: float, dst: Path, stream_copy: bool) -> bool:
    from . import container, scene
    idx = container.probe(src)
    if idx is None or idx.n_frames == 0 or idx.fps_num <= 0:
        return False
    keyframes = np.nonzero(idx.keyframe)[0]
    first, last = scene.frames_for_window(start, end, idx.n_frames, idx.fps_num, idx.fps_den, keyframes, stream_copy)
    if last <= first:
        return False
    data = np.memmap(src, dtype=np.uint8, mode="r")
    if idx.kind == "h264":
        sps, pps = container.find_parameter_sets(data)
    else:
        sps, pps = idx.sps, idx.pps
    samples = [bytes(data[int(o):int(o) + int(s)]) for o, s in
               zip(idx.nal_offsets[first:last], idx.nal_sizes[first:last])]
    keys = [bool(k) for k in idx.keyframe[first:last]]
    if not keys[0]:
        # frame-accurate cut that starts inside a GOP (the reference re-encodes here).  The PCM-intra subset
        # lets us re-express the first picture exactly: it repeats its reference IDR, so that IDR's samples
        # become the first sample.  Anything else would need an encoder.
        ref = keyframes[keyframes <= first]
        if idx.extra.get("pcm_intra_only") is False or ref.size == 0:
            return False
        k = int(ref[-1])
        samples[0] = bytes(data[int(idx.nal_offsets[k]):int(idx.nal_offsets[k]) + int(idx.nal_sizes[k])])
        keys[0] = True
    container.write_mp4(dst, sps=sps, pps=pps, samples=samples, width=idx.width, height=idx.height,
                        fps_num=idx.fps_num, fps_den=idx.fps_den, keyframes=keys)
    if _OPTIONS["frame_buffers"]:
        _ingest_to_files(idx, data, first, last, dst)
    return True


def _ingest_to_files(idx, data, first: int, last: int, dst: Path) -> None:
    """GPU pass for pictures [first,last): writes <dst>.frames and <dst>.json.  Raises on any failure."""
    from . import ingest
    opts = ingest.IngestOptions(target_height=_OPTIONS["target_height"], sws_flags=_OPTIONS["sws_flags"],
                                batch_frames=_OPTIONS["batch_frames"], scene_threshold=_OPTIONS["scene_threshold"],
                                output=_OPTIONS["output"], rgb_size=_OPTIONS["rgb_size"],
                                sample_every=_OPTIONS["sample_every"], device=_OPTIONS["device"])
    eng = ingest.SegmentIngestor(idx, opts, host_bytes=data)
    sink = ingest.FileSink(_frames_path(dst))
    try:
        res = eng.run(first, last, sink)
    finally:
        sink.close()
    side = {
        "source": str(idx.path), "first_picture": first, "last_picture": last, "fps": [idx.fps_num, idx.fps_den],
        "source_size": [idx.width, idx.height], "frame_size": [res.out_width, res.out_height],
        "pixel_format": opts.output, "sample_every": opts.sample_every, "frame_bytes": res.frame_bytes,
        "frames": sink.frames,
        "scene_threshold": opts.scene_threshold, "cuts": [int(c) for c in res.cuts],
        "sad": [int(s) for s in res.sad], "score": [float(s) for s in res.scores],
    }
    _sidecar_path(dst).write_text(json.dumps(side), encoding="utf-8")
    log.info("event=segment_ingest frames=%d size=%dx%d cuts=%d", sink.frames, res.out_width, res.out_height,
             len(res.cuts))


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

"""Batch pre-ingest across the GPUs of one box (BASELINE.json configs[3]).

The reference processes a URL list strictly one video at a time (/root/reference/src/pipeline.py:361-396) and keeps a
video-level checkpoint in `progress.json` (/root/reference/src/utils/progress_tracker.py:36-132: `processed` list,
`failed` map, `last_updated`).  Here the local part of that loop -- probe, plan, manifest, cut + GPU pass for every
pending segment -- is sharded per VIDEO over the ranks (one process per GPU, no collective on the data path):

    rank r of W   ingests   shard.assign_videos(n_pictures, W)[r]      (longest-first greedy, identical on every rank)

Each rank writes its own `progress.rank<r>.json` with the reference's schema; `merge_progress` folds them into the
`progress.json` the reference's ProgressTracker reads, so a later run of the unmodified pipeline finds every segment
file already present (content_analyzer.py:749 skips extraction when the file exists) and every video's manifest in
place.  Segment planning and manifests go through the same functions the analyzer calls.
"""
from __future__ import annotations

import json
import logging
import os
from dataclasses import dataclass, field
from datetime import datetime
from pathlib import Path

from . import budget_planner, container, shard, video_segmenter
from .video_utils import probe_duration

log = logging.getLogger(__name__)


@dataclass
class BatchReport:
    rank: int
    world: int
    assigned: list[str] = field(default_factory=list)
    processed: list[str] = field(default_factory=list)
    failed: dict = field(default_factory=dict)
    segments_done: int = 0
    pictures: int = 0


def _progress_path(temp_dir: Path, rank: int | None) -> Path:
    return temp_dir / ("progress.json" if rank is None else "progress.rank%d.json" % rank)


def _load_progress(path: Path) -> dict:
    if path.exists():
        try:
            return json.loads(path.read_text(encoding="utf-8"))
        except (OSError, ValueError):
            pass
    return {"processed": [], "failed": {}, "last_updated": None}


def _save_progress(path: Path, data: dict) -> None:
    data["last_updated"] = datetime.now().isoformat()
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(json.dumps(data, ensure_ascii=False, indent=2), encoding="utf-8")


def plan_batch(video_paths, world: int) -> list[list[int]]:
    """Which videos each rank ingests.  Deterministic, so every rank computes the same answer without talking."""
    counts = []
    for p in video_paths:
        idx = container.probe(Path(p))
        counts.append(int(idx.n_frames) if idx is not None else 0)
    return shard.assign_videos(counts, world)


def ingest_video(video_path, temp_dir, config: dict | None = None, current_api_count: int = 0) -> tuple[int, int]:
    """Probe -> budget plan -> manifest -> extract every pending segment.  Returns (segments done, pictures).
    Mirrors the local steps of ContentAnalyzer._analyze_video_segments (content_analyzer.py:822-942)."""
    video_path, temp_dir = Path(video_path), Path(temp_dir)
    duration = probe_duration(video_path)
    if duration <= 0:
        raise RuntimeError("probe_duration returned 0 for %s" % video_path)
    plan = budget_planner.plan_segments_with_budget(duration, config or {}, current_api_count)
    if plan.num_segments <= 0 or plan.segment_duration <= 0:
        raise RuntimeError("no segment plan fits the API budget for %s" % video_path)
    manifest = video_segmenter.load_or_create_manifest(video_id=video_path.stem, duration=duration,
                                                       segment_seconds=plan.segment_duration,
                                                       overlap_seconds=plan.overlap, temp_dir=temp_dir)
    mpath = video_segmenter.get_manifest_path(video_path.stem, temp_dir)
    done = pictures = 0
    for entry in sorted(manifest["segments"], key=lambda e: e["id"]):
        seg = Path(entry["file_path"])
        if not (seg.exists() and seg.stat().st_size > 0):        # same reuse rule as content_analyzer.py:749
            ok = video_segmenter.extract_segment(input_path=video_path, start=entry["start"], end=entry["end"],
                                                 output_path=seg, stream_copy=True)
            if not ok:
                video_segmenter.update_segment_status(manifest, entry["id"], "failed", error="extract_segment failed",
                                                      increment_attempts=True)
                video_segmenter.save_manifest(mpath, manifest)
                raise RuntimeError("extract_segment failed for %s segment %d" % (video_path, entry["id"]))
        side = seg.with_suffix(".json")
        if side.exists():
            info = json.loads(side.read_text(encoding="utf-8"))
            pictures += int(info.get("frames") or 0)
            b = info.get("boundaries")
            if b and (b.get("start_snapped") or b.get("end_snapped")):
                # boundary selection moved this segment onto detected scene cuts (opt-in, ingest option
                # snap_tolerance_s): recorded as extra keys, the reference's own keys keep their planned values
                entry["snapped_start"], entry["snapped_end"] = b["start"], b["end"]
        done += 1
    # statuses stay "pending": the analyzer owns the status machine (processing/completed) when it runs later
    video_segmenter.save_manifest(mpath, manifest)
    return done, pictures


def ingest_batch(video_paths, temp_dir, *, rank: int | None = None, world: int | None = None,
                 config: dict | None = None) -> BatchReport:
    """Ingest this rank's share of `video_paths`.  rank/world default to RANK/WORLD_SIZE (torchrun) or 0/1."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    temp_dir = Path(temp_dir)
    paths = [str(p) for p in video_paths]
    mine = plan_batch(paths, world)[rank]
    rep = BatchReport(rank=rank, world=world, assigned=[paths[i] for i in mine])
    ppath = _progress_path(temp_dir, rank)
    prog = _load_progress(ppath)
    for i in mine:
        vid = Path(paths[i]).stem
        if vid in prog["processed"]:
            rep.processed.append(vid)
            continue
        try:
            done, pics = ingest_video(paths[i], temp_dir, config)
            rep.segments_done += done
            rep.pictures += pics
            rep.processed.append(vid)
            prog["processed"].append(vid)
            prog["failed"].pop(vid, None)
        except Exception as exc:  # noqa: BLE001 - one bad video must not stop the batch (pipeline.py:340-359)
            log.warning("event=batch_ingest_failed video=%s error=%s", vid, exc)
            rep.failed[vid] = str(exc)
            prog["failed"][vid] = {"error": str(exc), "timestamp": datetime.now().isoformat()}
        _save_progress(ppath, prog)
    _save_progress(ppath, prog)
    return rep


# ---- dynamic per-video queue and the overlapped batch loop ----------------------------------------------------------
def _claim(temp_dir: Path, video_id: str, rank: int) -> bool:
    """Atomically claim a video for this rank (O_EXCL create of temp_dir/claims/<video_id>): the shared queue of the
    N worker processes is the file system, no collective and no server."""
    d = temp_dir / "claims"
    d.mkdir(parents=True, exist_ok=True)
    try:
        fd = os.open(d / (video_id + ".claim"), os.O_CREAT | os.O_EXCL | os.O_WRONLY, 0o644)
    except FileExistsError:
        return False
    os.write(fd, str(rank).encode())
    os.close(fd)
    return True


def ingest_batch_dynamic(video_paths, temp_dir, *, rank: int | None = None, world: int | None = None,
                         config: dict | None = None, on_done=None) -> BatchReport:
    """Like ingest_batch, but ranks PULL videos (longest first) instead of owning a static share: a rank behind a slow
    host link or with longer videos simply claims fewer.  Every rank walks the same order and claims atomically.
    on_done(video_id, ok) is called after each video this rank ingested (the scheduler's hand-off)."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    temp_dir = Path(temp_dir)
    paths = [str(p) for p in video_paths]
    counts = []
    for p in paths:
        idx = container.probe(Path(p))
        counts.append(int(idx.n_frames) if idx is not None else 0)
    order = sorted(range(len(paths)), key=lambda k: (-counts[k], k))
    rep = BatchReport(rank=rank, world=world)
    ppath = _progress_path(temp_dir, rank)
    prog = _load_progress(ppath)
    for i in order:
        vid = Path(paths[i]).stem
        if not _claim(temp_dir, vid, rank):
            continue
        rep.assigned.append(paths[i])
        ok = True
        try:
            done, pics = ingest_video(paths[i], temp_dir, config)
            rep.segments_done += done
            rep.pictures += pics
            rep.processed.append(vid)
            prog["processed"].append(vid)
            prog["failed"].pop(vid, None)
        except Exception as exc:  # noqa: BLE001 - one bad video must not stop the batch (pipeline.py:340-359)
            ok = False
            log.warning("event=batch_ingest_failed video=%s error=%s", vid, exc)
            rep.failed[vid] = str(exc)
            prog["failed"][vid] = {"error": str(exc), "timestamp": datetime.now().isoformat()}
        _save_progress(ppath, prog)
        if on_done is not None:
            on_done(vid, ok)
    _save_progress(ppath, prog)
    return rep


class IngestScheduler:
    """Pre-ingests a list of videos on a worker thread while the caller analyses earlier ones.

    The reference's batch loop (/root/reference/src/pipeline.py:361-396) handles one video at a time: probe, cut, upload,
    wait for the remote model, next video -- the GPU would idle during every remote call.  Here the local half (probe,
    plan, manifest, cut + GPU pass of every segment) of video i+1 runs while video i is with the remote model; when the
    unmodified analyzer reaches video i+1 it finds every segment file present and skips extraction
    (/root/reference/src/analyzer/content_analyzer.py:749).  Several processes (one per GPU) may run a scheduler over
    the same list: videos are claimed atomically, whoever is free takes the next one."""

    def __init__(self, video_paths, temp_dir, *, rank: int | None = None, world: int | None = None,
                 config: dict | None = None):
        import threading
        self.paths = [str(p) for p in video_paths]
        self.temp_dir = Path(temp_dir)
        self.rank, self.world, self.config = rank, world, config
        self._done: dict[str, bool] = {}
        self._cv = threading.Condition()
        self._finished = False
        self.report: BatchReport | None = None
        self._thread = threading.Thread(target=self._work, daemon=True)

    def start(self) -> "IngestScheduler":
        self._thread.start()
        return self

    def _work(self) -> None:
        def on_done(vid, ok):
            with self._cv:
                self._done[vid] = ok
                self._cv.notify_all()
        try:
            self.report = ingest_batch_dynamic(self.paths, self.temp_dir, rank=self.rank, world=self.world,
                                               config=self.config, on_done=on_done)
        finally:
            with self._cv:
                self._finished = True
                self._cv.notify_all()

    def wait(self, video_path, timeout: float | None = None) -> bool:
        """Block until this video's pre-ingest has finished here.  True when its segments are in place; False when it
        failed, was claimed by another process, or the timeout expired (the analyzer then cuts on demand, as the
        reference does)."""
        vid = Path(video_path).stem
        with self._cv:
            self._cv.wait_for(lambda: vid in self._done or self._finished, timeout)
            return bool(self._done.get(vid, False))

    def join(self) -> BatchReport | None:
        self._thread.join()
        return self.report


def process_batch_overlapped(video_paths, temp_dir, analyze, *, config: dict | None = None,
                             rank: int | None = None, world: int | None = None) -> list:
    """The reference's batch loop with the ingest of later videos overlapped: for each video in order, wait for its
    pre-ingest, then call analyze(video_path) (in the reference: VideoPipeline.process_single_video -> the analyzer's
    segment loop -> remote model).  Returns analyze's results in order."""
    sched = IngestScheduler(video_paths, temp_dir, rank=rank, world=world, config=config).start()
    out = []
    for p in video_paths:
        sched.wait(p)
        out.append(analyze(p))
    sched.join()
    return out


def merge_progress(temp_dir, world: int) -> dict:
    """Fold progress.rank*.json into the progress.json the reference's ProgressTracker loads."""
    temp_dir = Path(temp_dir)
    merged = _load_progress(_progress_path(temp_dir, None))
    for r in range(world):
        part = _load_progress(_progress_path(temp_dir, r))
        for v in part["processed"]:
            if v not in merged["processed"]:
                merged["processed"].append(v)
            merged["failed"].pop(v, None)
        for v, why in part["failed"].items():
            if v not in merged["processed"]:
                merged["failed"][v] = why
    _save_progress(_progress_path(temp_dir, None), merged)
    return merged


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="Pre-ingest a list of local videos on this rank's GPU")
    ap.add_argument("--list", required=True, help="text file, one video path per line (URL.txt style)")
    ap.add_argument("--temp-dir", required=True)
    ap.add_argument("--target-height", type=int, default=720)
    a = ap.parse_args(argv)
    paths = [ln.strip() for ln in Path(a.list).read_text(encoding="utf-8").splitlines() if ln.strip()]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    video_segmenter.configure(target_height=a.target_height, device="cuda:%d" % local)
    rep = ingest_batch(paths, a.temp_dir)
    print(json.dumps({"rank": rep.rank, "world": rep.world, "videos": len(rep.assigned), "processed": len(rep.processed),
                      "failed": rep.failed, "segments": rep.segments_done, "pictures": rep.pictures}))
    return 0 if not rep.failed else 1


if __name__ == "__main__":
    raise SystemExit(main())

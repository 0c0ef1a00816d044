"""K4: scene scores, cut selection and time -> picture index mapping.  Host float64 from the GPU's integer SADs.

Definitions (nothing in the reference computes these, so the in-repo oracle
oracle/scene_oracle.py is the authority and tests compare index-for-index):
  mafd_t  = sad_t / (W*H)                       (float64, one divide)
  diff_t  = |mafd_t - mafd_{t-1}|               (mafd_{-1} = 0; picture 0 has sad = 0)
  score_t = min(max(min(mafd_t, diff_t) / 100, 0), 1)
  cut at t  iff  t > 0 and score_t > threshold  (strict)
Time plan -> pictures follows what the reference's ffmpeg command line selects
(/root/reference/src/utils/video_segmenter.py:118-137): start and duration are rounded to milliseconds by the
`.3f` formatting, `-ss` before `-i` seeks on the input, and with `-c copy` output begins at the keyframe at or
before the seek point.
"""
from __future__ import annotations

import numpy as np


def scene_scores(sad: np.ndarray, width: int, height: int) -> np.ndarray:
    """Per-picture score in [0, 1] from per-picture SAD (uint64).  Sequential by definition (diff of mafd)."""
    sad = np.asarray(sad, dtype=np.uint64)
    area = float(width * height)
    mafd = sad.astype(np.float64) / area
    prev = np.concatenate(([0.0], mafd[:-1]))
    diff = np.abs(mafd - prev)
    score = np.minimum(np.maximum(np.minimum(mafd, diff) / 100.0, 0.0), 1.0)
    if score.size:
        score[0] = 0.0
    return score


def select_cuts(score: np.ndarray, threshold: float, first_picture: int = 0) -> np.ndarray:
    """Absolute picture indices t > 0 with score_t > threshold; score[0] belongs to picture `first_picture`."""
    idx = np.nonzero(np.asarray(score) > threshold)[0].astype(np.int64) + int(first_picture)
    return idx[idx > 0]


def pts(k, fps_num: int, fps_den: int):
    """Presentation time of picture k for constant frame rate: k*den/num (multiply, then divide)."""
    return np.asarray(k, dtype=np.float64) * float(fps_den) / float(fps_num)


def frames_for_window(start: float, end: float, n_frames: int, fps_num: int, fps_den: int,
                      keyframes: np.ndarray | None = None, stream_copy: bool = False) -> tuple[int, int]:
    """Half-open picture range [first, last) that `ffmpeg -ss start -i IN -t (end-start)` keeps.

    accurate seek : { k : s <= pts_k < s + d }, s = round(start, 3), d = round(end - start, 3)
    stream copy   : the same range, extended back to the last keyframe at or before its first picture.
    """
    s = float("%.3f" % start)
    d = float("%.3f" % (end - start))
    if d <= 0 or n_frames <= 0:
        return 0, 0
    t = pts(np.arange(n_frames), fps_num, fps_den)
    keep = np.nonzero((t >= s) & (t < s + d))[0]
    if keep.size == 0:
        return 0, 0
    first, last = int(keep[0]), int(keep[-1]) + 1
    if stream_copy and keyframes is not None:
        kf = np.asarray(keyframes)
        kf = kf[kf <= first]
        if kf.size:
            first = int(kf[-1])
    return first, last


def boundary_frame(t: float, n_frames: int, fps_num: int, fps_den: int) -> int:
    """First picture whose pts is >= t (n_frames when t is past the end)."""
    tt = pts(np.arange(n_frames), fps_num, fps_den)
    return int(np.searchsorted(tt, t, side="left"))


def snap_boundaries(boundaries_s, cuts: np.ndarray, n_frames: int, fps_num: int, fps_den: int,
                    tolerance_s: float) -> list[dict]:
    """Move each planned boundary (seconds) to the nearest detected cut within +-tolerance_s.

    Ties go to the earlier cut.  A boundary with no cut in range keeps its time-plan picture.  Returns one
    record per boundary: planned time, planned picture, chosen picture, chosen time, whether it snapped.
    """
    cuts = np.asarray(cuts, dtype=np.int64)
    out = []
    for t in boundaries_s:
        k0 = boundary_frame(float(t), n_frames, fps_num, fps_den)
        best, snapped = k0, False
        if cuts.size:
            ct = pts(cuts, fps_num, fps_den)
            dist = np.abs(ct - float(t))
            ok = np.nonzero(dist <= tolerance_s)[0]
            if ok.size:
                j = ok[np.argmin(dist[ok])]     # argmin returns the first (earliest) minimum
                best, snapped = int(cuts[j]), True
        out.append({"planned_time": float(t), "planned_frame": k0, "frame": best,
                    "time": float(pts(best, fps_num, fps_den)), "snapped": snapped})
    return out

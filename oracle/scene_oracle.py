"""Oracle (O) for K4: plain-Python restatement of scene scoring / cut selection / picture ranges.

TEST INFRASTRUCTURE ONLY.  Parity unpinned by the reference (it has no content-based detection:
/root/reference/src/utils/video_segmenter.py:157-159 is a stub); the definition is SURVEY.md section 8a K4,
restated here with scalar loops so that it shares no code with video_transformer_b200/scene.py.
"""
from __future__ import annotations


def scene_scores(sads, width, height):
    out = []
    prev_mafd = 0.0
    for t, s in enumerate(sads):
        mafd = float(int(s)) / float(width * height)
        diff = abs(mafd - prev_mafd)
        sc = min(mafd, diff) / 100.0
        sc = 0.0 if sc < 0.0 else (1.0 if sc > 1.0 else sc)
        out.append(0.0 if t == 0 else sc)
        prev_mafd = mafd
    return out


def select_cuts(scores, thr):
    return [t for t, s in enumerate(scores) if t > 0 and s > thr]


def frames_for_window(start, end, n_frames, fps_num, fps_den, keyframes=None, stream_copy=False):
    s = float(format(start, ".3f"))
    d = float(format(end - start, ".3f"))
    if d <= 0:
        return (0, 0)
    kept = [k for k in range(n_frames) if s <= (k * float(fps_den)) / float(fps_num) < s + d]
    if not kept:
        return (0, 0)
    first, last = kept[0], kept[-1] + 1
    if stream_copy and keyframes is not None:
        earlier = [int(k) for k in keyframes if int(k) <= first]
        if earlier:
            first = max(earlier)
    return (first, last)


def boundary_frame(t, n_frames, fps_num, fps_den):
    """First picture whose presentation time is >= t (n_frames when t lies past the last picture)."""
    for k in range(n_frames):
        if (k * float(fps_den)) / float(fps_num) >= t:
            return k
    return n_frames


def snap_boundaries(boundaries_s, cuts, n_frames, fps_num, fps_den, tolerance_s):
    """Boundary selection (the hook the reference leaves as a stub, src/utils/video_segmenter.py:157-159): every
    planned boundary moves to the detected cut nearest in time within +-tolerance_s; on a tie the earlier cut wins;
    without a cut in range the boundary keeps its time-plan picture."""
    out = []
    for t in boundaries_s:
        t = float(t)
        planned = boundary_frame(t, n_frames, fps_num, fps_den)
        best, best_d, snapped = planned, None, False
        for c in cuts:
            ct = (int(c) * float(fps_den)) / float(fps_num)
            d = abs(ct - t)
            if d <= tolerance_s and (best_d is None or d < best_d):
                best, best_d, snapped = int(c), d, True
        out.append({"planned_time": t, "planned_frame": planned, "frame": best,
                    "time": (best * float(fps_den)) / float(fps_num), "snapped": snapped})
    return out

"""Multi-GPU sharding of the ingest path: independent units, no exchange step.

* per video   -- a URL.txt-style batch (the reference loops sequentially, /root/reference/src/pipeline.py:376-393):
                 longest-first greedy assignment by picture count, one process per GPU;
* per segment -- one long video: contiguous, GOP-aligned picture ranges, one per rank.  A shard recomputes the
                 only cross-shard state (previous luma / previous mafd) from 2 lead-in pictures
                 (ingest.SegmentIngestor.run), so nothing is communicated on the data path.
The host then concatenates per-picture SADs (8 bytes each) in picture order and runs K4 once, which makes the
boundaries independent of the GPU count.
"""
from __future__ import annotations

import numpy as np

from . import scene


def assign_videos(n_frames: list[int], world: int) -> list[list[int]]:
    """Longest-first greedy: returns, per rank, the indices of the videos it ingests."""
    loads = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(n_frames)), key=lambda k: (-n_frames[k], k)):
        r = min(range(world), key=lambda q: (loads[q], q))
        out[r].append(i)
        loads[r] += n_frames[i]
    return [sorted(v) for v in out]


def split_gop_aligned(keyframes, first: int, last: int, world: int) -> list[tuple[int, int]]:
    """Cut [first,last) into `world` contiguous ranges whose starts are keyframes (except the first).

    Ranges may be empty when there are fewer GOPs than ranks.  Balanced by picture count."""
    kf = np.asarray([k for k in keyframes if first < k < last], dtype=np.int64)
    cuts = [first]
    for r in range(1, world):
        target = first + (last - first) * r // world
        if kf.size:
            j = int(np.argmin(np.abs(kf - target)))
            c = int(kf[j])
        else:
            c = cuts[-1]
        cuts.append(max(c, cuts[-1]))
    cuts.append(last)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def merge_and_score(parts: list[tuple[int, np.ndarray]], width: int, height: int, threshold: float):
    """parts: (first_picture, sad[uint64]) per shard, any order.  Returns (sad, scores, cuts) for the whole range.

    Each shard's sad[0] is SAD(first, first-1) computed from its lead-in, so concatenation equals the single-pass
    array; scoring once here is what makes cuts identical at 1/2/4/8 GPUs."""
    parts = sorted(((int(a), np.asarray(s, dtype=np.uint64)) for a, s in parts if len(s)), key=lambda t: t[0])
    pos = parts[0][0]
    for a, s in parts:
        if a != pos:
            raise ValueError("shards are not contiguous at picture %d" % a)
        pos += len(s)
    sad = np.concatenate([s for _, s in parts])
    if parts[0][0] != 0:
        raise ValueError("merge_and_score expects the range to start at picture 0")
    sad[0] = 0
    scores = scene.scene_scores(sad, width, height)
    return sad, scores, scene.select_cuts(scores, threshold)

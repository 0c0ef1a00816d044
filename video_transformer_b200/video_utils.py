"""probe_duration without ffprobe.

Drop-in for /root/reference/src/utils/video_utils.py:7-38 (`ffprobe ... format=duration`): returns the
container duration in seconds as a float, and 0.0 on ANY failure (missing file, unknown container, parse
error) -- the caller treats 0.0 as "do not segment" (/root/reference/src/analyzer/content_analyzer.py:497-498).
Containers understood (codec-agnostic: the duration lives in the container, not in the samples): ISO-BMFF
(MP4/MOV/M4A: mvhd duration/timescale rescaled to microseconds, what ffprobe reports for format=duration; fragmented
files through mehd or their fragments), Matroska/WebM (Segment Info Duration), AVI (stream headers), and raw Annex-B
H.264 elementary streams (picture count / VUI frame rate; ffprobe itself prints N/A there).
"""
from __future__ import annotations

from pathlib import Path

from . import container


def probe_duration(video_path: str | Path) -> float:
    try:
        d = float(container.container_duration(Path(video_path)))
    except Exception:  # noqa: BLE001 - the reference swallows every failure into 0.0
        return 0.0
    return d if d > 0 else 0.0

"""probe_duration without ffprobe.

Drop-in for /root/reference/src/utils/video_utils.py:7-38 (`ffprobe ... format=duration`): returns the
container duration in seconds as a float, and 0.0 on ANY failure (missing file, unknown container, parse
error) -- the caller treats 0.0 as "do not segment" (/root/reference/src/analyzer/content_analyzer.py:497-498).
Containers understood: ISO-BMFF/MP4 (mvhd duration/timescale, what ffprobe reports for format=duration) and
raw Annex-B H.264 elementary streams (picture count / VUI frame rate; ffprobe itself prints N/A there).
"""
from __future__ import annotations

from pathlib import Path

from . import container


def probe_duration(video_path: str | Path) -> float:
    try:
        info = container.probe(Path(video_path))
    except Exception:  # noqa: BLE001 - the reference swallows every failure into 0.0
        return 0.0
    if info is None:
        return 0.0
    d = float(info.duration)
    return d if d > 0 else 0.0

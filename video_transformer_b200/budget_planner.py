"""Segment-length planning under an API-call budget.

Drop-in for /root/reference/src/utils/budget_planner.py (SegmentPlan :9-17, _estimate_segments :43-53,
_estimate_calls :56-70, plan_segments_with_budget :73-194): same names, same arguments, same results --
tests/test_host_parity.py replays golden vectors generated from the reference's own code.  Pure integer and
float64 host arithmetic; it fixes where the segment boundaries fall, so it is part of the boundary parity.
"""
from __future__ import annotations

import math
from collections.abc import Mapping
from dataclasses import dataclass

_TRUE_WORDS = frozenset(("true", "1", "yes", "y", "on"))
_FALSE_WORDS = frozenset(("false", "0", "no", "n", "off"))


@dataclass(frozen=True)
class SegmentPlan:
    segment_duration: int
    overlap: int
    num_segments: int
    estimated_calls: int
    available_calls: int
    hard_max_calls: int
    fits_budget: bool


def _coerce_int(value: object, default: int) -> int:
    """int() of numbers and numeric strings; anything else (None, lists, bad strings) gives the default."""
    if not isinstance(value, (int, float, str)):
        return default
    try:
        return int(value)
    except ValueError:
        return default


def _coerce_bool(value: object, default: bool) -> bool:
    if isinstance(value, bool):
        return value
    if isinstance(value, (int, float)):
        return value != 0
    if isinstance(value, str):
        word = value.strip().lower()
        if word in _TRUE_WORDS:
            return True
        if word in _FALSE_WORDS:
            return False
    return default


def _estimate_segments(duration: float, segment_duration: int, overlap: int) -> int:
    """Estimated segment count.  Note it strides by (segment - overlap) although plan_segments strides by
    segment; the reference does the same (budget_planner.py:50-53) and the chosen plan depends on it."""
    if duration <= 0:
        return 0
    seg = segment_duration if segment_duration > 1 else 1
    ovl = min(overlap, seg - 1)
    if ovl < 0:
        ovl = 0
    if duration <= seg:
        return 1
    step = seg - ovl
    if step <= 0:
        step = 1
    return int(math.ceil((duration - seg) / step)) + 1


def _estimate_calls(num_segments: int, max_continuations: int, retry_buffer: int, extra_calls: int = 0) -> int:
    if num_segments <= 0:
        return 0
    # one call per segment, its continuations, the merge call, optional consolidation, retries
    return num_segments * (1 + max_continuations) + 1 + extra_calls + retry_buffer


def _section(parent: object, key: str) -> dict:
    if isinstance(parent, Mapping):
        child = parent.get(key)
        if isinstance(child, dict):
            return child
    return {}


def plan_segments_with_budget(duration: float, config: Mapping[str, object], current_api_count: int) -> SegmentPlan:
    analyzer = _section(config, "analyzer")
    long_video = _section(analyzer, "long_video")

    default_segment = _coerce_int(long_video.get("default_segment_seconds"), 480)
    overlap = _coerce_int(long_video.get("overlap_seconds"), 20)
    min_segment = _coerce_int(long_video.get("min_segment_seconds"), 90)
    hard_max = _coerce_int(long_video.get("hard_max_api_calls"), 50)
    continuations = _coerce_int(analyzer.get("max_continuations"), 3)
    retries = _coerce_int(analyzer.get("retry_times"), 0)
    consolidate = _coerce_bool(long_video.get("consolidate"), True)
    raw_threshold = long_video.get("duration_threshold_seconds")

    duration = max(float(duration), 0.0)
    available = max(hard_max - int(current_api_count), 0)

    def empty() -> SegmentPlan:
        return SegmentPlan(0, 0, 0, 0, available, hard_max, False)

    if duration <= 0 or available == 0:
        return empty()

    threshold = None
    if isinstance(raw_threshold, (int, float, str)):
        try:
            threshold = float(raw_threshold)
        except ValueError:
            threshold = None

    if threshold is not None and duration < threshold:
        segment = max(int(math.ceil(duration)), 1)
        overlap = 0
    else:
        segment = max(default_segment, min_segment, 1)
        overlap = max(min(overlap, segment - 1), 0)

    extra = 1 if consolidate else 0

    def cost(seg: int, ovl: int) -> tuple[int, int]:
        n = _estimate_segments(duration, seg, ovl)
        return n, _estimate_calls(n, continuations, retries, extra)

    count, calls = cost(segment, overlap)
    if calls > available:
        overlap = 0
        count, calls = cost(segment, overlap)

    if calls > available and available > 0:
        room = (available - (1 + extra + retries)) // (1 + continuations)
        if room < 1:
            return empty()
        room = max(int(room), 1)
        overlap = 0
        while True:
            segment = max(int(math.ceil(duration / room)), min_segment, 1)
            count, calls = cost(segment, overlap)
            if calls <= available or room <= 1:
                break
            room -= 1
        if calls > available:
            return empty()

    return SegmentPlan(segment, overlap, count, calls, available, hard_max, calls <= available)

"""Host-side drop-in parity: plan, budget and manifest functions against vectors generated from the reference's
own Python (tests/golden/make_golden.py -> plan_vectors.json).  Floats are compared as bit patterns."""
import json
import math
import os
from pathlib import Path

import pytest

from video_transformer_b200 import budget_planner, video_segmenter
from video_transformer_b200.video_utils import probe_duration

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "plan_vectors.json")))


def fx(h):
    return float.fromhex(h)


def test_plan_segments_bit_exact_against_reference_vectors():
    assert len(GOLD["plan_segments"]) > 200
    for case in GOLD["plan_segments"]:
        got = video_segmenter.plan_segments(fx(case["duration"]), fx(case["segment_seconds"]),
                                            fx(case["overlap_seconds"]))
        exp = case["segments"]
        assert len(got) == len(exp), case
        for g, e in zip(got, exp):
            assert g.segment_id == e[0]
            assert [float(v).hex() for v in (g.start, g.end, g.effective_start, g.effective_end)] == e[1:], case


def test_plan_segments_float_accumulation_kat():
    # repeated addition, not id * segment_seconds
    segs = video_segmenter.plan_segments(100.0, 33.3, 1.1)
    assert [s.effective_end.hex() for s in segs] == ["0x1.0a66666666666p+5", "0x1.0a66666666666p+6",
                                                     "0x1.8f99999999999p+6", "0x1.9000000000000p+6"]
    assert segs[1].start.hex() == "0x1.0199999999999p+5" and segs[2].end.hex() == "0x1.9000000000000p+6"


def test_plan_segments_reference_test_cases():
    # same cases as /root/reference/tests/test_video_segmenter.py:89-117
    s = video_segmenter.plan_segments(duration=100.0, segment_seconds=30.0, overlap_seconds=5.0)
    assert len(s) == 4
    assert (s[0].start, s[0].end, s[0].effective_start, s[0].effective_end) == (0.0, 35.0, 0.0, 30.0)
    assert (s[1].start, s[1].end, s[1].effective_start, s[1].effective_end) == (25.0, 65.0, 30.0, 60.0)
    assert (s[3].start, s[3].end, s[3].effective_start, s[3].effective_end) == (85.0, 100.0, 90.0, 100.0)
    s = video_segmenter.plan_segments(duration=50.0, segment_seconds=20.0, overlap_seconds=-3.0)
    assert len(s) == 3 and (s[1].start, s[1].end) == (20.0, 40.0)
    assert video_segmenter.plan_segments(0.0, 10.0, 1.0) == [] and video_segmenter.plan_segments(10.0, 0.0, 1.0) == []


def test_budget_planner_matches_reference_vectors():
    assert len(GOLD["budget"]) > 300
    for case in GOLD["budget"]:
        p = budget_planner.plan_segments_with_budget(fx(case["duration"]), case["config"], case["current"])
        got = [p.segment_duration, p.overlap, p.num_segments, p.estimated_calls, p.available_calls, p.hard_max_calls,
               p.fits_budget]
        assert got == case["plan"], case
    for d, s, o, n in GOLD["estimate_segments"]:
        assert budget_planner._estimate_segments(fx(d), s, o) == n


def test_budget_planner_reference_test_cases():
    # /root/reference/tests/test_budget_planner.py:25-61
    base = {"analyzer": {"max_continuations": 3, "retry_times": 5,
                         "long_video": {"enabled": True, "default_segment_seconds": 480, "overlap_seconds": 20,
                                        "min_segment_seconds": 90, "hard_max_api_calls": 50, "consolidate": True}}}
    p = budget_planner.plan_segments_with_budget(3 * 3600, base, current_api_count=0)
    assert p.num_segments >= 1 and p.estimated_calls <= p.hard_max_calls
    base["analyzer"]["long_video"]["duration_threshold_seconds"] = 600
    p = budget_planner.plan_segments_with_budget(9 * 60, base, current_api_count=0)
    assert p.num_segments == 1 and p.overlap == 0
    exact = {"analyzer": {"max_continuations": 2, "retry_times": 0,
                          "long_video": {"default_segment_seconds": 400, "overlap_seconds": 0,
                                         "min_segment_seconds": 90, "hard_max_api_calls": 8, "consolidate": True}}}
    p = budget_planner.plan_segments_with_budget(1200, exact, current_api_count=0)
    assert p.estimated_calls == p.hard_max_calls == 8


def test_shipped_config_plans_from_survey_appendix_a():
    # values of /root/reference/config/config.yaml:84-96
    cfg = {"analyzer": {"max_continuations": 3, "retry_times": 5,
                        "long_video": {"default_segment_seconds": 480, "overlap_seconds": 20, "min_segment_seconds": 90,
                                       "hard_max_api_calls": 50, "consolidate": True}}}
    expect = {600.0: (480, 20, 2, 15), 1800.0: (480, 20, 4, 23), 7200.0: (720, 0, 10, 47), 36000.0: (3600, 0, 10, 47)}
    for d, e in expect.items():
        p = budget_planner.plan_segments_with_budget(d, cfg, 0)
        assert (p.segment_duration, p.overlap, p.num_segments, p.estimated_calls) == e


def test_manifest_lifecycle_matches_reference(tmp_path: Path):
    g = GOLD["manifest"]
    m = video_segmenter.create_manifest(video_id="vid_A", duration=65.0, segment_seconds=30.0, overlap_seconds=5.0,
                                        temp_dir=tmp_path)
    path = video_segmenter.get_manifest_path("vid_A", tmp_path)
    assert path == tmp_path / "segments" / "vid_A" / "manifest.json" and path.exists()
    text = path.read_text(encoding="utf-8").replace(m["created_at"], "@CREATED@").replace(str(tmp_path), "@TMP@")
    assert text == g["text"]                       # byte-identical file: key order, indent=2, ascii
    assert len(m["created_at"]) == len(g["created_at_sample"]) and m["created_at"].endswith("+00:00")
    assert m["segments"][0]["status"] == "pending" and m["segments"][0]["file_path"].endswith("segment_0000.mp4")
    # resume returns the file on disk, not a fresh plan
    m["segments"][0]["status"] = "completed"
    video_segmenter.save_manifest(path, m)
    again = video_segmenter.load_or_create_manifest(video_id="vid_A", duration=999.0, segment_seconds=1.0,
                                                    overlap_seconds=0.0, temp_dir=tmp_path)
    assert again["segments"][0]["status"] == "completed" and len(again["segments"]) == len(m["segments"])
    assert all(e["id"] != 0 for e in video_segmenter.pending_segments(again))
    m2 = video_segmenter.load_manifest(path)
    m2["segments"][0]["status"] = "pending"
    video_segmenter.update_segment_status(m2, 1, "failed", error="boom", increment_attempts=True)
    exp = json.loads(json.dumps(g["after_update"]).replace("@CREATED@", m["created_at"]).replace("@TMP@", str(tmp_path)))
    assert m2 == exp
    video_segmenter.update_segment_status(m2, 99, "completed")      # unknown id: warning only
    assert m2 == exp


def test_snap_to_keyframe_is_the_reference_stub():
    assert video_segmenter.snap_to_keyframe("whatever.mp4", -3) == 0.0
    assert video_segmenter.snap_to_keyframe("whatever.mp4", 12.5) == 12.5
    assert isinstance(video_segmenter.snap_to_keyframe("x", 3), float)


def test_extract_segment_failure_contract(tmp_path: Path):
    out = tmp_path / "a" / "b" / "seg.mp4"
    assert video_segmenter.extract_segment(tmp_path / "missing.mp4", 0.0, 1.0, out) is False
    assert out.parent.is_dir() and not out.exists()          # parent is created, nothing is left behind
    assert video_segmenter.extract_segment(tmp_path / "missing.mp4", 5.0, 5.0, out) is False
    junk = tmp_path / "junk.mp4"
    junk.write_bytes(bytes(1024))                            # the reference's tests use 1 KB of zeros
    assert video_segmenter.extract_segment(input_path=junk, start=0.0, end=1.0, output_path=out,
                                           stream_copy=True) is False


def test_probe_duration_contract(tmp_path: Path):
    assert probe_duration(tmp_path / "nope.mp4") == 0.0
    z = tmp_path / "zeros.mp4"
    z.write_bytes(bytes(1024))
    assert probe_duration(z) == 0.0
    assert isinstance(probe_duration(str(z)), float)

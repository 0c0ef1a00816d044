"""K4 boundary selection: the product's snap_boundaries against its scalar oracle twin (CPU), and the opt-in consumer
inside extract_segment on the GPU (BASELINE.json configs[2] shape: scene cuts on a 60 fps clip).

The reference leaves this as a stub (/root/reference/src/utils/video_segmenter.py:157-159); the rule is defined in
oracle/scene_oracle.py:snap_boundaries."""
import json

import numpy as np
import pytest

from video_transformer_b200 import container, scene, synth, video_segmenter


def test_snap_boundaries_equals_the_scalar_oracle():
    from oracle import scene_oracle
    rng = np.random.default_rng(7)
    for trial in range(200):
        fps_num, fps_den = [(30, 1), (60, 1), (30000, 1001), (25, 1), (24000, 1001)][trial % 5]
        n = int(rng.integers(50, 4000))
        cuts = np.unique(rng.integers(1, n, int(rng.integers(0, 12))))
        dur = n * fps_den / fps_num
        bounds = sorted(float(x) for x in rng.uniform(0, dur * 1.05, int(rng.integers(1, 6))))
        if trial % 3 == 0 and cuts.size:                 # exact hits and exact ties
            c = int(cuts[0])
            bounds.append(c * fps_den / fps_num)
            if cuts.size > 1:
                bounds.append((cuts[0] + cuts[1]) / 2.0 * fps_den / fps_num)
        tol = float(rng.choice([0.0, 0.04, 0.5, 2.0, 10.0]))
        got = scene.snap_boundaries(bounds, cuts, n, fps_num, fps_den, tol)
        exp = scene_oracle.snap_boundaries(bounds, cuts.tolist(), n, fps_num, fps_den, tol)
        assert len(got) == len(exp)
        for g, e in zip(got, exp):
            assert g["planned_frame"] == e["planned_frame"] and g["frame"] == e["frame"], (trial, g, e)
            assert g["snapped"] == e["snapped"] and float(g["time"]).hex() == float(e["time"]).hex()
        for t in bounds:
            assert scene.boundary_frame(t, n, fps_num, fps_den) == scene_oracle.boundary_frame(t, n, fps_num, fps_den)


def test_snapping_is_off_by_default_and_keeps_reference_values():
    assert video_segmenter.configure()["snap_tolerance_s"] == 0.0


@pytest.mark.gpu
def test_extract_segment_snaps_boundaries_to_detected_cuts(cuda, tmp_path):
    """Opt-in consumer: a requested boundary within the tolerance of a detected cut moves onto that cut; one without a
    cut in range keeps its time.  60 fps, cuts at pictures 137 and 301."""
    from oracle import scene_oracle
    w, h, n, fps = 640, 360, 420, 60
    bs, meta = synth.make_testsrc_h264(w, h, n, fps=fps, gop=20, cuts=[137, 301])
    raw = tmp_path / "clip.h264"
    raw.write_bytes(bs)
    src = tmp_path / "clip.mp4"
    container.annexb_to_mp4(raw, src)
    saved = video_segmenter.configure()
    try:
        video_segmenter.configure(target_height=180, batch_frames=16, scene_threshold=0.10, snap_tolerance_s=0.5)
        out = tmp_path / "seg" / "segment_0001.mp4"
        # start 2.0 s = picture 120 (cut 137 is 0.283 s away: snaps); end 6.0 s = picture 360 (cut 301 is 0.98 s away: stays)
        assert video_segmenter.extract_segment(src, 2.0, 6.0, out, stream_copy=False) is True
        side = json.loads(out.with_suffix(".json").read_text())
        b = side["boundaries"]
        assert b["start_snapped"] is True and b["start_frame"] == 137 and b["end_snapped"] is False
        assert b["start"] == 137 / fps and b["end"] == 6.0 and b["requested_start"] == 2.0
        exp = scene_oracle.snap_boundaries([2.0, 6.0], [137, 301], n, fps, 1, 0.5)
        assert [exp[0]["frame"], exp[1]["snapped"]] == [137, False]
        first, last = scene_oracle.frames_for_window(137 / fps, 6.0, n, fps, 1, meta["idr_frames"], False)
        assert (side["first_picture"], side["last_picture"]) == (first, last) == (137, 360)
        assert side["frames"] == last - first
        # the same call without the option keeps the requested times
        video_segmenter.configure(snap_tolerance_s=0.0)
        assert video_segmenter.extract_segment(src, 2.0, 6.0, out, stream_copy=False) is True
        side = json.loads(out.with_suffix(".json").read_text())
        assert "boundaries" not in side and side["first_picture"] == 120
    finally:
        video_segmenter.configure(**saved)

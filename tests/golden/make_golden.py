"""Generates the committed golden fixtures.  Run ONLY in the build container (it imports the reference from
/root/reference and calls the image's libswscale); the fixtures travel, this script's inputs do not.

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  plan_vectors.json   plan_segments / plan_segments_with_budget / manifest vectors from the reference's own
                      Python (/root/reference/src/utils/video_segmenter.py, budget_planner.py), floats as hex
  sws_vectors.npz     libswscale 9.1.100 outputs (SWS_ACCURATE_RND|SWS_BITEXACT and default flags) for seeded
                      planes, the pin for oracle/vt_oracle.c and the CUDA scaler
  sws_rgb_vectors.npz libswscale 9.1.100 nv12 -> rgb24 outputs (same size and scaled, default flags = what
                      `ffmpeg -vf scale=W:H -pix_fmt rgb24` produces), the pin for vto_yuv_to_rgb24 / K1b
                      (`python tests/golden/make_golden.py rgb` regenerates only this file)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")

from utils import budget_planner as ref_bp  # noqa: E402
from utils import video_segmenter as ref_vs  # noqa: E402

from oracle import ffsws  # noqa: E402


def hx(x):
    return float(x).hex()


def plan_vectors():
    rng = np.random.default_rng(20261018)
    cases = [(100.0, 30.0, 5.0), (50.0, 20.0, -3.0), (100.0, 33.3, 1.1), (600.0, 480, 20), (600.5, 480, 20),
             (1800.0, 480, 20), (3600.0, 480, 20), (7200.0, 720, 0), (36000.0, 3600, 0), (0.0, 10, 1), (-5.0, 10, 1),
             (10.0, 0, 1), (10.0, -1, 1), (10.0, 10, 0), (10.0, 10, 5), (10.0, 3, 3), (10.0, 3, 7), (0.1, 0.03, 0.01),
             (65.0, 30.0, 5.0), (180.0, 60, 0), (61.0, 60, 0), (7200.0, 3600, 0), (1e-9, 1.0, 0.0)]
    for _ in range(200):
        d = float(rng.uniform(0.01, 40000))
        s = float(rng.choice([rng.uniform(0.5, 4000), float(rng.integers(1, 4000))]))
        o = float(rng.choice([0.0, rng.uniform(-5, 50), float(rng.integers(0, 60))]))
        cases.append((d, s, o))
    out = []
    for d, s, o in cases:
        segs = ref_vs.plan_segments(d, s, o)
        out.append({"duration": hx(d), "segment_seconds": hx(s), "overlap_seconds": hx(o),
                    "segments": [[g.segment_id, hx(g.start), hx(g.end), hx(g.effective_start), hx(g.effective_end)]
                                 for g in segs]})
    return out


def budget_vectors():
    rng = np.random.default_rng(7)
    shipped = {"analyzer": {"max_continuations": 3, "retry_times": 0,
                            "long_video": {"enabled": True, "default_segment_seconds": 480, "overlap_seconds": 20,
                                           "min_segment_seconds": 90, "hard_max_api_calls": 50, "consolidate": True,
                                           "duration_threshold_seconds": 600}}}
    cfgs = [(shipped, d, c) for d in (0, -1, 30, 540, 599.9, 600, 600.5, 1800, 3600, 7200, 10800, 36000, 86400)
            for c in (0, 10, 45, 49, 50, 60)]
    for _ in range(300):
        lv = {"default_segment_seconds": int(rng.integers(1, 2000)), "overlap_seconds": int(rng.integers(-5, 200)),
              "min_segment_seconds": int(rng.integers(0, 400)), "hard_max_api_calls": int(rng.integers(0, 80)),
              "consolidate": bool(rng.integers(0, 2))}
        if rng.integers(0, 2):
            lv["duration_threshold_seconds"] = float(rng.uniform(0, 3000))
        if rng.integers(0, 8) == 0:
            lv["default_segment_seconds"] = str(lv["default_segment_seconds"])
        if rng.integers(0, 8) == 0:
            lv["consolidate"] = str(rng.choice(["yes", "off", "1", "maybe"]))
        cfg = {"analyzer": {"max_continuations": int(rng.integers(0, 5)), "retry_times": int(rng.integers(0, 6)),
                            "long_video": lv}}
        cfgs.append((cfg, float(rng.uniform(0, 40000)), int(rng.integers(0, 60))))
    cfgs.append(({}, 1000.0, 0))
    cfgs.append(({"analyzer": "nope"}, 1000.0, 3))
    out = []
    for cfg, d, c in cfgs:
        p = ref_bp.plan_segments_with_budget(d, cfg, c)
        out.append({"config": cfg, "duration": hx(d), "current": c,
                    "plan": [p.segment_duration, p.overlap, p.num_segments, p.estimated_calls, p.available_calls,
                             p.hard_max_calls, p.fits_budget]})
    est = [[hx(d), s, o, ref_bp._estimate_segments(d, s, o)]
           for d, s, o in [(7200.0, 480, 20), (600.0, 480, 20), (1800.0, 480, 20), (0.0, 5, 1), (10.0, 0, 0),
                           (10.0, 3, 9), (100.5, 10, 3)]]
    return out, est


def manifest_vector(tmp):
    m = ref_vs.create_manifest(video_id="vid_A", duration=65.0, segment_seconds=30.0, overlap_seconds=5.0,
                               temp_dir=tmp)
    text = (ref_vs.get_manifest_path("vid_A", tmp)).read_text(encoding="utf-8")
    m2 = json.loads(text)
    ref_vs.update_segment_status(m2, 1, "failed", error="boom", increment_attempts=True)
    return {"manifest": m, "text": text.replace(m["created_at"], "@CREATED@").replace(str(tmp), "@TMP@"),
            "after_update": json.loads(json.dumps(m2).replace(m["created_at"], "@CREATED@").replace(str(tmp), "@TMP@")),
            "created_at_sample": m["created_at"]}


def sws_vectors():
    rng = np.random.default_rng(99)
    ex = ffsws.SWS_ACCURATE_RND | ffsws.SWS_BITEXACT
    out = {}
    for name, (sw, sh, dw, dh) in {"a": (192, 108, 128, 72), "b": (128, 72, 64, 36), "c": (384, 216, 128, 72),
                                   "d": (128, 72, 76, 76), "e": (101, 57, 50, 28)}.items():
        y = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        u = rng.integers(0, 256, ((sh + 1) // 2, (sw + 1) // 2), dtype=np.uint8)
        v = rng.integers(0, 256, ((sh + 1) // 2, (sw + 1) // 2), dtype=np.uint8)
        out[name + "_src_y"], out[name + "_src_u"], out[name + "_src_v"] = y, u, v
        out[name + "_dims"] = np.array([sw, sh, dw, dh])
        for fl, fn in ((ffsws.SWS_BICUBIC, "bicubic"), (ffsws.SWS_BILINEAR, "bilinear"), (ffsws.SWS_AREA, "area")):
            ey, eu, ev = ffsws.scale_yuv420p(y, u, v, dw, dh, fl | ex)
            out["%s_%s_y" % (name, fn)], out["%s_%s_u" % (name, fn)], out["%s_%s_v" % (name, fn)] = ey, eu, ev
            dy, _, _ = ffsws.scale_yuv420p(y, u, v, dw, dh, fl)
            out["%s_%s_default_y" % (name, fn)] = dy
    return out


def sws_rgb_vectors():
    rng = np.random.default_rng(2026)
    out = {}
    for name, (sw, sh, dw, dh) in {"a": (64, 48, 64, 48), "b": (128, 72, 64, 36), "c": (160, 90, 96, 96),
                                   "d": (101, 61, 78, 52), "e": (192, 108, 48, 48)}.items():
        pitch = (sw + 15) // 16 * 16
        buf = rng.integers(0, 256, (sh + (sh + 1) // 2, pitch), dtype=np.uint8)
        if name == "c":                      # limited-range content with flat areas, like real video
            buf[:sh] = np.clip(buf[:sh], 16, 235)
            buf[: sh // 3, : sw // 2] = 235
            buf[sh:] = np.clip(buf[sh:], 16, 240)
        out[name + "_nv12"] = buf
        out[name + "_dims"] = np.array([sw, sh, pitch, dw, dh])
        y, uv = buf[:sh, :sw], buf[sh:, : 2 * ((sw + 1) // 2)]
        out[name + "_rgb"] = ffsws.nv12_to_rgb24(y, uv, dw, dh, ffsws.SWS_BICUBIC)
        out[name + "_rgb_bitexact"] = ffsws.nv12_to_rgb24(y, uv, dw, dh, ffsws.SWS_BICUBIC | ffsws.SWS_ACCURATE_RND |
                                                          ffsws.SWS_BITEXACT)
    return out


if __name__ == "__main__":
    if sys.argv[1:] == ["rgb"]:
        np.savez_compressed(os.path.join(HERE, "sws_rgb_vectors.npz"), swscale_version=np.array(ffsws.version()),
                            **sws_rgb_vectors())
        print("wrote sws_rgb_vectors.npz")
        sys.exit(0)
    import pathlib
    import tempfile
    bv, est = budget_vectors()
    with tempfile.TemporaryDirectory() as t:
        mv = manifest_vector(pathlib.Path(t))
    doc = {"source": "generated from /root/reference/src/utils/{video_segmenter,budget_planner}.py",
           "plan_segments": plan_vectors(), "budget": bv, "estimate_segments": est, "manifest": mv}
    with open(os.path.join(HERE, "plan_vectors.json"), "w") as f:
        json.dump(doc, f, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "sws_vectors.npz"), swscale_version=np.array(ffsws.version()),
                        **sws_vectors())
    np.savez_compressed(os.path.join(HERE, "sws_rgb_vectors.npz"), swscale_version=np.array(ffsws.version()),
                        **sws_rgb_vectors())
    print("wrote", os.listdir(HERE))

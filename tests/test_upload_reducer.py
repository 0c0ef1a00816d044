"""Drop-in for ContentAnalyzer._compress_video_for_upload (/root/reference/src/analyzer/content_analyzer.py:167-236):
gate, cache and fall-back-to-input behaviour on the CPU; the GPU product against the oracle."""
import json

import numpy as np
import pytest

from video_transformer_b200 import upload_reducer as ur


def test_small_file_is_returned_unchanged(tmp_path):
    p = tmp_path / "small.mp4"
    p.write_bytes(b"x" * 1000)
    assert ur.compress_video_for_upload(p) == p                          # <= 30 MiB: no work, as in the reference
    assert not ur.compressed_path_for(p).exists()


def test_existing_compressed_file_is_reused(tmp_path):
    p = tmp_path / "big.mp4"
    p.write_bytes(b"x" * 4096)
    c = tmp_path / "compressed_big.mp4"
    c.write_bytes(b"cached")
    assert ur.compressed_path_for(p) == c
    assert ur.compress_video_for_upload(p, max_size_mb=0.001) == c       # cache hit: returned without decoding
    assert c.read_bytes() == b"cached"


def test_failure_returns_the_input_and_leaves_nothing_behind(tmp_path):
    p = tmp_path / "notavideo.mp4"
    p.write_bytes(b"\x00" * 8192)
    assert ur.compress_video_for_upload(p, max_size_mb=0.001) == p       # unparseable input: original path, no raise
    assert not ur.compressed_path_for(p).exists()
    with pytest.raises(OSError):
        ur.compress_video_for_upload(tmp_path / "missing.mp4")          # the reference's stat() raises too


@pytest.mark.gpu
def test_reduced_artifact_holds_the_swscale_exact_360p_frames(cuda, oracle_c, tmp_path):
    from video_transformer_b200 import container, synth
    w, h, n, gop = 1280, 720, 95, 10
    bs, meta = synth.make_testsrc_h264(w, h, n, fps=30, gop=gop, cuts=[33])
    raw = tmp_path / "clip.h264"
    raw.write_bytes(bs)
    src = tmp_path / "clip.mp4"
    container.annexb_to_mp4(raw, src)
    out = ur.compress_video_for_upload(src, max_size_mb=1.0)
    assert out == tmp_path / "compressed_clip.mp4" and out.stat().st_size < src.stat().st_size
    idx = container.probe(out)
    assert (idx.width, idx.height, idx.n_frames) == (640, 360, 4) and all(idx.keyframe)   # pictures 0, 30, 60, 90
    side = json.loads(out.with_suffix(".json").read_text())
    assert side["frame_size"] == [640, 360] and side["sample_every"] == 30 and side["frames"] == 4
    assert side["cuts"] == [33]
    frames = np.fromfile(out.with_suffix(".frames"), np.uint8).reshape(4, -1)
    # expected: the decoded source pictures (samples >= 1, held from the last IDR) through the oracle's bicubic scaler
    scene, ref, exp = 0, None, {}
    for k in range(n):
        if k == 33:
            scene += 1
        if k in meta["idr_frames"]:
            ref = tuple(np.maximum(p, 1) for p in synth.testsrc_frame(w, h, k, scene))
        if k % 30 == 0:
            exp[k] = ref
    for i, k in enumerate(sorted(exp)):
        ey, eu, ev = oracle_c.scale_yuv420p(*exp[k], 640, 360, oracle_c.BICUBIC)
        assert np.array_equal(frames[i], np.concatenate([ey.reshape(-1), eu.reshape(-1), ev.reshape(-1)])), k
    # the MP4 pictures are those frames with PCM's "no zero sample" rule applied; any H.264 decoder reads them
    cv2 = pytest.importorskip("cv2")
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    ok, img = cap.read()
    assert ok
    y0 = np.asarray(img).reshape(-1)[:640 * 360]
    assert np.array_equal(y0, np.maximum(frames[0][:640 * 360], 1))
    # second call: cache hit
    assert ur.compress_video_for_upload(src, max_size_mb=1.0) == out


@pytest.mark.gpu
def test_reducer_never_enlarges(cuda, tmp_path):
    from video_transformer_b200 import container, synth
    bs, _ = synth.make_testsrc_h264(320, 240, 40, fps=30, gop=40)       # one IDR + 39 skips: tiny; every picture sampled
    raw = tmp_path / "c.h264"
    raw.write_bytes(bs)
    src = tmp_path / "c.mp4"
    container.annexb_to_mp4(raw, src)
    assert ur.compress_video_for_upload(src, max_size_mb=0.01, target_height=120, sample_fps=30.0) == src
    assert not ur.compressed_path_for(src).exists()

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
def test_reduced_artifact_is_a_smaller_mjpeg_mp4_with_the_audio_kept(cuda, oracle_c, tmp_path):
    """720p source with a PCM audio trak -> compressed_<name>: a Motion-JPEG MP4 whose samples are byte-identical to the
    oracle's JPEG of the oracle's bicubic 360p frames (pictures 0, 30, 60, 90), decodable by libavcodec, with the audio
    track's samples copied verbatim; the artefact is smaller than the input."""
    import struct
    from mp4_fixture import write_av_mp4
    from video_transformer_b200 import container, isobmff, synth
    cv2 = pytest.importorskip("cv2")
    w, h, n, gop, fps = 1280, 720, 95, 10, 30
    wr = synth.H264PcmWriter(w, h, fps, 1)
    nal_samples, keys, exp, ref, scene = [], [], {}, None, 0
    for k in range(n):
        if k == 33:
            scene += 1
        if k % gop == 0 or k == 33:
            pic = synth.testsrc_frame(w, h, k, scene)
            nal_samples.append([wr.idr(*pic, with_params=False)[4:]])
            keys.append(True)
            ref = tuple(np.maximum(p, 1) for p in pic)
        else:
            nal_samples.append([wr.skip()[4:]])
            keys.append(False)
        if k % 30 == 0:
            exp[k] = ref
    rate = 16000
    pcm = (np.arange(n * rate // fps) % 977).astype(np.int16)[:, None]
    src = tmp_path / "clip.mp4"
    meta = write_av_mp4(src, sps=wr._sps[4:], pps=wr._pps[4:], video_samples=nal_samples, keyframes=keys, width=w,
                        height=h, timescale=30000, delta=1000, audio_pcm=pcm, audio_rate=rate, audio_channels=1)
    out = ur.compress_video_for_upload(src, max_size_mb=1.0)
    assert out == tmp_path / "compressed_clip.mp4" and out.stat().st_size < src.stat().st_size / 10
    movie = isobmff.read_movie(out)
    assert [(t.codec, t.n) for t in movie.tracks] == [(b"jpeg", 4), (b"sowt", len(pcm))]
    side = json.loads(out.with_suffix(".json").read_text())
    assert side["frame_size"] == [640, 360] and side["sample_every"] == 30 and side["frames"] == 4
    assert side["cuts"] == [33] and side["audio_tracks"] == 1 and side["codec"] == "mjpeg"
    data = out.read_bytes()
    vt_, at_ = movie.tracks
    for i, k in enumerate(sorted(exp)):
        ey, eu, ev = oracle_c.scale_yuv420p(*exp[k], 640, 360, oracle_c.BICUBIC)
        want = oracle_c.jpeg_encode(ey, eu, ev, quality=ur.JPEG_QUALITY, restart_interval=40, expand_range=True)
        got = data[int(vt_.offsets[i]):int(vt_.offsets[i]) + int(vt_.sizes[i])]
        assert got == want, k
    # audio: every PCM frame of the source, byte for byte, same duration
    offs, sizes = at_.offsets.astype(np.int64), at_.sizes.astype(np.int64)
    brk = np.nonzero(offs[1:] != offs[:-1] + sizes[:-1])[0] + 1
    audio = b"".join(data[offs[a_]:offs[b_ - 1] + sizes[b_ - 1]]
                     for a_, b_ in zip(np.concatenate(([0], brk)), np.concatenate((brk, [offs.size]))))
    assert audio == meta["audio_bytes"]
    assert abs(movie.duration_seconds() - max(4 * 1.0, len(pcm) / rate)) < 2e-3
    # libavformat/libavcodec read it: 4 pictures at 1 fps whose luma is the JPEG's luma
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 4 and abs(cap.get(cv2.CAP_PROP_FPS) - 1.0) < 1e-6
    shown = 0
    while True:
        ok, img = cap.read()
        if not ok:
            break
        want = cv2.imdecode(np.frombuffer(data[int(vt_.offsets[shown]):int(vt_.offsets[shown]) + int(vt_.sizes[shown])],
                                          np.uint8), cv2.IMREAD_COLOR)
        assert img.shape == (360, 640, 3) and np.abs(img.astype(int) - want.astype(int)).mean() < 4.0
        shown += 1
    assert shown == 4
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

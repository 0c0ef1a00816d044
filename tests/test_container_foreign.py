"""The drop-in on files it did NOT write.

Reference behaviour being matched: `ffprobe format=duration` for any container/codec
(/root/reference/src/utils/video_utils.py:7-38) and `ffmpeg -ss S -i IN -t D -movflags +faststart -c copy OUT`, which
copies every stream without decoding (/root/reference/src/utils/video_segmenter.py:118-137); the shape of the checks
follows the reference's own integration test (/root/reference/tests/test_video_segmenter.py:147-178: success, file
exists, non-empty) and then goes further (picture counts, byte-identical samples, all tracks present).

Foreign inputs: (i) MP4/MOV/MKV/AVI files muxed by libavformat through OpenCV's VideoWriter (MPEG-4 part 2, VP9,
MJPEG); (ii) MP4s from the test-side writer tests/mp4_fixture.py with AUD+SEI+slice samples, a `ctts` box, an edit list
and a second (PCM `sowt`) audio trak.
"""
import json
from pathlib import Path
import struct

import numpy as np
import pytest

from mp4_fixture import write_av_mp4, write_fragmented_av, write_fragmented_mp4, write_mkv
from video_transformer_b200 import container, isobmff, synth, video_segmenter
from video_transformer_b200.video_utils import probe_duration

cv2 = pytest.importorskip("cv2")


def _cv_write(path, fourcc, n=75, fps=25.0, size=(320, 240)):
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), fps, size)
    if not vw.isOpened():
        pytest.skip("OpenCV cannot write %s to %s here" % (fourcc, path.suffix))
    for k in range(n):
        img = np.zeros((size[1], size[0], 3), np.uint8)
        img[:, :] = (k * 3 % 255, 50, 200)
        cv2.putText(img, str(k), (20, 100), cv2.FONT_HERSHEY_SIMPLEX, 2, (255, 255, 255), 3)
        cv2.rectangle(img, (k * 3, 150), (k * 3 + 40, 200), (0, 255, 0), -1)
        vw.write(img)
    vw.release()
    if not path.exists() or path.stat().st_size == 0:
        pytest.skip("OpenCV wrote nothing for %s" % fourcc)


def _cv_frames(path):
    cap = cv2.VideoCapture(str(path), cv2.CAP_FFMPEG)
    out = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        out.append(fr)
    return out


def _cv_duration(path):
    cap = cv2.VideoCapture(str(path), cv2.CAP_FFMPEG)
    return cap.get(cv2.CAP_PROP_FRAME_COUNT) / cap.get(cv2.CAP_PROP_FPS)


@pytest.fixture(autouse=True)
def _options():
    saved = video_segmenter.configure()
    yield
    video_segmenter.configure(**saved)


@pytest.mark.parametrize("name,fourcc", [("a.mp4", "mp4v"), ("b.mp4", "vp09"), ("c.mkv", "VP90"), ("d.mkv", "mp4v"),
                                         ("e.avi", "MJPG"), ("f.mov", "mp4v"), ("g.webm", "VP90"), ("h.flv", "FLV1")])
def test_probe_duration_matches_libavformat(tmp_path, name, fourcc):
    p = tmp_path / name
    _cv_write(p, fourcc, n=83, fps=25.0)
    assert abs(probe_duration(p) - _cv_duration(p)) < 1e-3
    assert abs(probe_duration(p) - 83 / 25.0) < 1e-3


def test_flv_without_metadata_duration_uses_its_last_tag(tmp_path):
    p = tmp_path / "a.flv"
    _cv_write(p, "FLV1", n=50, fps=25.0)
    b = bytearray(p.read_bytes())
    at = b.find(b"\x00\x08duration\x00")
    assert at > 0 and abs(probe_duration(p) - 2.0) < 1e-3
    b[at + 2:at + 10] = b"xuration"
    p.write_bytes(bytes(b))
    assert abs(probe_duration(p) - 49 / 25.0) < 1e-3      # timestamp of the last picture
    p.write_bytes(bytes(b[:len(b) // 2]))                  # truncated: the trailing tag size is garbage
    assert probe_duration(p) >= 0.0


@pytest.mark.parametrize("fourcc", ["mp4v", "vp09"])
@pytest.mark.parametrize("stream_copy", [True, False])
def test_cut_of_libavformat_mp4(tmp_path, fourcc, stream_copy):
    """A stream copy needs no decoder: the cut succeeds for codecs the pixel pass cannot touch, the sidecar says why
    the frame buffers are absent, and libavcodec decodes the cut to exactly the source's pictures."""
    from oracle import scene_oracle
    src = tmp_path / ("src_%s.mp4" % fourcc)
    n, fps = 90, 25
    _cv_write(src, fourcc, n=n, fps=float(fps))
    idx = container.probe(src)
    assert idx is not None and idx.n_frames == n and idx.extra["codec"] == fourcc and not idx.extra["decodable"]
    keys = np.nonzero(idx.keyframe)[0]
    assert keys.size >= 2 and keys[0] == 0
    start, end = 1.30, 2.9
    out = tmp_path / "segments" / "v" / "segment_0001.mp4"
    assert video_segmenter.extract_segment(src, start, end, out, stream_copy) is True
    first, last = scene_oracle.frames_for_window(start, end, n, fps, 1, keys, True)      # file content: from the keyframe
    first_acc, _ = scene_oracle.frames_for_window(start, end, n, fps, 1, keys, False)
    side = json.loads(out.with_suffix(".json").read_text())
    assert side["frames"] is None and "NVDEC" in side["reason"] and not out.with_suffix(".frames").exists()
    assert (side["first_picture"], side["last_picture"]) == (first, last)
    cut = container.probe(out)
    assert cut.n_frames == last - first and cut.extra["codec"] == fourcc and bool(cut.keyframe[0])
    ref = _cv_frames(src)
    got = _cv_frames(out)
    shown_from = first if stream_copy else first_acc      # the accurate cut hides the lead-in through its edit list
    assert len(got) == last - shown_from
    for i, fr in enumerate(got):
        assert np.array_equal(fr, ref[shown_from + i]), i
    exp = (last - shown_from) / fps
    assert abs(probe_duration(out) - exp) < 2e-3
    # moov before mdat (+faststart)
    kinds = [k for k, *_ in isobmff.top_level(out)]
    assert kinds.index(b"moov") < kinds.index(b"mdat")


def _pcm_samples(w, h, n, gop):
    """(sps, pps, [[AUD, SEI, slice], ...], keyframes, expected luma per picture)."""
    wr = synth.H264PcmWriter(w, h, 30, 1)
    aud = b"\x09\xf0"
    samples, keys, luma = [], [], []
    cur = None
    for k in range(n):
        sei = b"\x06\x05\x14" + bytes(range(16)) + struct.pack(">I", k) + b"\x80"
        if k % gop == 0:
            y, u, v = synth.testsrc_frame(w, h, k, k // gop)
            nal = wr.idr(y, u, v, with_params=False)[4:]
            cur = np.maximum(y, 1)
            keys.append(True)
        else:
            nal = wr.skip()[4:]
            keys.append(False)
        samples.append([aud, sei, nal])
        luma.append(cur)
    return wr._sps[4:], wr._pps[4:], samples, keys, luma


@pytest.mark.parametrize("moov_first,co64,version1,stz2", [(False, False, False, False), (True, True, False, False),
                                                          (False, True, True, True)])
def test_cut_keeps_every_track_byte_identical(tmp_path, moov_first, co64, version1, stz2):
    """AUD+SEI+slice samples, a ctts box with an edit list, and a PCM audio trak: the cut carries both tracks, sample
    bytes and sample descriptions verbatim, and libavcodec decodes it to the expected pictures."""
    from oracle import scene_oracle
    w, h, n, gop, fps = 192, 160, 120, 12, 30
    sps, pps, samples, keys, luma = _pcm_samples(w, h, n, gop)
    rate = 48000
    t = np.arange(n * rate // fps)
    pcm = np.stack([(8000 * np.sin(t * 0.05)).astype(np.int16), (t % 1000).astype(np.int16)], 1)
    src = tmp_path / "src.mp4"
    delta = 512
    meta = write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h,
                        timescale=fps * delta, delta=delta, ctts=[delta] * n, video_media_time=delta, audio_pcm=pcm,
                        audio_rate=rate, moov_first=moov_first, co64=co64, version1=version1, stz2=stz2)
    assert abs(probe_duration(src) - n / fps) < 1e-3
    assert abs(probe_duration(src) - _cv_duration(src)) < 1e-3
    movie = isobmff.read_movie(src)
    assert [t_.codec for t_ in movie.tracks] == [b"avc1", b"sowt"]
    idx = container.probe(src)
    assert idx.extra["decodable"] and idx.extra["single_slice"] and idx.extra["cfr"]
    assert container.classify_pcm(idx) is True            # AUD and SEI were walked over; the slice is what is indexed
    video_segmenter.configure(frame_buffers=False)
    start, end = 1.25, 3.0
    out = tmp_path / "seg" / "segment_0002.mp4"
    assert video_segmenter.extract_segment(input_path=src, start=start, end=end, output_path=out) is True
    first, last = scene_oracle.frames_for_window(start, end, n, fps, 1, np.nonzero(keys)[0], True)
    cut = isobmff.read_movie(out)
    assert [t_.codec for t_ in cut.tracks] == [b"avc1", b"sowt"]
    v_src, a_src = movie.tracks
    v_cut, a_cut = cut.tracks
    assert v_cut.stsd == v_src.stsd and a_cut.stsd == a_src.stsd          # sample descriptions verbatim
    data = out.read_bytes()
    got = [data[int(o):int(o) + int(z)] for o, z in zip(v_cut.offsets, v_cut.sizes)]
    assert got == meta["video_samples"][first:last]
    assert v_cut.cts_off is not None and (v_cut.cts_off == delta).all()
    assert v_cut.edit_shift(cut.timescale) == (0.0, delta)
    assert bool(v_cut.sync[0]) and v_cut.sync.tolist() == keys[first:last]
    # audio: exactly the PCM frames that overlap the video's time span, byte for byte
    t_lo, t_hi = first / fps, last / fps
    a0, a1 = int(np.floor(t_lo * rate + 1e-9)), int(np.ceil(t_hi * rate - 1e-9))
    bpf = meta["bytes_per_audio_frame"]
    a_got = b"".join(data[int(o):int(o) + int(z)] for o, z in zip(a_cut.offsets[::a_cut.n // 50 or 1], a_cut.sizes[::a_cut.n // 50 or 1]))
    a_all = bytearray()
    offs, sizes = a_cut.offsets.astype(np.int64), a_cut.sizes.astype(np.int64)
    brk = np.nonzero(offs[1:] != offs[:-1] + sizes[:-1])[0] + 1
    for s_, e_ in zip(np.concatenate(([0], brk)), np.concatenate((brk, [offs.size]))):
        a_all += data[offs[s_]:offs[e_ - 1] + sizes[e_ - 1]]
    assert a_cut.n == a1 - a0 and bytes(a_all) == meta["audio_bytes"][a0 * bpf:a1 * bpf] and a_got
    assert abs(a_cut.n / rate - (last - first) / fps) <= 2.0 / rate
    assert abs(probe_duration(out) - (last - first) / fps) < 2e-3
    # libavformat/libavcodec accept the file and decode the expected pictures
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    k = first
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), luma[k]), k
        k += 1
    assert k == last


def test_frame_accurate_cut_of_pcm_stream_with_audio(tmp_path):
    """stream_copy=False inside a GOP of a PCM-intra stream: the first picture is re-expressed by its IDR's sample
    (exact), the audio track starts at that picture."""
    w, h, n, gop, fps = 128, 96, 60, 10, 30
    sps, pps, samples, keys, luma = _pcm_samples(w, h, n, gop)
    rate = 8000
    pcm = (np.arange(n * rate // fps) % 251).astype(np.int16)[:, None]
    src = tmp_path / "src.mp4"
    meta = write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h,
                        timescale=30000, delta=1000, audio_pcm=pcm, audio_rate=rate, audio_channels=1)
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.5, 1.5, out, stream_copy=False) is True
    cut = isobmff.read_movie(out)
    v_cut, a_cut = cut.tracks
    assert v_cut.n == 30 and bool(v_cut.sync[0])
    data = out.read_bytes()
    assert data[int(v_cut.offsets[0]):int(v_cut.offsets[0]) + int(v_cut.sizes[0])] == meta["video_samples"][10]
    assert a_cut.n == rate and int(a_cut.offsets[0]) > 0
    frames = _cv_frames(out)
    assert len(frames) == 30


def test_unknown_h264_stream_fails_closed_for_the_pcm_trick(tmp_path):
    """A mid-GOP frame-accurate cut of a stream that is not provably PCM-intra must not splice samples: it keeps the
    keyframe lead-in and hides it with the edit list instead."""
    w, h, n, gop = 128, 96, 40, 10
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    for s in samples:
        if s[2][0] & 31 == 1:
            s.append(s[2])                   # a second slice NAL per P picture: outside what K0 indexes
    src = tmp_path / "multi.mp4"
    meta = write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h,
                        timescale=30000, delta=1000)
    idx = container.probe(src)
    assert idx.extra["single_slice"] is False and idx.extra["decodable"] is False
    assert container.classify_pcm(idx) is False
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.5, 1.0, out, stream_copy=False) is True
    cut = isobmff.read_movie(out)
    v = cut.tracks[0]
    data = out.read_bytes()
    got = [data[int(o):int(o) + int(z)] for o, z in zip(v.offsets, v.sizes)]
    assert got == meta["video_samples"][10:30]          # from the keyframe at 0.333 s, verbatim
    empty, media_time = v.edit_shift(cut.timescale)
    assert empty == 0.0 and media_time == 5 * 1000       # presentation starts at picture 15 = 0.5 s
    side = json.loads(out.with_suffix(".json").read_text())
    assert side["frames"] is None and (side["first_picture"], side["last_picture"]) == (10, 30)


def test_audio_that_starts_late_keeps_its_offset(tmp_path):
    w, h, n, gop, fps = 128, 96, 60, 15, 30
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    rate = 8000
    pcm = (np.arange(rate) % 199).astype(np.int16)[:, None]          # 1 s of audio, presented from t = 0.75 s
    src = tmp_path / "late.mp4"
    write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h, timescale=30000,
                 delta=1000, audio_pcm=pcm, audio_rate=rate, audio_channels=1, audio_empty_edit=750)
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.5, 1.5, out) is True          # video from the keyframe at 0.5 s
    cut = isobmff.read_movie(out)
    v, a = cut.tracks
    assert v.n == 30
    empty, mt = a.edit_shift(cut.timescale)
    assert abs(empty - 0.25) < 2e-3 and mt == 0           # audio begins 0.25 s into the cut
    assert a.n == int(0.75 * rate)                         # and runs to the end of the video span


def test_transform_coded_audio_keeps_its_preroll_behind_the_edit_list(tmp_path):
    """An `mp4a` track (the payload is irrelevant to a stream copy) gets one extra sample before the first audible one,
    and the edit list starts the presentation after it."""
    w, h, n, gop = 128, 96, 60, 15
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    rate = 8000
    pcm = (np.arange(2 * rate) % 199).astype(np.int16)[:, None]
    src = tmp_path / "aac.mp4"
    meta = write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h, timescale=30000,
                        delta=1000, audio_pcm=pcm, audio_rate=rate, audio_channels=1, audio_codec=b"mp4a")
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.5, 1.5, out) is True
    cut = isobmff.read_movie(out)
    v, a = cut.tracks
    assert a.codec == b"mp4a" and a.n == rate + 1                     # one pre-roll sample before t = 0.5 s
    empty, mt = a.edit_shift(cut.timescale)
    assert empty == 0.0 and mt == 1                                    # ... which the edit list skips
    data = out.read_bytes()
    first = data[int(a.offsets[0]):int(a.offsets[0]) + 2]
    assert first == meta["audio_bytes"][(rate // 2 - 1) * 2:(rate // 2) * 2]


@pytest.mark.parametrize("with_mehd", [True, False])
def test_fragmented_mp4_duration(tmp_path, with_mehd):
    p = tmp_path / "frag.mp4"
    exp = write_fragmented_mp4(p, with_mehd=with_mehd)
    assert abs(probe_duration(p) - exp) < 1e-3
    movie = isobmff.read_movie(p)
    assert movie.fragmented and movie.tracks[0].n == 100 and int(movie.tracks[0].deltas.sum()) == 100 * 512


@pytest.mark.parametrize("base_is_moof,with_tfdt", [(True, True), (False, False)])
def test_fragmented_source_is_cut_like_any_other(tmp_path, base_is_moof, with_tfdt):
    """moof/traf/tfhd/tfdt/trun are indexed into ordinary sample tables: the cut of a fragmented file is a plain faststart
    MP4 with the same sample bytes, which libavcodec decodes to the expected pictures."""
    from oracle import scene_oracle
    w, h, n, gop, fps = 128, 96, 60, 10, 30
    sps, pps, samples, keys, luma = _pcm_samples(w, h, n, gop)
    src = tmp_path / "frag.mp4"
    meta = write_fragmented_av(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h,
                               base_is_moof=base_is_moof, with_tfdt=with_tfdt)
    assert abs(probe_duration(src) - n / fps) < 1e-3
    assert len(_cv_frames(src)) == n                      # libavformat agrees that this is a fragmented file with n pictures
    idx = container.probe(src)
    assert idx.n_frames == n and idx.keyframe.tolist() == keys and idx.extra["decodable"]
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    start, end = 0.7, 1.6
    assert video_segmenter.extract_segment(src, start, end, out) is True
    first, last = scene_oracle.frames_for_window(start, end, n, fps, 1, np.nonzero(keys)[0], True)
    cut = isobmff.read_movie(out)
    assert not cut.fragmented and cut.tracks[0].n == last - first
    data = out.read_bytes()
    v = cut.tracks[0]
    assert [data[int(o):int(o) + int(z)] for o, z in zip(v.offsets, v.sizes)] == meta["video_samples"][first:last]
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    k = first
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), luma[k]), k
        k += 1
    assert k == last


def test_keyframe_at_or_before_on_foreign_file(tmp_path):
    src = tmp_path / "k.mp4"
    _cv_write(src, "mp4v", n=60, fps=25.0)
    idx = container.probe(src)
    keys = np.nonzero(idx.keyframe)[0]
    t = (int(keys[1]) + 3) / 25.0
    assert video_segmenter.keyframe_at_or_before(src, t) == int(keys[1]) / 25.0
    assert video_segmenter.keyframe_at_or_before(src, 0.01) == 0.0
    assert video_segmenter.snap_to_keyframe(src, t) == t          # the reference's stub semantics are kept


def test_truncated_and_garbage_inputs_are_refused(tmp_path):
    src = tmp_path / "t.mp4"
    _cv_write(src, "mp4v", n=40, fps=25.0)
    raw = src.read_bytes()
    cutoff = tmp_path / "trunc.mp4"
    kinds = isobmff.top_level(src)
    mdat = next(b for b in kinds if b[0] == b"mdat")
    moov = next(b for b in kinds if b[0] == b"moov")
    if moov[1] > mdat[1]:                   # moov last: losing the tail loses the index
        cutoff.write_bytes(raw[: moov[1] + 20])
        assert probe_duration(cutoff) == 0.0
        assert video_segmenter.extract_segment(cutoff, 0.0, 1.0, tmp_path / "o.mp4") is False
    junk = tmp_path / "junk.mp4"
    junk.write_bytes(b"\x00\x00\x00\x18ftypisom" + bytes(range(200)))
    assert probe_duration(junk) == 0.0
    assert video_segmenter.extract_segment(junk, 0.0, 1.0, tmp_path / "o2.mp4") is False


@pytest.mark.parametrize("name,fourcc", [("m.mkv", "VP90"), ("n.mkv", "mp4v"), ("o.webm", "VP90"), ("p.mkv", "VP80")])
def test_matroska_source_is_stream_copied_into_mp4(tmp_path, name, fourcc):
    """`ffmpeg -ss S -i IN.webm -t D -c copy OUT.mp4` re-wraps Matroska streams: VP9 and MPEG-4 tracks written by
    libavformat's matroska muxer are indexed, cut at a keyframe and decode to exactly the source's pictures."""
    from oracle import scene_oracle
    src = tmp_path / name
    n, fps = 90, 25
    _cv_write(src, fourcc, n=n, fps=float(fps))
    idx = container.probe(src)
    assert idx is not None and idx.n_frames == n and (idx.fps_num, idx.fps_den) == (fps, 1)
    assert idx.extra["container"] == "matroska" and not idx.extra["decodable"]
    keys = np.nonzero(idx.keyframe)[0]
    out = tmp_path / "seg" / "segment_0000.mp4"
    start, end = 1.3, 2.9
    assert video_segmenter.extract_segment(src, start, end, out) is True
    first, last = scene_oracle.frames_for_window(start, end, n, fps, 1, keys, True)
    side = json.loads(out.with_suffix(".json").read_text())
    assert (side["first_picture"], side["last_picture"]) == (first, last) and side["frames"] is None
    ref, got = _cv_frames(src), _cv_frames(out)
    assert len(got) == last - first
    for i, fr in enumerate(got):
        assert np.array_equal(fr, ref[first + i]), i
    assert abs(probe_duration(out) - (last - first) / fps) < 2e-3


def test_matroska_avc_and_opus_tracks_become_avc1_and_opus(tmp_path):
    """A hand-written MKV (AVC video with AUD+SEI+slice samples, Opus audio): both tracks arrive in the MP4 with their
    sample bytes unchanged, avcC verbatim, OpusHead re-expressed as dOps (big endian), and libavcodec decodes the video."""
    from oracle import scene_oracle
    w, h, n, gop, fps = 128, 96, 60, 10, 30
    sps, pps, samples, keys, luma = _pcm_samples(w, h, n, gop)
    packets = [bytes([0xFC, k & 0xFF]) + bytes((k * 7 + j) & 0xFF for j in range(40 + k % 5)) for k in range(100)]   # 2 s
    src = tmp_path / "av.mkv"
    meta = write_mkv(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h, fps=fps,
                     opus_packets=packets)
    assert abs(probe_duration(src) - 2.0) < 1e-3
    assert len(_cv_frames(src)) == n                     # libavformat reads the fixture as 60 pictures
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.7, 1.6, out) is True
    first, last = scene_oracle.frames_for_window(0.7, 1.6, n, fps, 1, np.nonzero(keys)[0], True)
    cut = isobmff.read_movie(out)
    assert [t.codec for t in cut.tracks] == [b"avc1", b"Opus"]
    v, a = cut.tracks
    data = out.read_bytes()
    assert [data[int(o):int(o) + int(z)] for o, z in zip(v.offsets, v.sizes)] == meta["video_samples"][first:last]
    t_lo, t_hi = round(first * 1000 / fps) / 1000, round(last * 1000 / fps) / 1000
    hit = [k for k in range(len(packets)) if (k * 20 + 20) / 1000 > t_lo and k * 20 / 1000 < t_hi]
    want = packets[max(0, hit[0] - 4):hit[-1] + 1]                      # 80 ms of Opus pre-roll, hidden by the edit list
    assert [data[int(o):int(o) + int(z)] for o, z in zip(a.offsets, a.sizes)] == want
    assert a.edit_shift(cut.timescale)[1] >= 60                        # media time skips the pre-roll packets
    dops = a.stsd[a.stsd.index(b"dOps") + 4:]
    assert dops[:2] == bytes([0, 2]) and struct.unpack_from(">HIh", dops, 2) == (312, 48000, 0) and dops[10] == 0
    assert b"avcC" in v.stsd and meta["video_samples"][0][4:6] != b""      # avcC carried
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    k = first
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), luma[k]), k
        k += 1
    assert k == last


@pytest.mark.parametrize("parallel", [False, True])
def test_mapped_and_in_kernel_copies_write_the_same_file(tmp_path, monkeypatch, parallel):
    """The segment MP4 written through the recycled mapping (RAM-backed output directory) is byte-identical to the one
    copy_file_range writes, including a re-cut into recycled pages and a first-sample replacement; `parallel` lowers
    the size from which a sample range is copied by several threads (gigabyte ranges in production)."""
    import shutil
    import tempfile
    from video_transformer_b200 import landing
    if parallel:
        monkeypatch.setattr(isobmff, "_PARALLEL_COPY_MIN", 4096)
    if not landing.on_memory_fs("/dev/shm"):
        pytest.skip("no RAM-backed file system here")
    w, h, n, gop, fps = 96, 80, 90, 10, 30
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    rate = 48000
    t = np.arange(n * rate // fps)
    pcm = np.stack([(t % 311).astype(np.int16), (t % 1000).astype(np.int16)], 1)
    src = tmp_path / "src.mp4"
    write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h, timescale=fps * 512,
                 delta=512, audio_pcm=pcm, audio_rate=rate)
    movie = isobmff.read_movie(src)
    shm = tempfile.mkdtemp(prefix="vt_test_", dir="/dev/shm")
    try:
        for k, (start, end, copy) in enumerate([(0.5, 2.0, True), (1.0, 2.6, True), (0.7, 1.9, False)]):
            plain = tmp_path / ("plain_%d.mp4" % k)
            fast = Path(shm) / "seg.mp4"
            kw = {}
            if not copy:
                vt = movie.video_track()
                kw["first_sample"] = src.read_bytes()[int(vt.offsets[20]):int(vt.offsets[20]) + int(vt.sizes[20])]
            r0 = isobmff.cut_movie(movie, start, end, plain, stream_copy=copy, **kw)
            if fast.exists():
                fast.unlink()                   # the consumer deleted the previous segment: its pages are reused
            r1 = isobmff.cut_movie(movie, start, end, fast, stream_copy=copy, mapped=True, **kw)
            assert r0 == r1 and fast.read_bytes() == plain.read_bytes()
            assert landing.stats()["mapped_files"] == 1
    finally:
        landing.release_all()
        shutil.rmtree(shm, ignore_errors=True)


def test_uncompressed_audio_with_millions_of_frames_is_handled_in_groups(tmp_path, monkeypatch):
    """A PCM track has one sample per audio frame (two hours at 48 kHz: 345 million).  Large uniform tracks are read as
    groups of frames inside their chunks; the cut still carries a true per-frame table (fixed-size stsz, one stts run),
    byte-identical audio, and starts its presentation on the same frame as the per-sample path."""
    w, h, n, gop, fps = 96, 80, 120, 12, 30
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    rate = 48000
    t = np.arange(n * rate // fps)
    pcm = np.stack([(t % 32749).astype(np.int16), (-(t % 1021)).astype(np.int16)], 1)
    src = tmp_path / "src.mp4"
    meta = write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h,
                        timescale=fps * 512, delta=512, audio_pcm=pcm, audio_rate=rate, audio_chunk=1500)
    bpf = meta["bytes_per_audio_frame"]
    exact = isobmff.read_movie(src)
    assert exact.tracks[1].unit is None and exact.tracks[1].n == t.size
    monkeypatch.setattr(isobmff, "UNIFORM_MIN_SAMPLES", 1000)
    monkeypatch.setattr(isobmff, "UNIFORM_GROUP", 256)
    grouped = isobmff.read_movie(src)
    a = grouped.tracks[1]
    assert a.unit == (bpf, 1) and a.n == -(-1500 // 256) * (t.size // 1500) and int(a.sizes.sum()) == t.size * bpf
    assert int(a.deltas.sum()) == t.size and a.sync.all() and a.cts_off is None
    data = src.read_bytes()
    assert b"".join(data[int(o):int(o) + int(z)] for o, z in zip(a.offsets, a.sizes)) == meta["audio_bytes"]
    assert abs(grouped.duration_seconds() - exact.duration_seconds()) < 1e-9
    start, end = 1.25, 3.0
    out_g, out_e = tmp_path / "grouped.mp4", tmp_path / "exact.mp4"
    r_g = isobmff.cut_movie(grouped, start, end, out_g)
    r_e = isobmff.cut_movie(exact, start, end, out_e)
    assert (r_g.first, r_g.last) == (r_e.first, r_e.last)
    monkeypatch.undo()                                    # read both outputs sample by sample
    cut_g, cut_e = isobmff.read_movie(out_g), isobmff.read_movie(out_e)
    ag, ae = cut_g.tracks[1], cut_e.tracks[1]
    assert ag.stsd == ae.stsd == exact.tracks[1].stsd and ag.unit is None
    assert (ag.sizes == bpf).all() and (ag.deltas == 1).all()
    # the grouped cut holds whole groups: a few more frames on either side, the same frames in between
    t_lo, t_hi = r_e.first / fps, r_e.last / fps
    a0, a1 = int(np.floor(t_lo * rate + 1e-9)), int(np.ceil(t_hi * rate - 1e-9))
    assert ae.n == a1 - a0
    g_lo = (a0 // 1500) * 1500 + ((a0 % 1500) // 256) * 256
    g_hi = min(((a1 - 1) // 1500) * 1500 + (((a1 - 1) % 1500) // 256 + 1) * 256, ((a1 - 1) // 1500 + 1) * 1500)
    assert ag.n == g_hi - g_lo and g_lo <= a0 and g_hi >= a1
    dg = out_g.read_bytes()
    offs, sizes = ag.offsets.astype(np.int64), ag.sizes.astype(np.int64)
    brk = np.nonzero(offs[1:] != offs[:-1] + sizes[:-1])[0] + 1
    got = b"".join(dg[offs[s_]:offs[e_ - 1] + sizes[e_ - 1]]
                   for s_, e_ in zip(np.concatenate(([0], brk)), np.concatenate((brk, [offs.size]))))
    assert got == meta["audio_bytes"][g_lo * bpf:g_hi * bpf]
    # presentation: both edit lists start the audio on the same source frame (the video's first picture)
    assert ag.edit_shift(cut_g.timescale)[1] + g_lo == ae.edit_shift(cut_e.timescale)[1] + a0
    assert abs(probe_duration(out_g) - probe_duration(out_e)) < 2e-3
    cap = cv2.VideoCapture(str(out_g), cv2.CAP_FFMPEG)
    k = 0
    while cap.read()[0]:
        k += 1
    assert k == r_g.last - r_g.first


def test_corrupt_run_lengths_do_not_allocate_the_world(tmp_path):
    """A flipped byte in a stts/ctts/stsc run length (0xCE000000 samples ...) must not expand into gigabytes: tables are
    expanded up to the track's sample count only, and a sample count beyond the limit is refused."""
    import time
    import torch  # noqa: F401  (imported lazily by the cut path: keep its seconds out of the timings below)
    w, h, n, gop = 64, 48, 40, 8
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    src = tmp_path / "src.mp4"
    write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h, timescale=30 * 512,
                 delta=512, ctts=[512] * n, video_media_time=512)
    good = src.read_bytes()
    video_segmenter.configure(frame_buffers=False)
    video_segmenter.extract_segment(src, 0.2, 1.0, tmp_path / "warm.mp4")
    for tag in (b"stts", b"ctts", b"stsc"):
        at = good.find(tag)
        assert at > 0
        bad = bytearray(good)
        first_count = at + 4 + 4 + 4 + (4 if tag == b"stsc" else 0)       # version/flags, entry count, [first_chunk]
        bad[first_count] = 0xCE
        f = tmp_path / ("bad_%s.mp4" % tag.decode())
        f.write_bytes(bytes(bad))
        t0 = time.perf_counter()
        assert isinstance(probe_duration(f), float)
        video_segmenter.configure(frame_buffers=False)
        assert video_segmenter.extract_segment(f, 0.2, 1.0, tmp_path / "o.mp4") in (True, False)
        assert time.perf_counter() - t0 < 2.0, tag
    at = good.find(b"stsz")
    bad = bytearray(good)
    bad[at + 8:at + 16] = struct.pack(">II", 4, 0xFFFFFFF0)                # fixed size, four billion samples
    f = tmp_path / "bad_stsz.mp4"
    f.write_bytes(bytes(bad))
    t0 = time.perf_counter()
    assert isinstance(probe_duration(f), float)
    assert video_segmenter.extract_segment(f, 0.2, 1.0, tmp_path / "o2.mp4") is False
    assert time.perf_counter() - t0 < 2.0


def test_several_sample_entries_keep_their_chunks(tmp_path):
    """A track may use more than one sample entry (`stsd`), chosen per chunk by stsc's third column (files assembled
    from several encodes).  The cut copies all entries verbatim and every copied sample keeps its entry."""
    w, h, n, gop, fps = 96, 80, 90, 10, 30
    sps, pps, samples, keys, luma = _pcm_samples(w, h, n, gop)
    src = tmp_path / "src.mp4"
    meta = write_av_mp4(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h,
                        timescale=fps * 512, delta=512, video_chunk=5, second_entry_chunk=9)   # samples 45.. use entry 2
    movie = isobmff.read_movie(src)
    v = movie.video_track()
    assert v.desc is not None and v.desc[:45].tolist() == [1] * 45 and v.desc[45:].tolist() == [2] * 45
    assert struct.unpack_from(">I", v.stsd, 12)[0] == 2 and b"second entry" in v.stsd
    out = tmp_path / "cut.mp4"
    r = isobmff.cut_movie(movie, 1.0, 2.5, out)                      # pictures 30..75: both entries are in use
    assert (r.first, r.last) == (30, 75)
    cut = isobmff.read_movie(out)
    c = cut.video_track()
    assert c.stsd == v.stsd
    assert c.desc is not None and c.desc.tolist() == [1] * 15 + [2] * 30
    data = out.read_bytes()
    assert [data[int(o):int(o) + int(z)] for o, z in zip(c.offsets, c.sizes)] == meta["video_samples"][30:75]
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)                 # libavformat follows the entries, too
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    k = 30
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), luma[k]), k
        k += 1
    assert k == 75
    # a window inside the second entry's range: only entry 2 is referenced, the table still names it
    r2 = isobmff.cut_movie(movie, 2.0, 2.9, tmp_path / "late.mp4")
    late = isobmff.read_movie(tmp_path / "late.mp4").video_track()
    assert late.desc is not None and set(late.desc.tolist()) == {2} and late.n == r2.last - r2.first


@pytest.mark.parametrize("codec_id,private,rate,fourcc", [
    (b"A_MPEG/L3", b"", 44100.0, b"mp4a"),
    (b"A_FLAC", b"fLaC" + bytes([0x80, 0, 0, 34]) + bytes(10) + bytes([0x0A, 0xC4, 0x42, 0xF0]) + bytes(20), 44100.0, b"fLaC"),
    (b"A_VORBIS", b"\x02\x1e\x20" + bytes(60), 44100.0, None),
])
def test_matroska_audio_codecs_map_to_mp4_sample_entries(tmp_path, codec_id, private, rate, fourcc):
    """MP3 -> `mp4a` with object type 0x6B and no decoder-specific info, FLAC -> `fLaC` + dfLa (the CodecPrivate's
    metadata blocks); a codec without an MP4 mapping (Vorbis: ffmpeg's mov muxer refuses it, too) is left out, the video
    still arrives.  Packets are opaque to a stream copy, so the fixture's are arbitrary bytes."""
    w, h, n, gop, fps = 64, 48, 40, 10, 30
    sps, pps, samples, keys, _ = _pcm_samples(w, h, n, gop)
    packets = [bytes([0xFF, 0xFB, k & 0xFF]) + bytes((k + j) & 0xFF for j in range(50)) for k in range(50)]
    src = tmp_path / "a.mkv"
    write_mkv(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, width=w, height=h, fps=fps,
              opus_packets=packets, opus_ms=26, audio_codec=(codec_id, private, rate, 2))
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.0, 1.0, out) is True
    cut = isobmff.read_movie(out)
    if fourcc is None:
        assert [t.codec for t in cut.tracks] == [b"avc1"]
        return
    assert [t.codec for t in cut.tracks] == [b"avc1", fourcc]
    a = cut.tracks[1]
    data = out.read_bytes()
    got = [data[int(o):int(o) + int(z)] for o, z in zip(a.offsets, a.sizes)]
    assert got == packets[:len(got)] and len(got) >= 38                 # one second of 26 ms packets
    if fourcc == b"mp4a":
        assert b"esds" in a.stsd and bytes([0x04]) in a.stsd and b"\x6b\x15" in a.stsd and b"\x05\x80" not in a.stsd
    else:
        assert a.stsd.count(b"dfLa") == 1 and private[4:] in a.stsd
    assert len(_cv_frames(out)) == 30                                   # libavformat opens the file, video intact


def test_flv_source_is_stream_copied_into_mp4(tmp_path):
    """FLV (AVC + AAC): the cut is an MP4 whose `avc1` entry carries the file's AVCDecoderConfigurationRecord verbatim,
    whose `mp4a` entry carries the AudioSpecificConfig, with sample bytes unchanged, audio pre-roll behind the edit
    list, and pictures libavcodec decodes to the generator's."""
    from mp4_fixture import write_flv
    from oracle import scene_oracle
    w, h, n, gop, fps = 128, 96, 60, 10, 30
    sps, pps, samples, keys, luma = _pcm_samples(w, h, n, gop)
    frames = [bytes([0x21, k & 0xFF]) + bytes((k * 5 + j) & 0xFF for j in range(60 + k % 7)) for k in range(95)]   # ~2 s
    src = tmp_path / "a.flv"
    meta = write_flv(src, sps=sps, pps=pps, video_samples=samples, keyframes=keys, fps=fps, aac_frames=frames)
    assert abs(probe_duration(src) - 2.0) < 1e-3
    idx = container.probe(src)
    assert idx is not None and idx.extra["container"] == "flv" and idx.n_frames == n and (idx.width, idx.height) == (w, h)
    assert len(_cv_frames(src)) == n                          # libavformat reads the fixture as 60 pictures
    video_segmenter.configure(frame_buffers=False)
    out = tmp_path / "cut.mp4"
    assert video_segmenter.extract_segment(src, 0.7, 1.6, out) is True
    first, last = scene_oracle.frames_for_window(0.7, 1.6, n, fps, 1, np.nonzero(keys)[0], True)
    cut = isobmff.read_movie(out)
    assert [t.codec for t in cut.tracks] == [b"avc1", b"mp4a"]
    v, a = cut.tracks
    assert meta["avcc"] in v.stsd and meta["asc"] in a.stsd and b"esds" in a.stsd
    data = out.read_bytes()
    assert [data[int(o):int(o) + int(z)] for o, z in zip(v.offsets, v.sizes)] == meta["video_samples"][first:last]
    got = [data[int(o):int(o) + int(z)] for o, z in zip(a.offsets, a.sizes)]
    k0 = frames.index(got[0])
    assert got == frames[k0:k0 + len(got)]
    t_lo = first / fps
    assert k0 * 1024 / 48000.0 <= t_lo and (k0 + 3) * 1024 / 48000.0 > t_lo - 1e-9      # one frame of AAC pre-roll kept
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    k = first
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), luma[k]), k
        k += 1
    assert k == last
    assert abs(probe_duration(out) - (last - first) / fps) < 0.03
    # truncated download: whatever tags are complete still index; garbage does not raise
    (tmp_path / "t.flv").write_bytes(src.read_bytes()[: src.stat().st_size * 2 // 3])
    assert container.probe(tmp_path / "t.flv").n_frames < n
    (tmp_path / "g.flv").write_bytes(b"FLV\x01\x05\x00\x00\x00\x09" + bytes(range(200)))
    assert video_segmenter.extract_segment(tmp_path / "g.flv", 0.0, 1.0, tmp_path / "g.mp4") is False

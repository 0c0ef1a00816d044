"""End-to-end: extract_segment() on the GPU against the oracles (picture range, MP4 artefact, frame buffers,
SAD chain, scene cuts), plus SegmentIngestor range/lead-in behaviour."""
import json

import numpy as np
import pytest

from video_transformer_b200 import container, synth, video_segmenter
from video_transformer_b200.video_utils import probe_duration

pytestmark = pytest.mark.gpu


def _clip(tmp_path, w, h, n, gop, cuts, mp4=True):
    bs, meta = synth.make_testsrc_h264(w, h, n, fps=30, gop=gop, cuts=cuts)
    raw = tmp_path / "clip.h264"
    raw.write_bytes(bs)
    if not mp4:
        return raw, meta
    out = tmp_path / "clip.mp4"
    container.annexb_to_mp4(raw, out)
    return out, meta


def _expected(w, h, meta):
    scene, ref, out = 0, None, []
    cuts = set(meta["cuts"])
    for k in range(meta["n_frames"]):
        if k in cuts:
            scene += 1
        if k in meta["idr_frames"]:
            ref = tuple(np.maximum(p, 1) for p in synth.testsrc_frame(w, h, k, scene))
        out.append(ref)
    return out


@pytest.mark.parametrize("kind", ["mp4", "h264"])
@pytest.mark.parametrize("stream_copy", [True, False])
def test_extract_segment_gpu_matches_oracles(cuda, oracle_c, tmp_path, kind, stream_copy):
    from oracle import scene_oracle
    w, h, n, gop = 640, 480, 150, 15
    src, meta = _clip(tmp_path, w, h, n, gop, cuts=[37, 95], mp4=(kind == "mp4"))
    assert abs(probe_duration(src) - 5.0) < 1e-9
    video_segmenter.configure(target_height=240, batch_frames=16, scene_threshold=0.05)
    start, end = 1.2345, 4.0
    out = tmp_path / "segments" / "vid" / "segment_0001.mp4"
    assert video_segmenter.extract_segment(input_path=src, start=start, end=end, output_path=out,
                                           stream_copy=stream_copy) is True
    assert out.exists() and out.stat().st_size > 0
    first, last = scene_oracle.frames_for_window(start, end, n, 30, 1, meta["idr_frames"], stream_copy)
    side = json.loads(out.with_suffix(".json").read_text())
    assert (side["first_picture"], side["last_picture"]) == (first, last)
    assert side["frame_size"] == [320, 240] and side["frames"] == last - first
    # the MP4 artefact: same pictures, decodable from its first sample
    cut = container.probe(out)
    assert cut.n_frames == last - first and bool(cut.keyframe[0])
    cv2 = pytest.importorskip("cv2")
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    exp = _expected(w, h, meta)
    k = first
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), exp[k][0]), k
        k += 1
    assert k == last
    # frame buffers: bit-exact against the C oracle applied to the expected decoded pictures
    fb = 320 * 240 * 3 // 2
    frames = np.fromfile(out.with_suffix(".frames"), np.uint8).reshape(-1, fb)
    assert frames.shape[0] == last - first
    cache = {}
    for i, k in enumerate(range(first, last)):
        key = id(exp[k])
        if key not in cache:
            ey, eu, ev = oracle_c.scale_yuv420p(*exp[k], 320, 240, oracle_c.BICUBIC)
            cache[key] = np.concatenate([ey.reshape(-1), eu.reshape(-1), ev.reshape(-1)])
        assert np.array_equal(frames[i], cache[key]), k
    # SAD chain and cuts: integers exact, float64 scores bit-identical to the scalar oracle
    sads = [0 if k == 0 else oracle_c.sad_hist(exp[k][0], exp[k - 1][0])[0] for k in range(max(first - 1, 0), last)]
    if first == 0:
        sads[0] = 0
    sc = scene_oracle.scene_scores(sads, w, h)
    if first > 0:
        sads, sc = sads[1:], sc[1:]
    assert side["sad"] == sads
    assert [float(x).hex() for x in side["score"]] == [float(x).hex() for x in sc]
    exp_cuts = [first + t for t, v in enumerate(sc) if v > 0.05 and first + t > 0]
    assert side["cuts"] == exp_cuts
    assert set(c for c in meta["cuts"] if first < c < last) <= set(side["cuts"])
    video_segmenter.configure(target_height=720, batch_frames=32, scene_threshold=0.10)


def test_ingest_is_shard_invariant(cuda, tmp_path):
    """Boundaries must not depend on how pictures are sharded: per-shard SADs concatenate to the single-pass SADs."""
    from video_transformer_b200 import ingest
    src, meta = _clip(tmp_path, 320, 240, 120, 10, cuts=[33, 71])
    idx = container.probe(src)
    opts = ingest.IngestOptions(target_height=120, batch_frames=8, scene_threshold=0.05)
    whole = ingest.SegmentIngestor(idx, opts).run(0, 120, None)
    parts = [ingest.SegmentIngestor(idx, opts).run(a, b, None) for a, b in [(0, 33), (33, 34), (34, 95), (95, 120)]]
    assert np.array_equal(np.concatenate([p.sad for p in parts]), whole.sad)
    assert np.array_equal(np.concatenate([p.hist for p in parts]), whole.hist)
    assert np.array_equal(np.concatenate([p.scores for p in parts]), whole.scores)
    assert sorted(np.concatenate([p.cuts for p in parts]).tolist()) == whole.cuts.tolist()
    assert set(meta["cuts"]) <= set(whole.cuts.tolist())


def test_same_height_source_is_converted_not_resized(cuda, oracle_c, tmp_path):
    from video_transformer_b200 import ingest
    src, meta = _clip(tmp_path, 320, 240, 12, 4, cuts=[])
    idx = container.probe(src)
    got = []
    res = ingest.SegmentIngestor(idx, ingest.IngestOptions(target_height=240, batch_frames=5)).run(
        2, 11, lambda chunk, k0: got.append((k0, chunk.numpy().copy())))
    assert (res.out_width, res.out_height) == (320, 240)
    exp = _expected(320, 240, meta)
    k = 2
    for k0, chunk in got:
        assert k0 == k
        for row in chunk:
            y, u, v = exp[k]
            assert np.array_equal(row, np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)])), k
            k += 1
    assert k == 11


def test_config5_sampled_rgb_frames(cuda, oracle_c, tmp_path):
    """BASELINE.json configs[4]: 1 fps sampling + fixed-size RGB for upload.  Every picture is scored, only pictures
    whose index is a multiple of sample_every are converted; RGB is bit-exact against the libswscale-pinned oracle."""
    from video_transformer_b200 import ingest
    w, h, n = 320, 240, 75
    src, meta = _clip(tmp_path, w, h, n, 10, cuts=[41])
    idx = container.probe(src)
    got = []
    opts = ingest.IngestOptions(output="rgb24", rgb_size=(96, 96), sample_every=30, batch_frames=8, scene_threshold=0.05)
    eng = ingest.SegmentIngestor(idx, opts)
    res = eng.run(5, n, lambda chunk, k0: got.append((k0, chunk.numpy().copy())))
    assert (res.out_width, res.out_height, res.frame_bytes) == (96, 96, 96 * 96 * 3)
    assert [k0 for k0, _ in got] == [30, 60] and all(c.shape[0] == 1 for _, c in got)
    exp = _expected(w, h, meta)
    for k0, chunk in got:
        y, u, v = exp[k0]
        nv12 = synth.planar_to_nv12(y, u, v, eng.pitch)
        want = oracle_c.nv12_to_rgb24(nv12.reshape(-1), w, h, eng.pitch, 96, 96)
        assert np.array_equal(chunk[0].reshape(96, 96, 3), want), k0
    # scoring is unaffected by the sampling
    whole = ingest.SegmentIngestor(idx, ingest.IngestOptions(target_height=120, batch_frames=8, scene_threshold=0.05)).run(5, n, None)
    assert np.array_equal(res.sad, whole.sad) and res.cuts.tolist() == whole.cuts.tolist() and 41 in res.cuts.tolist()


def test_config3_4k_scene_cuts_scores_only(cuda, oracle_c, tmp_path):
    """BASELINE.json configs[2] shape (3840x2160, 60 fps), score-only pass: SADs are exact integers, scores are
    bit-identical float64 and the cut list equals the scalar oracle's on the same decoded pictures.
    (HEVC is not available on this pool: the clip is H.264 PCM-intra)"""
    from oracle import scene_oracle
    from video_transformer_b200 import ingest
    w, h, n = 3840, 2160, 36
    bs, meta = synth.make_testsrc_h264(w, h, n, fps=60, gop=12, cuts=[7, 20, 29])
    raw = tmp_path / "uhd.h264"
    raw.write_bytes(bs)
    idx = container.probe(raw)
    assert (idx.width, idx.height, idx.fps_num // idx.fps_den) == (w, h, 60)
    opts = ingest.IngestOptions(keep_frames=False, batch_frames=8, scene_threshold=0.10)
    res = ingest.SegmentIngestor(idx, opts).run(0, n, None)
    exp = _expected(w, h, meta)
    sads = [0] + [oracle_c.sad_hist(exp[k][0], exp[k - 1][0])[0] for k in range(1, n)]
    assert res.sad.tolist() == sads
    sc = scene_oracle.scene_scores(sads, w, h)
    assert [float(x).hex() for x in res.scores] == [float(x).hex() for x in sc]
    assert res.cuts.tolist() == scene_oracle.select_cuts(sc, 0.10)
    assert set(meta["cuts"]) <= set(res.cuts.tolist())
    for k in (0, 7, 35):
        assert np.array_equal(res.hist[k], oracle_c.sad_hist(exp[k][0], None)[1])


def test_device_sink_hands_over_the_same_frames_without_a_host_copy(cuda, tmp_path):
    """A consumer on the GPU takes each batch on the compute stream; frames equal the host-sink frames and the pass
    copies only scores to the host."""
    import torch
    from video_transformer_b200 import ingest
    src, meta = _clip(tmp_path, 640, 480, 50, 10, cuts=[23])
    idx = container.probe(src)
    opts = ingest.IngestOptions(target_height=240, batch_frames=8, scene_threshold=0.05)
    host = []
    ref = ingest.SegmentIngestor(idx, opts).run(3, 47, lambda chunk, k0: host.append((k0, chunk.numpy().copy())))
    eng = ingest.SegmentIngestor(idx, opts)
    dev = torch.zeros((44, eng.frame_bytes), dtype=torch.uint8, device="cuda")
    seen = []

    def on_device(chunk, k0):
        assert chunk.is_cuda
        dev[k0 - 3:k0 - 3 + chunk.shape[0]].copy_(chunk, non_blocking=True)   # enqueued on the compute stream
        seen.append((k0, chunk.shape[0]))

    res = eng.run(3, 47, None, device_sink=on_device)
    torch.cuda.synchronize()
    assert [k for k, _ in seen] == [k for k, _ in host] and sum(n for _, n in seen) == 44
    assert np.array_equal(dev.cpu().numpy(), np.concatenate([c for _, c in host]))
    assert np.array_equal(res.sad, ref.sad) and res.cuts.tolist() == ref.cuts.tolist()
    assert eng.d2h_bytes == 50 * 0 + sum(min(8, 47 - b) for b in range(0, 47, 8)) * (8 + 1024)   # scores only


@pytest.mark.parametrize("where", ["shm", "disk"])
@pytest.mark.parametrize("sample_every", [1, 7])
def test_frames_land_directly_in_the_file_and_landing_files_are_recycled(cuda, tmp_path, where, sample_every):
    """K5: on tmpfs the copy engine writes into a registered mapping of the `.frames` file (no host copy); elsewhere a
    writer thread drains the pinned ring.  Either way the file equals the frames the engine API delivers, and a
    replaced artefact lands in the same registered pages (landing_recycled)."""
    import os
    import shutil
    import tempfile
    from video_transformer_b200 import ingest, landing
    src, meta = _clip(tmp_path, 640, 480, 90, 10, cuts=[23, 61])
    idx = container.probe(src)
    opts = ingest.IngestOptions(target_height=240, batch_frames=16, scene_threshold=0.05, sample_every=sample_every)
    first, last = 10, 90
    ref = []
    ingest.SegmentIngestor(idx, opts).run(first, last, lambda chunk, k0: ref.append(chunk.numpy().copy()))
    ref = np.concatenate(ref)
    if where == "shm":
        if not os.path.isdir("/dev/shm"):
            pytest.skip("no /dev/shm")
        out_dir = tempfile.mkdtemp(prefix="vt_landing_test_", dir="/dev/shm")
    else:
        out_dir = str(tmp_path / "out")
    try:
        video_segmenter.configure(target_height=240, batch_frames=16, scene_threshold=0.05, sample_every=sample_every)
        out = os.path.join(out_dir, "segment_0000.mp4")
        sides = []
        for rep in range(3):
            assert video_segmenter.extract_segment(src, first / 30.0, last / 30.0, out) is True
            side = json.loads(open(out[:-4] + ".json").read())
            sides.append(side)
            frames = np.fromfile(out[:-4] + ".frames", np.uint8).reshape(-1, side["frame_bytes"])
            assert side["frames"] == ref.shape[0] == frames.shape[0]
            assert np.array_equal(frames, ref), (where, rep)
        if where == "shm" and sides[0]["landing"] == "direct":
            assert [s["landing_recycled"] for s in sides] == [False, True, True]
            assert landing.stats()["files"] >= 1
            # the consumer deletes the artefact: its landing file becomes free, the next segment reuses it
            os.unlink(out[:-4] + ".frames")
            out2 = os.path.join(out_dir, "segment_0001.mp4")
            assert video_segmenter.extract_segment(src, first / 30.0, last / 30.0, out2) is True
            assert json.loads(open(out2[:-4] + ".json").read())["landing_recycled"] is True
            assert np.array_equal(np.fromfile(out2[:-4] + ".frames", np.uint8).reshape(ref.shape), ref)
        else:
            assert sides[0]["landing"] in ("staged", "direct")
    finally:
        video_segmenter.configure(target_height=720, batch_frames=32, scene_threshold=0.10, sample_every=1)
        landing.release_all()
        shutil.rmtree(out_dir, ignore_errors=True)


def test_pixel_pass_reads_matroska_and_flv_sources(cuda, oracle_c, tmp_path):
    """The decode front end indexes slice NALs by file offset, so a PCM-intra AVC track inside Matroska or FLV goes
    through the same GPU pass as one inside MP4: same frame buffers, same SADs."""
    import struct
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).parent))
    from mp4_fixture import write_flv, write_mkv
    w, h, n, gop = 320, 240, 40, 8
    wr = synth.H264PcmWriter(w, h, 30, 1)
    samples, keys = [], []
    for k in range(n):
        if k % gop == 0:
            samples.append([wr.idr(*synth.testsrc_frame(w, h, k, k // gop), with_params=False)[4:]])
            keys.append(True)
        else:
            samples.append([wr.skip()[4:]])
            keys.append(False)
    mkv = tmp_path / "clip.mkv"
    write_mkv(mkv, sps=wr._sps[4:], pps=wr._pps[4:], video_samples=samples, keyframes=keys, width=w, height=h, fps=30)
    flv = tmp_path / "clip.flv"
    write_flv(flv, sps=wr._sps[4:], pps=wr._pps[4:], video_samples=samples, keyframes=keys, fps=30)
    mp4 = tmp_path / "clip.mp4"
    sys.path.insert(0, str(tmp_path))
    from mp4_fixture import write_av_mp4
    write_av_mp4(mp4, sps=wr._sps[4:], pps=wr._pps[4:], video_samples=samples, keyframes=keys, width=w, height=h,
                 timescale=30000, delta=1000)
    saved = video_segmenter.configure()
    try:
        video_segmenter.configure(target_height=120, batch_frames=8, scene_threshold=0.05)
        outs = []
        for src in (mkv, flv, mp4):
            out = tmp_path / (src.suffix[1:]) / "segment_0000.mp4"
            assert video_segmenter.extract_segment(src, 0.3, 1.2, out) is True
            side = json.loads(out.with_suffix(".json").read_text())
            outs.append((side, np.fromfile(out.with_suffix(".frames"), np.uint8)))
        for other in outs[1:]:
            assert outs[0][0]["frames"] == other[0]["frames"] > 0
            assert outs[0][0]["sad"] == other[0]["sad"] and outs[0][0]["cuts"] == other[0]["cuts"]
            assert np.array_equal(outs[0][1], other[1])
    finally:
        video_segmenter.configure(**saved)


def test_direct_and_staged_copy_in_deliver_the_same_pass(cuda, tmp_path, monkeypatch):
    """The bitstream reaches the device either straight from a page-locked mapping of the source file (files of
    VT_INGEST_DIRECT_MIN_MB and more) or through the pinned staging buffers (small files): same frames, same SADs."""
    from video_transformer_b200 import ingest
    w, h, n = 320, 240, 75
    src, _meta = _clip(tmp_path, w, h, n, gop=10, cuts=[31], mp4=True)
    idx = container.probe(src)
    got = {}
    for mode, min_mb in (("staged", "128"), ("direct", "0")):
        monkeypatch.setenv("VT_INGEST_DIRECT_MIN_MB", min_mb)
        eng = ingest.SegmentIngestor(idx, ingest.IngestOptions(target_height=120, batch_frames=16))
        assert (eng._src_map is not None) == (mode == "direct"), mode
        frames = np.zeros((n - 12, eng.frame_bytes), np.uint8)

        def sink(chunk, first_picture, frames=frames):
            frames[first_picture - 12:first_picture - 12 + chunk.shape[0]] = chunk.numpy()

        res = eng.run(12, n, sink=sink)
        got[mode] = (frames, res.sad.copy(), res.hist.copy(), res.cuts.copy())
        eng.release()
    for a, b in zip(got["staged"], got["direct"]):
        assert np.array_equal(a, b)
    assert got["direct"][0].any()

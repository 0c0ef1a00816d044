"""Upload-size reducer: the drop-in for `ContentAnalyzer._compress_video_for_upload`
(/root/reference/src/analyzer/content_analyzer.py:167-236).

The reference shells out to `ffmpeg -i IN -vf scale=-2:360 -c:v libx264 -crf 28 ...` and returns the path of
`compressed_<name>` beside the input -- or the ORIGINAL path when the file is small (<= 30 MiB), when the compressed
file already exists (cache), or when anything fails.  This module keeps that contract (same gate, same name, same
"never raises, fall back to the input" behaviour) and replaces the decode -> swscale -> encode chain with the GPU pass:

    decode (K0) -> SAD/hist (K3) -> libswscale-exact bicubic to height 360, width rounded to even (K2) -> frames (K5)

A B200 has no NVENC and the image has no software H.264 encoder, so the artefact is not libx264 output.  It is a
Motion-JPEG MP4: the downscaled pictures sampled at `sample_fps` (default 1 picture per second, the granularity the
remote model consumes), each encoded on the GPU as a baseline JPEG (csrc/vt_jpeg.cu: IJG integer DCT, Annex-K Huffman
tables, quality 75 by default; limited-range video samples are expanded to JFIF's full range), in a `jpeg` video track
at the sampled rate -- readable by libavformat/libavcodec and every MP4 player -- with the source's AUDIO TRACKS COPIED
VERBATIM (the reference re-encodes audio to AAC 64k; a stream copy keeps the speech the lecture analysis depends on
without needing an audio encoder).  Scene scores of every picture land in `compressed_<stem>.json`.  When the artefact
would not be smaller than the input, the input path is returned: the function exists to reduce upload size and must
never enlarge it.
"""
from __future__ import annotations

import json
import logging
from pathlib import Path

import numpy as np

log = logging.getLogger(__name__)

MAX_SIZE_MB = 30            # content_analyzer.py:174
TARGET_HEIGHT = 360         # `scale=-2:360`, content_analyzer.py:199
JPEG_QUALITY = 75           # IJG quality of the Motion-JPEG pictures (the reference's knob is x264's CRF 28)


def compressed_path_for(video_path: str | Path) -> Path:
    video_path = Path(video_path)
    return video_path.parent / ("compressed_%s" % video_path.name)     # content_analyzer.py:186


def compress_video_for_upload(video_path: str | Path, *, max_size_mb: float = MAX_SIZE_MB,
                              target_height: int = TARGET_HEIGHT, sample_fps: float = 1.0,
                              device: str = "cuda") -> Path:
    """Returns the path to upload: `compressed_<name>` when a smaller artefact exists or could be made, else the input."""
    video_path = Path(video_path)
    size_mb = video_path.stat().st_size / (1024 * 1024)                # a missing input raises, as in the reference
    if size_mb <= max_size_mb:
        log.info("event=upload_reduce_skip size_mb=%.1f limit_mb=%s", size_mb, max_size_mb)
        return video_path
    out = compressed_path_for(video_path)
    if out.exists() and out.stat().st_size > 0:                        # content_analyzer.py:189-191
        log.info("event=upload_reduce_cached file=%s", out.name)
        return out
    try:
        ok = _reduce(video_path, out, target_height, sample_fps, device)
    except Exception as exc:  # noqa: BLE001 -- the reference swallows every failure and uploads the original
        log.warning("event=upload_reduce_failed error=%s", str(exc)[:200])
        ok = False
    if not ok:
        for p in (out, out.with_suffix(".json")):
            if p.exists():
                p.unlink()
        return video_path
    log.info("event=upload_reduce_done size_mb=%.1f new_size_mb=%.1f", size_mb, out.stat().st_size / (1024 * 1024))
    return out


def _reduce(src: Path, out: Path, target_height: int, sample_fps: float, device: str, quality: int = JPEG_QUALITY) -> bool:
    import torch
    from . import container, ingest, isobmff, ops
    idx = container.probe(src)
    if idx is None or idx.n_frames == 0 or idx.fps_num <= 0:
        return False
    if idx.kind == "mp4" and not (idx.extra.get("decodable") and container.classify_pcm(idx)):
        raise RuntimeError("video track %r cannot be decoded by this build (K0 handles PCM-intra H.264; NVDEC is not "
                           "available on this host)" % idx.extra.get("codec"))
    fps = idx.fps_num / idx.fps_den
    every = max(1, int(round(fps / sample_fps))) if sample_fps > 0 else 1
    opts = ingest.IngestOptions(target_height=target_height, sample_every=every, device=device)
    eng = ingest.SegmentIngestor(idx, opts)
    dw, dh = eng.out_w, eng.out_h
    jpeg = ops.JpegPlan(dw, dh, quality, expand_range=True)
    samples: list[bytes] = []
    pending: list = []

    def flush():
        for data, offsets, status in pending:
            off = offsets.cpu().numpy()                       # synchronises the compute stream up to this batch
            if int(status.cpu()[0]) != 0:
                raise RuntimeError("JPEG encoder refused a batch (status %d)" % int(status.cpu()[0]))
            host = data[: int(off[-1])].cpu().numpy()
            samples.extend(host[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1))
        pending.clear()

    def on_device(chunk, first_picture):
        # called with the compute stream current, right after the kernels that produced `chunk` were enqueued: the
        # encoder's launches are ordered after them and before the buffers are reused
        pending.append(jpeg.encode(chunk))
        if len(pending) >= 32:
            flush()

    try:
        with torch.cuda.device(eng.dev):
            res = eng.run(0, idx.n_frames, None, device_sink=on_device)
            flush()
    finally:
        jpeg.close()
    if not samples:
        return False
    g = int(np.gcd(idx.fps_num, idx.fps_den * every))
    timescale, delta = int(idx.fps_num // g), int(idx.fps_den * every // g)
    movie = idx.extra.get("movie")
    mts = (movie.timescale if movie is not None else 0) or 1000
    audio = [t for t in movie.tracks if t.handler == b"soun" and t.n] if movie is not None else []
    vid = isobmff.make_video_track(max([t.track_id for t in audio] + [0]) + 1, b"jpeg", dw, dh, timescale,
                                   b"Photo - JPEG")
    plans = [isobmff.plan_memory_track(vid, samples, [delta] * len(samples))]
    plans += [isobmff.plan_whole_track(t, mts) for t in audio]
    isobmff.write_plans(out, plans, mts, b"", movie.path if movie is not None else None)
    out.with_suffix(".json").write_text(json.dumps({
        "source": str(src), "source_size": [idx.width, idx.height], "frame_size": [dw, dh], "codec": "mjpeg",
        "jpeg_quality": quality, "sample_every": every, "frames": len(samples), "audio_tracks": len(audio),
        "cuts": [int(c) for c in res.cuts], "sad": [int(s) for s in res.sad],
        "score": [float(s) for s in res.scores]}), encoding="utf-8")
    return out.stat().st_size < src.stat().st_size

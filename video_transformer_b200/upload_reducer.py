"""Upload-size reducer: the drop-in for `ContentAnalyzer._compress_video_for_upload`
(/root/reference/src/analyzer/content_analyzer.py:167-236).

The reference shells out to `ffmpeg -i IN -vf scale=-2:360 -c:v libx264 -crf 28 ...` and returns the path of
`compressed_<name>` beside the input -- or the ORIGINAL path when the file is small (<= 30 MiB), when the compressed
file already exists (cache), or when anything fails.  This module keeps that contract (same gate, same name, same
"never raises, fall back to the input" behaviour) and replaces the decode -> swscale -> encode chain with the GPU pass:

    decode (K0) -> SAD/hist (K3) -> libswscale-exact bicubic to height 360, width rounded to even (K2) -> frames (K5)

A B200 has no NVENC, so the artefact is not libx264 output.  It is a valid H.264 MP4 whose pictures are the downscaled
frames sampled at `sample_fps` (default 1 picture per second, the granularity the remote model consumes), each
written as an I_PCM IDR picture -- decodable by any H.264 decoder, bit-identical to the scaled frames except that
PCM samples cannot be 0 and are raised to 1 -- plus the exact frames and scene scores in `compressed_<stem>.frames`
/ `.json`.  When that artefact would not be smaller than the input, the input path is returned: the function exists
to reduce upload size and must never enlarge it.
"""
from __future__ import annotations

import json
import logging
from pathlib import Path

import numpy as np

log = logging.getLogger(__name__)

MAX_SIZE_MB = 30            # content_analyzer.py:174
TARGET_HEIGHT = 360         # `scale=-2:360`, content_analyzer.py:199


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
        for p in (out, out.with_suffix(".frames"), out.with_suffix(".json")):
            if p.exists():
                p.unlink()
        return video_path
    log.info("event=upload_reduce_done size_mb=%.1f new_size_mb=%.1f", size_mb, out.stat().st_size / (1024 * 1024))
    return out


def _reduce(src: Path, out: Path, target_height: int, sample_fps: float, device: str) -> bool:
    from . import container, ingest, synth
    idx = container.probe(src)
    if idx is None or idx.n_frames == 0 or idx.fps_num <= 0:
        return False
    fps = idx.fps_num / idx.fps_den
    every = max(1, int(round(fps / sample_fps))) if sample_fps > 0 else 1
    opts = ingest.IngestOptions(target_height=target_height, sample_every=every, device=device)
    eng = ingest.SegmentIngestor(idx, opts)
    dw, dh = eng.out_w, eng.out_h
    n_keep = (idx.n_frames + every - 1) // every
    # the I_PCM picture is a fixed 384 bytes per macroblock plus a few header bytes: decide before running the pass
    est = n_keep * (((dw + 15) // 16) * ((dh + 15) // 16) * 384 + 64)
    if est >= src.stat().st_size:
        log.info("event=upload_reduce_not_smaller est_bytes=%d input_bytes=%d", est, src.stat().st_size)
        return False
    wr = synth.H264PcmWriter(dw, dh, max(1, int(round(fps))), every)   # picture k of the output shows at k*every/fps
    start = len(synth._START)
    samples: list[bytes] = []
    ysz, csz = dw * dh, (dw // 2) * (dh // 2)
    frames_file = open(out.with_suffix(".frames"), "wb")

    def sink(chunk, first_picture):
        a = chunk.numpy()
        frames_file.write(a.tobytes())
        for row in a:
            y = row[:ysz].reshape(dh, dw)
            u = row[ysz:ysz + csz].reshape(dh // 2, dw // 2)
            v = row[ysz + csz:ysz + 2 * csz].reshape(dh // 2, dw // 2)
            samples.append(wr.idr(y, u, v, with_params=False)[start:])

    try:
        res = eng.run(0, idx.n_frames, sink)
    finally:
        frames_file.close()
    if not samples:
        return False
    g = np.gcd(idx.fps_num, idx.fps_den * every)
    container.write_mp4(out, sps=wr._sps[start:], pps=wr._pps[start:], samples=samples, width=dw, height=dh,
                        fps_num=int(idx.fps_num // g), fps_den=int(idx.fps_den * every // g),
                        keyframes=[True] * len(samples))
    out.with_suffix(".json").write_text(json.dumps({
        "source": str(src), "source_size": [idx.width, idx.height], "frame_size": [dw, dh], "pixel_format": "yuv420p",
        "sample_every": every, "frames": len(samples), "frame_bytes": res.frame_bytes,
        "cuts": [int(c) for c in res.cuts], "sad": [int(s) for s in res.sad],
        "score": [float(s) for s in res.scores]}), encoding="utf-8")
    return out.stat().st_size < src.stat().st_size

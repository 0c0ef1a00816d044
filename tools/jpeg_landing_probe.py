"""What would a COMPRESSED landing format buy?    Runs the upload reducer's path on the
bench clip with EVERY picture kept (decode -> SAD/hist -> 720p bicubic -> baseline JPEG on the GPU -> Motion-JPEG MP4),
i.e. the frames leave the GPU as ~100 KB JPEGs instead of 1.38 MB of raw YUV.

    gpurun -- python tools/jpeg_landing_probe.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from video_transformer_b200 import container, upload_reducer  # noqa: E402


def main():
    import shutil
    from pathlib import Path
    work = "/dev/shm/vt_jpegland_%d" % os.getpid()
    os.makedirs(work, exist_ok=True)
    raw, mp4 = os.path.join(work, "clip.h264"), os.path.join(work, "clip.mp4")
    n = bench.CLIP_FRAMES
    bench.make_clip(n, raw)
    container.annexb_to_mp4(raw, mp4)
    for target in (720, 360):
        for rep in range(3):
            out = Path(work) / ("mjpeg_%d_%d.mp4" % (target, rep))
            t0 = time.perf_counter()
            upload_reducer._reduce(Path(mp4), out, target, float(bench.FPS), "cuda")
            dt = time.perf_counter() - t0
            print("target %dp  pass %d: %d pictures in %.3f s = %.0f pictures/s; MJPEG file %.1f MB (%.1f KB per picture)" % (
                target, rep, n, dt, n / dt, out.stat().st_size / 1e6, out.stat().st_size / n / 1e3), flush=True)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

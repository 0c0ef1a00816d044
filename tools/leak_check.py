"""Many extract_segment calls in one process: resident memory, open descriptors, arena size and device memory must stay flat.

    gpurun -- python tools/leak_check.py [calls]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from video_transformer_b200 import container, landing, video_segmenter  # noqa: E402


def rss_mb():
    for line in open("/proc/self/status"):
        if line.startswith("VmRSS"):
            return int(line.split()[1]) / 1024.0
    return 0.0


def main():
    import torch
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    work = "/dev/shm/vt_leak_%d" % os.getpid()
    os.makedirs(work, exist_ok=True)
    raw, mp4 = os.path.join(work, "clip.h264"), os.path.join(work, "clip.mp4")
    bench.make_clip(1920, raw)
    container.annexb_to_mp4(raw, mp4)
    other = os.path.join(work, "other.mp4")                      # a second source: engines and mappings get replaced
    import shutil
    shutil.copyfile(mp4, other)
    video_segmenter.configure(target_height=720, frame_buffers=True)
    t0 = time.perf_counter()
    for k in range(calls):
        src = mp4 if (k // 10) % 2 == 0 else other
        a = (k % 6) * 8.0
        out = os.path.join(work, "seg_%d.mp4" % (k % 3))
        assert video_segmenter.extract_segment(input_path=src, start=a, end=a + 16.0 + (k % 4), output_path=out)
        for ext in (".frames", ".mp4", ".json"):                 # the consumer
            os.unlink(out[:-4] + ext)
        if k % 50 == 0 or k == calls - 1:
            print("call %4d  rss %7.1f MB  fds %3d  arena %s  cuda %6.1f MB reserved  %.1f s" % (
                k, rss_mb(), len(os.listdir("/proc/self/fd")), landing.stats(),
                torch.cuda.memory_reserved() / 1e6, time.perf_counter() - t0), flush=True)
    landing.release_all()
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

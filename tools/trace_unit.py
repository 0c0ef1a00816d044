"""Per-batch timeline of one warm extract_segment call: where do the three streams of the pipeline wait?

    gpurun -- python tools/trace_unit.py            (writes the clip to /dev/shm, runs 4 units, prints the last)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VT_INGEST_TRACE"] = "1"

import bench  # noqa: E402
from video_transformer_b200 import video_segmenter  # noqa: E402


def main():
    import torch
    work = "/dev/shm/vt_trace_%d" % os.getpid()
    os.makedirs(work, exist_ok=True)
    from video_transformer_b200 import container
    raw, mp4 = os.path.join(work, "clip.h264"), os.path.join(work, "clip.mp4")
    bench.make_clip(bench.CLIP_FRAMES, raw)
    container.annexb_to_mp4(raw, mp4)
    unit_s = bench.UNIT_PICTURES / bench.FPS
    video_segmenter.configure(target_height=720, frame_buffers=True)
    out = os.path.join(work, "seg.mp4")
    for j in range(4):
        t0 = time.perf_counter()
        assert video_segmenter.extract_segment(input_path=mp4, start=(j % 2) * unit_s, end=(j % 2 + 1) * unit_s,
                                               output_path=out, stream_copy=True)
        wall = time.perf_counter() - t0
    eng = video_segmenter._ENGINE_CACHE["engine"][1]
    tr = eng.last_trace
    print("wall %.2f ms; timings %s" % (wall * 1e3, {k: round(v * 1e3, 2) for k, v in video_segmenter.LAST_TIMINGS.items()}))
    print("host ms since run(): enter, loop start, batch 1 issued, last batch issued, drained:", eng.last_host_ms)
    print("batch  h2d[start,end]   cmp[start,end]   d2h[start,end]   d2h_ms  gap_before_d2h")
    prev_end = None
    busy = 0.0
    for i, r in enumerate(tr):
        gap = (r[4] - prev_end) if prev_end is not None else r[4]
        busy += r[5] - r[4]
        print("%3d   %7.2f %7.2f   %7.2f %7.2f   %7.2f %7.2f   %6.2f  %6.2f" % (i, r[0], r[1], r[2], r[3], r[4], r[5],
                                                                            r[5] - r[4], gap))
        prev_end = r[5]
    print("d2h busy %.2f ms of %.2f ms (%.1f %%)" % (busy, tr[-1][5], 100 * busy / tr[-1][5]))
    torch.cuda.synchronize()
    # how long does the FIRST H2D after an idle gap take?  (a) from the engine's page-locked file mapping, (b) pinned
    import ctypes
    from video_transformer_b200._lib import lib
    L = lib()
    n = 9 << 20
    dev_buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    st = torch.cuda.Stream()
    base = eng._src_map[2] if eng._src_map is not None else None
    for idle_ms in (0, 1, 3, 10, 30):
        for kind in ("mapping", "pinned"):
            if kind == "mapping" and base is None:
                continue
            res = []
            for rep in range(3):
                torch.cuda.synchronize()
                time.sleep(idle_ms * 1e-3)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                with torch.cuda.stream(st):
                    a.record(st)
                    if kind == "mapping":
                        L.vt_copy_to_device_async(ctypes.c_void_p(dev_buf.data_ptr()),
                                                  ctypes.c_void_p(base + (rep + 1) * (64 << 20)), n,
                                                  ctypes.c_void_p(st.cuda_stream))
                    else:
                        dev_buf.copy_(pinned, non_blocking=True)
                    b.record(st)
                t1 = time.perf_counter()
                st.synchronize()
                t2 = time.perf_counter()
                res.append((round(a.elapsed_time(b), 3), round((t1 - t0) * 1e3, 3), round((t2 - t0) * 1e3, 3)))
            print("idle %2d ms  %-8s  (device ms, enqueue ms, wall ms) x3: %s" % (idle_ms, kind, res))
    import shutil
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

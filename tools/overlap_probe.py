"""Does the decode of pass p+1 overlap with score + scale of pass p when they sit on two streams?
(decode is a DRAM-bound re-layout, score is bound by the shared-memory atomic unit, scale by the multiply pipe.)

    gpurun -- python tools/overlap_probe.py
"""
import os
import sys
import tempfile
from ctypes import c_void_p

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from video_transformer_b200 import _lib, container, ingest  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    work = tempfile.mkdtemp(prefix="vt_ovl_", dir="/dev/shm")
    raw = os.path.join(work, "clip.h264")
    n_clip = 1024
    bench.make_clip(n_clip, raw)
    idx = container.probe(raw)
    eng = ingest.SegmentIngestor(idx, ingest.IngestOptions(target_height=720, batch_frames=64, device=str(dev)))
    fb = eng.frame_bytes
    host = np.memmap(raw, dtype=np.uint8, mode="r")
    bs_dev = torch.zeros(host.size + 64, dtype=torch.uint8, device=dev)
    bs_dev[:host.size].copy_(torch.from_numpy(np.array(host)))
    F = 256
    n_chunks = n_clip // F
    NB = 2
    surf = [torch.empty((F, eng.rows, eng.pitch), dtype=torch.uint8, device=dev) for _ in range(NB)]
    out = [torch.empty((F, fb), dtype=torch.uint8, device=dev) for _ in range(NB)]
    sad = [torch.empty(F, dtype=torch.int64, device=dev) for _ in range(NB)]
    hist = [torch.empty((F, 256), dtype=torch.int32, device=dev) for _ in range(NB)]
    sA, sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def decode(p, st):
        pay = eng.payload[(p % n_chunks) * F:(p % n_chunks) * F + F].copy()
        _lib.check(L.vt_h264_pcm_decode(c_void_p(bs_dev.data_ptr()), pay.ctypes.data, F, eng.w, eng.h, None,
                                        c_void_p(surf[p % NB].data_ptr()), eng.pitch, eng.surface_bytes,
                                        c_void_p(st.cuda_stream)))

    def rest(p, st):
        b = p % NB
        _lib.check(L.vt_sad_hist_u8(c_void_p(surf[b].data_ptr()), eng.pitch, eng.surface_bytes, eng.w, eng.h, None, F,
                                    c_void_p(sad[b].data_ptr()), c_void_p(hist[b].data_ptr()), c_void_p(st.cuda_stream)))
        _lib.check(L.vt_scale_nv12_to_yuv420p(eng.plan._h, c_void_p(surf[b].data_ptr()), eng.pitch, eng.surface_bytes,
                                              c_void_p(out[b].data_ptr()), fb, F, c_void_p(st.cuda_stream)))

    def run(mode, passes=64):
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dec_done = [torch.cuda.Event() for _ in range(passes)]
        rest_done = [torch.cuda.Event() for _ in range(passes)]
        t0.record(sA)
        sB.wait_event(t0)
        for p in range(passes):
            if mode == "serial":
                decode(p, sA)
                rest(p, sA)
            else:
                # decode on stream B (it may run ahead by one buffer), score + scale on stream A
                if p >= NB:
                    sB.wait_event(rest_done[p - NB])        # the surface buffer is free again
                decode(p, sB)
                dec_done[p].record(sB)
                sA.wait_event(dec_done[p])
                rest(p, sA)
                rest_done[p].record(sA)
        sA.wait_stream(sB)
        t1.record(sA)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / passes

    ref = None
    for mode in ("serial", "two-stream", "serial", "two-stream"):
        run(mode, 8)
        ms = run(mode)
        print("%-11s %.4f ms per 256-picture pass = %.0f pictures/s" % (mode, ms, F / ms * 1e3), flush=True)
    # parity of the overlapped arrangement: same outputs as serial for one pass
    run("serial", 2)
    a = (out[1].clone(), sad[1].clone(), hist[1].clone())
    run("two-stream", 2)
    print("outputs equal:", bool((a[0] == out[1]).all() and (a[1] == sad[1]).all() and (a[2] == hist[1]).all()))
    import shutil
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

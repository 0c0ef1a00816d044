#!/usr/bin/env python
"""bench.py -- 720p output frames/s of the long-video ingest path (BASELINE.json metric) on N B200s.

Workload (N=1): BASELINE.json configs[1], "2-hour synthetic 1080p30 H.264 -> 720p downscale + segment", run as a
bounded stream of GOP-aligned steps: one step = 256 pictures of a synthetic 1920x1080@30 testsrc clip (I_PCM IDR
every 30 pictures + all-skip P pictures, see video_transformer_b200/synth.py; entropy decoding cost is NOT
representative of real CABAC content) through decode -> SAD/histogram -> swscale-exact bicubic 1280x720 YUV420P.
`value`  : kernels only, bitstream already resident in HBM (decode + score + scale per step, CUDA events).
`e2e`    : SegmentIngestor.run() from host bytes to pinned host frame buffers, H2D/D2H inside the timed region.
N > 1    : one process per GPU (torchrun), every rank ingests its own GOP-aligned shard, no data-path
           collective (weak scaling); torch.distributed is used for the barrier and the max-over-ranks only.
`--impl reference`: the same work on the host cores (libswscale bicubic + OpenCV SAD/histogram + C oracle
           PCM re-layout, all cores), because the literal reference path (ffmpeg child processes) cannot run
           here: the image has no ffmpeg binary.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SRC_W, SRC_H, FPS, GOP = 1920, 1080, 30, 30
OUT_H = 720
STEP_FRAMES = 256
METRIC = "720p_output_frames_per_sec"
UNIT = "frames/s"
# algorithmic bytes per picture (SURVEY.md section 8d / BASELINE.md section 3)
BYTES_SCALE = 3110400 + 1382400          # 1080p NV12 in, 720p YUV420P out
BYTES_SCORE = 2 * SRC_W * SRC_H          # cur + prev luma
BYTES_DECODE = 2 * 3110400               # samples in, NV12 surface out


def load_traffic():
    """DRAM bytes per picture per kernel from the committed ncu capture (profiles/r01_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p))["per_picture_bytes"]
    except Exception:  # noqa: BLE001
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- host-core arm -----------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(path, payload, keyframe, w, h, dw, dh):
    import cv2  # noqa: F401
    from oracle import coracle, ffsws
    _W.update(buf=np.memmap(path, dtype=np.uint8, mode="r"), payload=payload, key=keyframe, w=w, h=h, dw=dw, dh=dh,
              coracle=coracle, ffsws=ffsws, sws=ffsws.available())
    try:
        cv2.setNumThreads(1)
    except Exception:  # noqa: BLE001
        pass


def _cpu_worker(rng):
    """Decode + score + downscale pictures [a,b) on one host core.  Returns (frames, checksum)."""
    import cv2
    a, b = rng
    W = _W
    co, ff = W["coracle"], W["ffsws"]
    w, h, dw, dh = W["w"], W["h"], W["dw"], W["dh"]
    out = np.empty((dw * dh * 3 // 2,), np.uint8)     # the segment frame buffer slot being filled
    prev_y = None
    cur = None
    last_payload = None
    acc = 0
    if a > 0:                                          # predecessor of the task's first picture, for its SAD
        last_payload = int(W["payload"][a - 1])
        cur = co.pcm_picture_to_yuv420p(W["buf"], last_payload, w, h)
        prev_y = cur[0]
    for k in range(a, b):
        p = int(W["payload"][k])
        if p != last_payload:                          # IDR: re-layout the PCM samples; skip pictures repeat
            cur = co.pcm_picture_to_yuv420p(W["buf"], p, w, h)
            last_payload = p
        y, u, v = cur
        if prev_y is not None:
            acc += int(cv2.norm(y, prev_y, cv2.NORM_L1))
        hist = cv2.calcHist([y], [0], None, [256], [0, 256])
        acc += int(hist[16, 0])
        if W["sws"]:
            sy, su, sv = ff.scale_yuv420p(y, u, v, dw, dh, ff.SWS_BICUBIC)
        else:
            sy, su, sv = co.scale_yuv420p(y, u, v, dw, dh, co.BICUBIC)
        n = dw * dh
        out[:n] = sy.reshape(-1); out[n:n + n // 4] = su.reshape(-1); out[n + n // 4:] = sv.reshape(-1)
        prev_y = y
    return b - a, acc


class CpuArm:
    """All host cores over pictures of the clip, split into small contiguous ranges per worker task.  The worker pool
    lives across calls (start-up and library warm-up are outside every timed region)."""

    def __init__(self, path, payload, keyframe, w, h, dw, dh, n_clip, cores):
        import multiprocessing as mp
        self.cores, self.n_clip = cores, n_clip
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init,
                                                initargs=(path, payload, keyframe, w, h, dw, dh))
        self.pool.map(_cpu_worker, [(a, min(a + 4, n_clip)) for a in range(0, min(n_clip, 4 * cores), 4)])   # warm

    def run(self, count):
        """Process `count` pictures (wrapping around the clip).  Returns (pictures/s, seconds, pictures)."""
        piece = 8
        tasks = []
        done = 0
        while done < count:
            a = done % self.n_clip
            b = min(a + piece, self.n_clip, a + (count - done))
            tasks.append((a, b))
            done += b - a
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, tasks, chunksize=4)
        dt = time.perf_counter() - t0
        frames = sum(r[0] for r in res)
        return frames / dt, dt, frames

    def close(self):
        self.pool.close()
        self.pool.join()


# ---- workload ---------------------------------------------------------------------------------------------------
def make_clip(n_frames: int, tmpdir: str):
    from video_transformer_b200 import container, synth
    wr = synth.H264PcmWriter(SRC_W, SRC_H, FPS, 1)
    cuts = set(synth.scene_cut_frames(n_frames, FPS, seed=42))
    path = os.path.join(tmpdir, "bench_1080p.h264")
    scene = 0
    with open(path, "wb") as f:
        for k in range(n_frames):
            if k in cuts:
                scene += 1
            if k % GOP == 0 or k in cuts:
                f.write(wr.idr(*synth.testsrc_frame(SRC_W, SRC_H, k, scene)))
            else:
                f.write(wr.skip())
    idx = container.probe(path)
    return path, idx, sorted(cuts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-frames", type=int, default=32768,
                    help="pictures the host-core baseline processes at N=1 (about 10 s of CPU work)")
    ap.add_argument("--batch-frames", type=int, default=int(os.environ.get("VT_BENCH_BATCH", "32")),
                    help="pictures per pipeline batch of the end-to-end arm")
    ap.add_argument("--cpu-step-frames", type=int, default=2048, help="pictures per step of --impl reference")
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0))
    tmpdir = tempfile.mkdtemp(prefix="vtbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

    if args.impl == "reference":
        if rank != 0:
            return 0
        from video_transformer_b200 import ops
        from video_transformer_b200.ingest import SegmentIngestor  # noqa: F401  (index/payload helper only)
        n_clip = 4 * STEP_FRAMES
        path, idx, _ = make_clip(n_clip, tmpdir)
        payload = _payload_for(idx, path)
        dw = ops.scale_width_for_height(SRC_W, SRC_H, OUT_H)
        arm = CpuArm(path, payload, idx.keyframe, SRC_W, SRC_H, dw, OUT_H, n_clip, cores)
        ref_step = args.cpu_step_frames               # pictures per reference step (a bounded sample of the workload)
        times = []
        for s in range(Wm + K):
            fps, dt, frames = arm.run(ref_step)
            if s >= Wm:
                times.append(dt)
        arm.close()
        total = sum(times)
        value = ref_step * K / total
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                "warmup": Wm, "ms_per_step": 1000 * total / K, "step_pictures": ref_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": _config(dw, world),
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%d pictures per step, %d steps; libswscale bicubic + OpenCV SAD/hist + "
                                           "C PCM re-layout, one process per core (ffmpeg binary absent)"
                                           % (ref_step, K)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from ctypes import c_void_p

    from video_transformer_b200 import _lib, ingest, ops
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    n_clip = (Wm + K) * STEP_FRAMES
    n_clip = min(n_clip, 16 * STEP_FRAMES)              # bounded clip; steps cycle through its chunks
    path, idx, cuts = make_clip(n_clip, tmpdir)
    n_chunks = n_clip // STEP_FRAMES
    opts = ingest.IngestOptions(target_height=OUT_H, batch_frames=args.batch_frames, device=str(dev))
    eng = ingest.SegmentIngestor(idx, opts)
    dw, dh, fb = eng.out_w, eng.out_h, eng.frame_bytes

    # ---- device-resident arm: whole bitstream in HBM, one step = decode + score + scale of 256 pictures ----------
    host = np.memmap(path, dtype=np.uint8, mode="r")
    bs_dev = torch.zeros(host.size + 64, dtype=torch.uint8, device=dev)
    bs_dev[:host.size].copy_(torch.from_numpy(np.array(host)))
    F = STEP_FRAMES
    surf = torch.empty((F, eng.rows, eng.pitch), dtype=torch.uint8, device=dev)
    out = torch.empty((F, fb), dtype=torch.uint8, device=dev)
    sad = torch.empty(F, dtype=torch.int64, device=dev)
    hist = torch.empty((F, 256), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    sp = c_void_p(st.cuda_stream)
    names = ("decode", "score", "scale")
    ev = {n: [] for n in names}

    def step(chunk: int, timed: bool):
        b0 = chunk * F
        pay = eng.payload[b0:b0 + F].copy()              # absolute offsets: the whole stream is resident
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
        if timed:
            marks[0].record(st)
        _lib.check(L.vt_h264_pcm_decode(c_void_p(bs_dev.data_ptr()), pay.ctypes.data, F, eng.w, eng.h, None,
                                        c_void_p(surf.data_ptr()), eng.pitch, eng.surface_bytes, sp))
        if timed:
            marks[1].record(st)
        _lib.check(L.vt_sad_hist_u8(c_void_p(surf.data_ptr()), eng.pitch, eng.surface_bytes, eng.w, eng.h, None, F,
                                    c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), sp))
        if timed:
            marks[2].record(st)
        _lib.check(L.vt_scale_nv12_to_yuv420p(eng.plan._h, c_void_p(surf.data_ptr()), eng.pitch, eng.surface_bytes,
                                              c_void_p(out.data_ptr()), fb, F, sp))
        if timed:
            marks[3].record(st)
            for i, n in enumerate(names):
                ev[n].append((marks[i], marks[i + 1]))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for s in range(max(Wm, 3)):
        step(s % n_chunks, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.vt_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(st)
    for s in range(K):
        step((Wm + s) % n_chunks, True)
    t_end.record(st)
    barrier()
    launches = L.vt_launch_count() - launches0
    dev_ms = t_start.elapsed_time(t_end)
    kern_ms = {n: sum(a.elapsed_time(b) for a, b in ev[n]) / K for n in names}

    # ---- end-to-end arm: host bytes -> pinned host frame buffers through SegmentIngestor.run ----------------------
    sink = ingest.PinnedRing()
    n_e2e = min(K, n_chunks) * F
    reps = (K * F + n_e2e - 1) // n_e2e
    eng.run(0, min(max(Wm, 1) * F, n_clip), sink)       # warm-up pass
    barrier()
    eng.h2d_bytes = eng.d2h_bytes = 0
    t0 = time.perf_counter()
    done = 0
    for r in range(reps):
        cnt = min(n_e2e, K * F - done)
        eng.run(0, cnt, sink)
        done += cnt
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_h2d, e2e_d2h = eng.h2d_bytes, eng.d2h_bytes
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    # the same pass with the frames handed over on the device (no D2H of frames): what a GPU-resident consumer sees
    t0 = time.perf_counter()
    done = 0
    for r in range(reps):
        cnt = min(n_e2e, K * F - done)
        eng.run(0, cnt, None, device_sink=lambda chunk, k0: None)
        done += cnt
    torch.cuda.synchronize(dev)
    dsink_s = time.perf_counter() - t0
    barrier()

    t = torch.tensor([dev_ms, e2e_s * 1000.0, dsink_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, dsink_ms_max = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_kind = load_peaks()
    dominant = max(names, key=lambda n: kern_ms[n])
    alg = {"decode": BYTES_DECODE, "score": BYTES_SCORE, "scale": BYTES_SCALE}
    ach = {n: alg[n] * F / (kern_ms[n] * 1e-3) / 1e9 for n in names}
    traffic = load_traffic()
    value = world * K * F / (dev_ms_max * 1e-3)
    e2e_value = world * K * F / (e2e_ms_max * 1e-3)
    segs_per_s = value / (720.0 * FPS)                   # shipped plan for 7200 s: 10 segments of 720 s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": _config(dw, world),
        "segments_per_sec": segs_per_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d // max(K, 1),
                "d2h_bytes_per_step": e2e_d2h // max(K, 1)},
        "e2e_device_sink": {"value": world * K * F / (dsink_ms_max * 1e-3), "unit": UNIT,
                            "note": "host bitstream in, frames consumed on the device (SegmentIngestor.run(device_sink=...)), "
                                    "scores to host; not the contract's e2e"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": ach[dominant], "peak": peak, "unit": "GB/s",
                     "frac": ach[dominant] / peak,
                     "traffic": (traffic[dominant] * F if traffic and dominant in traffic else None),
                     "traffic_source": "profiles/r01_traffic.json (ncu dram__bytes_read+write per launch of %d pictures)" % F,
                     "algorithmic_bytes": alg[dominant] * F, "peak_source": peak_kind,
                     "frac_of_nominal_8000": ach[dominant] / 8000.0},
        "kernels": {n: {"ms_per_step": kern_ms[n], "achieved_gbs": ach[n], "frac": ach[n] / peak,
                        "alg_bytes_per_picture": alg[n]} for n in names},
        "clocks": clocks,
    }
    if world == 1 and args.cpu_sample_frames > 0:
        payload = eng.payload
        arm = CpuArm(path, payload, idx.keyframe, SRC_W, SRC_H, dw, dh, n_clip, cores)
        cfps, cdt, cframes = arm.run(args.cpu_sample_frames)
        arm.close()
        line["cpu_baseline"] = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d pictures (the bench clip, wrapped) in %.1f s; libswscale bicubic + "
                                          "OpenCV SAD/hist + C PCM re-layout, one process per core" % (cframes, cdt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _payload_for(idx, path):
    """Payload offsets of every picture (host-side index; no GPU needed)."""
    import ctypes as C
    from video_transformer_b200 import _lib
    L = _lib.lib()
    host = np.memmap(path, dtype=np.uint8, mode="r")
    n = idx.n_frames
    pay = np.zeros(n, np.uint64)
    offs = np.ascontiguousarray(idx.nal_offsets, dtype=np.uint64)
    sizes = np.ascontiguousarray(idx.nal_sizes, dtype=np.uint32)
    _lib.check(L.vt_h264_pcm_layout(host.ctypes.data, host.size, offs.ctypes.data, sizes.ctypes.data, n,
                                    pay.ctypes.data))
    return pay


def _config(dw, world):
    return {"workload": "configs[1]: synthetic 1920x1080@30 H.264 (I_PCM IDR / GOP 30 + P_Skip) -> %dx%d yuv420p "
                        "bicubic + SAD/hist + segment plan" % (dw, OUT_H),
            "step_pictures": STEP_FRAMES, "l2": "inputs larger than L2 (step surfaces = 852 MB)",
            "sharding": "1 process/GPU, GOP-aligned shards, no collective" if world > 1 else "single GPU",
            "bitstream": "I_PCM/P_Skip synthetic; entropy decode cost not representative of CABAC content",
            "nvdec": "unavailable on this pool (driver refuses video decode); decode = CUDA PCM-intra kernel"}


if __name__ == "__main__":
    sys.exit(main())

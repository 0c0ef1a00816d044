#!/usr/bin/env python
"""bench.py -- 720p output frames/s of the long-video ingest path (BASELINE.json metric) on N B200s.

Workload: BASELINE.json configs[1], "2-hour synthetic 1080p30 H.264 -> 720p downscale + segment", run as a bounded
stream of GOP-aligned pieces of ONE synthetic 1920x1080@30 testsrc clip (I_PCM IDR every 30 pictures + all-skip P
pictures, video_transformer_b200/synth.py; entropy decoding cost is NOT representative of real CABAC content) through
decode -> SAD/histogram -> swscale-exact bicubic 1280x720 YUV420P.

`value`      kernels only, bitstream already resident in HBM: one step = R passes of 256 pictures (decode + score +
             scale), R chosen once so that the K timed steps last >= 1 s; CUDA events on the launching stream.
`e2e`        the reference-facing plugin call: video_segmenter.extract_segment(clip.mp4, start, end, segment.mp4) on
             /dev/shm, wall clock.  Every call reads host bytes, stream-copies its samples into a faststart MP4, runs
             the GPU pass and lands the frames in `<segment>.frames` (D2H straight into the registered mapping of the
             file) plus the JSON sidecar.  One step = UNITS_PER_STEP such calls per GPU, each UNIT_PICTURES long.
N > 1        ONE clip, sharded dynamically: ranks pull unit indices from a shared counter (torch's TCPStore -- control
             plane only; no collective touches pixels), so a rank whose D2H path is slower takes fewer units.  The
             record carries per-rank units / D2H GB/s, the concurrent D2H ceiling measured in the same run, and
             `cuts_equal_single_gpu` (cuts merged from all ranks' shards == a single-GPU pass over the whole clip).
`--impl reference`  the same work on the host cores (libswscale bicubic, the faster of OpenCV / tuned C for SAD+hist,
             C PCM re-layout; one process per core), because the literal reference path (ffmpeg child processes)
             cannot run here: the image has no ffmpeg binary.  It never loads libvtseg.so.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path
import os
import shutil
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
PASS_FRAMES = 256                      # pictures per kernel launch of the device-resident arm
UNIT_PICTURES = 1920                   # pictures per extract_segment call of the e2e arm (64 GOPs = 64 s of video)
UNITS_PER_STEP = 2                     # e2e units per GPU and step
CLIP_FRAMES = 3840                     # the bench clip: 128 s
METRIC = "720p_output_frames_per_sec"
UNIT = "frames/s"
# algorithmic bytes per picture (SURVEY.md section 8d / BASELINE.md section 3)
BYTES_SCALE = 3110400 + 1382400          # 1080p NV12 in, 720p YUV420P out
BYTES_SCORE = 2 * SRC_W * SRC_H          # cur + prev luma
BYTES_DECODE = 2 * 3110400               # samples in, NV12 surface out
BYTES_FUSED_C2 = 3110400 + 2073600 + 1382400   # scale + score on the source luma in one pass (SURVEY.md 8d "Fused C2")


def load_traffic():
    """DRAM bytes per picture per kernel from the newest committed ncu capture (profiles/r*_traffic.json), or None."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        names = sorted(n for n in os.listdir(pdir) if n.endswith("_traffic.json"))
        d = json.load(open(os.path.join(pdir, names[-1])))
        return d["per_picture_bytes"], "profiles/" + names[-1]
    except Exception:  # noqa: BLE001
        return None, None


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
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---- the clip -------------------------------------------------------------------------------------------------
def make_clip(n_frames: int, path: str):
    """Write the synthetic Annex-B clip.  Returns (payload offsets u64[n], keyframe flags, cut list): the generator knows
    where every picture's samples are, so the host-core arm needs no bitstream parser (and no libvtseg.so)."""
    from video_transformer_b200 import synth
    wr = synth.H264PcmWriter(SRC_W, SRC_H, FPS, 1)
    cuts = set(synth.scene_cut_frames(n_frames, FPS, seed=42))
    payload = np.zeros(n_frames, np.uint64)
    key = np.zeros(n_frames, bool)
    scene = 0
    pos = 0
    last = 0
    with open(path, "wb") as f:
        for k in range(n_frames):
            if k in cuts:
                scene += 1
            if k % GOP == 0 or k in cuts:
                b = wr.idr(*synth.testsrc_frame(SRC_W, SRC_H, k, scene))
                last = pos + wr.last_payload_offset
                key[k] = True
            else:
                b = wr.skip()
            payload[k] = last
            f.write(b)
            pos += len(b)
    return payload, key, sorted(cuts)


# ---- host-core arm -----------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(path, payload, w, h, dw, dh):
    import cv2
    from oracle import coracle, ffsws
    try:
        cv2.setNumThreads(1)
    except Exception:  # noqa: BLE001
        pass
    _W.update(buf=np.memmap(path, dtype=np.uint8, mode="r"), payload=payload, w=w, h=h, dw=dw, dh=dh,
              coracle=coracle, ffsws=ffsws, sws=ffsws.available(), cv2=cv2)
    # SAD + histogram: take the faster of OpenCV (norm L1 + calcHist) and the tuned C loop -- the stronger baseline
    y = np.frombuffer(np.random.default_rng(0).bytes(w * h), np.uint8).reshape(h, w)
    p = np.roll(y, 1, 1).copy()

    def t(fn):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        return (time.perf_counter() - t0) / 3

    t_cv = t(lambda: (cv2.norm(y, p, cv2.NORM_L1), cv2.calcHist([y], [0], None, [256], [0, 256])))
    t_c = t(lambda: coracle.sad_hist(y, p, True))
    _W["score_c"] = t_c < t_cv


def _cpu_worker(rng):
    """Decode + score + downscale pictures [a,b) on one host core.  Returns (frames, checksum, seconds per stage)."""
    a, b = rng
    W = _W
    co, ff, cv2 = W["coracle"], W["ffsws"], W["cv2"]
    w, h, dw, dh = W["w"], W["h"], W["dw"], W["dh"]
    out = np.empty((dw * dh * 3 // 2,), np.uint8)     # the segment frame buffer slot being filled
    prev_y = None
    cur = None
    last_payload = None
    acc = 0
    t_dec = t_score = t_scale = 0.0
    if a > 0:                                          # predecessor of the task's first picture, for its SAD
        last_payload = int(W["payload"][a - 1])
        cur = co.pcm_picture_to_yuv420p(W["buf"], last_payload, w, h)
        prev_y = cur[0]
    for k in range(a, b):
        t0 = time.perf_counter()
        p = int(W["payload"][k])
        if p != last_payload:                          # IDR: re-layout the PCM samples; skip pictures repeat
            cur = co.pcm_picture_to_yuv420p(W["buf"], p, w, h)
            last_payload = p
        y, u, v = cur
        t1 = time.perf_counter()
        if W["score_c"]:
            s, hist = co.sad_hist(y, prev_y, True)
            acc += s + int(hist[16])
        else:
            if prev_y is not None:
                acc += int(cv2.norm(y, prev_y, cv2.NORM_L1))
            hist = cv2.calcHist([y], [0], None, [256], [0, 256])
            acc += int(hist[16, 0])
        t2 = time.perf_counter()
        if W["sws"]:
            sy, su, sv = ff.scale_yuv420p(y, u, v, dw, dh, ff.SWS_BICUBIC)
        else:
            sy, su, sv = co.scale_yuv420p(y, u, v, dw, dh, co.BICUBIC)
        n = dw * dh
        out[:n] = sy.reshape(-1); out[n:n + n // 4] = su.reshape(-1); out[n + n // 4:] = sv.reshape(-1)
        t3 = time.perf_counter()
        t_dec += t1 - t0; t_score += t2 - t1; t_scale += t3 - t2
        prev_y = y
    return b - a, acc, (t_dec, t_score, t_scale), (W["sws"], W["score_c"])


class CpuArm:
    """All host cores over pictures of the clip, split into small contiguous ranges per worker task.  The worker pool
    lives across calls (start-up and library warm-up are outside every timed region)."""

    def __init__(self, path, payload, w, h, dw, dh, n_clip, cores):
        import multiprocessing as mp
        self.cores, self.n_clip = cores, n_clip
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init,
                                                initargs=(path, payload, w, h, dw, dh))
        self.pool.map(_cpu_worker, [(a, min(a + 4, n_clip)) for a in range(0, min(n_clip, 4 * cores), 4)])   # warm
        self.stage_s = [0.0, 0.0, 0.0]
        self.stage_frames = 0
        self.kind = (None, None)

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
        for r in res:
            for i in range(3):
                self.stage_s[i] += r[2][i]
        self.stage_frames += frames
        self.kind = res[0][3]
        return frames / dt, dt, frames

    def describe(self):
        n = max(self.stage_frames, 1)
        sws, score_c = self.kind
        return {"decode_ms_per_picture": 1e3 * self.stage_s[0] / n, "score_ms_per_picture": 1e3 * self.stage_s[1] / n,
                "scale_ms_per_picture": 1e3 * self.stage_s[2] / n,
                "sws": "live libswscale 9.1 (SIMD, bicubic)" if sws else "c-oracle (scalar restatement; libswscale absent)",
                "score": "tuned C (psadbw SAD + 4-way histogram)" if score_c else "OpenCV norm(L1) + calcHist"}

    def close(self):
        self.pool.close()
        self.pool.join()


def one_core_rate(path, payload, dw, dh, n_clip, frames=96):
    """The same worker on ONE core, in this process (for the per-core figure)."""
    _cpu_worker_init(path, payload, SRC_W, SRC_H, dw, dh)
    _cpu_worker((0, 8))
    t0 = time.perf_counter()
    n, _, _, _ = _cpu_worker((8, 8 + min(frames, n_clip - 8)))
    return n / (time.perf_counter() - t0)


def _out_width():
    # `-2:720` semantics of ffmpeg's scale filter: width from the aspect ratio, rounded to an even number
    return int(round(SRC_W * OUT_H / SRC_H / 2.0)) * 2


def run_reference(args, K, Wm, cores):
    tmpdir = tempfile.mkdtemp(prefix="vtbench_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        n_clip = 4 * PASS_FRAMES
        path = os.path.join(tmpdir, "bench_1080p.h264")
        payload, _key, _ = make_clip(n_clip, path)
        dw = _out_width()
        arm = CpuArm(path, payload, SRC_W, SRC_H, dw, OUT_H, n_clip, cores)
        ref_step = args.cpu_step_frames               # pictures per reference step (a bounded sample of the workload)
        times = []
        for s in range(Wm + K):
            fps, dt, frames = arm.run(ref_step)
            if s >= Wm:
                times.append(dt)
        desc = arm.describe()
        arm.close()
        one = one_core_rate(path, payload, dw, OUT_H, n_clip)
        total = sum(times)
        value = ref_step * K / total
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                "warmup": Wm, "ms_per_step": 1000 * total / K, "step_pictures": ref_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": _config(dw, 1),
                "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                      "one_core_value": one,
                                      "sample": "%d pictures per step, %d steps; one process per core; frames written "
                                                "into an in-memory segment buffer (no file); ffmpeg binary absent, so "
                                                "this is the restated reference work, not the ffmpeg child process"
                                                % (ref_step, K)}, **desc),
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "native_so": "oracle/libvtoracle.so + libswscale (ctypes) + OpenCV; libvtseg.so is not loaded"}
        print(json.dumps(line))
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    return 0


# ---- shared unit counter (control plane of the N-GPU arms) --------------------------------------------------------
class UnitCounter:
    """next() hands out 0, 1, 2, ... across all ranks: TCPStore.add on the process group's store (one small TCP round
    trip to rank 0 per unit; no tensor, no collective)."""

    def __init__(self, store, key: str):
        self.store, self.key, self.local = store, key, 0

    def next(self) -> int:
        if self.store is None:
            self.local += 1
            return self.local - 1
        return int(self.store.add(self.key, 1)) - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-frames", type=int, default=32768,
                    help="pictures the host-core baseline processes at N=1 (about 10 s of CPU work)")
    ap.add_argument("--batch-frames", type=int, default=int(os.environ.get("VT_BENCH_BATCH", "64")),
                    help="pictures per pipeline batch of the end-to-end arms")
    ap.add_argument("--cpu-step-frames", type=int, default=2048, help="pictures per step of --impl reference")
    ap.add_argument("--min-timed-s", type=float, default=1.0, help="the device-resident timed region lasts at least this long")
    ap.add_argument("--workload", default="config1", choices=["config1", "config3"],
                    help="config1 = BASELINE.json configs[1] (the contract's line); config3 = configs[3], a URL.txt-style batch "
                         "of 64 synthetic 720p clips through batch.ingest_batch_dynamic (its own JSON line)")
    ap.add_argument("--consume", action="store_true",
                    help="config3: a consumer deletes every clip's .frames right after its ingest (the landing files are "
                         "then recycled instead of being allocated and page-locked anew for every clip)")
    ap.add_argument("--diag", action="store_true", help="also time the engine API with a pinned ring / a direct landing on "
                    "every rank at once (which part of the plugin call limits multi-GPU scaling)")
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, K, Wm, cores)

    import torch
    import torch.distributed as dist
    from ctypes import c_void_p

    from video_transformer_b200 import _lib, container, ingest, landing, shard, video_segmenter
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    store = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        store = dist.distributed_c10d._get_default_store()
    L = _lib.lib()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- the clip: written once by rank 0 into /dev/shm, opened by every rank ---------------------------------------
    box = [None]
    if rank == 0:
        box[0] = tempfile.mkdtemp(prefix="vtbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    workdir = box[0]
    raw_path = os.path.join(workdir, "bench_1080p.h264")
    mp4_path = os.path.join(workdir, "bench_1080p.mp4")
    n_clip = CLIP_FRAMES
    if rank == 0:
        payload_gen, key_gen, cuts_truth = make_clip(n_clip, raw_path)
        container.annexb_to_mp4(raw_path, mp4_path)
        np.save(os.path.join(workdir, "cuts.npy"), np.asarray(cuts_truth, np.int64))
    barrier()
    cuts_truth = np.load(os.path.join(workdir, "cuts.npy")).tolist()
    try:
        if args.workload == "config3":
            return _run_config3(args, rank, world, local, dev, barrier, workdir, torch, dist, video_segmenter)
        return _run_ours(args, K, Wm, rank, world, local, cores, dev, store, L, barrier, workdir, raw_path, mp4_path,
                         n_clip, cuts_truth, torch, dist, c_void_p, _lib, container, ingest, landing, shard,
                         video_segmenter)
    finally:
        landing.release_all()
        if world > 1:
            try:
                dist.barrier()
            except Exception:  # noqa: BLE001
                pass
        if rank == 0:
            shutil.rmtree(workdir, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def _run_ours(args, K, Wm, rank, world, local, cores, dev, store, L, barrier, workdir, raw_path, mp4_path, n_clip,
              cuts_truth, torch, dist, c_void_p, _lib, container, ingest, landing, shard, video_segmenter):
    idx = container.probe(raw_path)
    opts = ingest.IngestOptions(target_height=OUT_H, batch_frames=args.batch_frames, device=str(dev))
    eng = ingest.SegmentIngestor(idx, opts)
    dw, dh, fb = eng.out_w, eng.out_h, eng.frame_bytes
    n_chunks = n_clip // PASS_FRAMES

    # ---- device-resident arm: whole bitstream in HBM, one pass = decode + score + scale of 256 pictures -------------
    host = np.memmap(raw_path, dtype=np.uint8, mode="r")
    bs_dev = torch.zeros(host.size + 64, dtype=torch.uint8, device=dev)
    bs_dev[:host.size].copy_(torch.from_numpy(np.array(host)))
    F = PASS_FRAMES
    surf = torch.empty((F, eng.rows, eng.pitch), dtype=torch.uint8, device=dev)
    out = torch.empty((F, fb), dtype=torch.uint8, device=dev)
    sad = torch.empty(F, dtype=torch.int64, device=dev)
    hist = torch.empty((F, 256), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    sp = c_void_p(st.cuda_stream)
    names = ("decode", "score", "scale")
    ev = {n: [] for n in names}

    def one_pass(chunk: int, timed: bool):
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

    for s in range(max(Wm, 3)):
        one_pass(s % n_chunks, False)
    torch.cuda.synchronize(dev)
    # calibrate R (passes per step) so that K steps last >= min_timed_s; identical on every rank (max over ranks)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(st)
    for s in range(4):
        one_pass(s % n_chunks, False)
    c1.record(st)
    torch.cuda.synchronize(dev)
    pass_ms = torch.tensor([c0.elapsed_time(c1) / 4.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pass_ms, op=dist.ReduceOp.MIN)
    R = int(max(1, min(4096, np.ceil(args.min_timed_s * 1e3 / max(K, 1) / float(pass_ms[0])))))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.vt_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(st)
    q = 0
    for s in range(K):
        for r in range(R):
            one_pass(q % n_chunks, r == 0 or R <= 8)     # per-kernel events on the first pass of each step
            q += 1
    t_end.record(st)
    barrier()
    launches = L.vt_launch_count() - launches0
    dev_ms = t_start.elapsed_time(t_end)
    kern_ms = {n: sum(a.elapsed_time(b) for a, b in ev[n]) / len(ev[n]) for n in names}   # per 256-picture launch
    clocks_dev = sampler.stop() if rank == 0 else None
    del bs_dev, surf, out

    # ---- end-to-end arm: the plugin call, units pulled from a shared counter ---------------------------------------
    video_segmenter.configure(target_height=OUT_H, batch_frames=args.batch_frames, device=str(dev), frame_buffers=True)
    unit_s = UNIT_PICTURES / FPS
    n_windows = n_clip // UNIT_PICTURES
    my_dir = os.path.join(workdir, "segments", "rank%d" % rank)
    os.makedirs(my_dir, exist_ok=True)

    def run_unit(u: int, slot: int):
        j = u % n_windows
        out_mp4 = os.path.join(my_dir, "segment_%04d.mp4" % slot)
        ok = video_segmenter.extract_segment(input_path=mp4_path, start=j * unit_s, end=(j + 1) * unit_s,
                                             output_path=out_mp4, stream_copy=True)
        if not ok:
            raise RuntimeError("extract_segment failed on unit %d" % u)
        return out_mp4

    t0 = time.perf_counter()
    first_out = run_unit(rank, 0)
    cold_s = time.perf_counter() - t0
    side = json.loads(open(first_out[:-4] + ".json").read())
    assert side["frames"] == UNIT_PICTURES and side["frame_size"] == [dw, dh], side
    landing_mode = side["landing"]
    for wu in range(1, max(2, min(Wm, 4))):
        run_unit(rank + wu, wu % 2)
    barrier()
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    counter = UnitCounter(store, "vtbench_units_timed")
    total_units = K * UNITS_PER_STEP * world
    eng_plugin = video_segmenter._ENGINE_CACHE["engine"][1]
    eng_plugin.h2d_bytes = eng_plugin.d2h_bytes = 0
    barrier()
    t0 = time.perf_counter()
    my_units = 0
    busy = 0.0
    while True:
        u = counter.next()
        if u >= total_units:
            break
        tu = time.perf_counter()
        run_unit(u, my_units % 2)
        busy += time.perf_counter() - tu
        my_units += 1
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    timings = dict(video_segmenter.LAST_TIMINGS)
    e2e_h2d, e2e_d2h = eng_plugin.h2d_bytes, eng_plugin.d2h_bytes
    clocks_e2e = sampler2.stop() if rank == 0 else None
    barrier()

    # ---- concurrent D2H ceiling, measured in the same run: every rank copies 44 MB chunks into pinned memory for the
    # same WINDOW of wall time (a fixed number of copies per rank would let the ranks on fast links finish first and
    # the slow ones then measure an emptier box: at N=8 that overstated the slow links by a third) ---------------------
    chunk = 44 << 20
    PROBE_S = 0.6
    d_src = torch.empty(chunk, dtype=torch.uint8, device=dev)
    h_dst = [torch.empty(chunk, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    s_copy = torch.cuda.Stream(dev)

    def d2h_window(seconds, each=None):
        """Copy chunks device -> pinned host, two in flight, until `seconds` have passed; `each(i)` runs per copy.
        Returns GB/s over the copies known complete at the last synchronisation inside the window."""
        from collections import deque
        with torch.cuda.stream(s_copy):
            for i in range(2):
                h_dst[i].copy_(d_src, non_blocking=True)
        s_copy.synchronize()
        barrier()
        pending = deque()
        n = done = 0
        t0 = t_done = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            with torch.cuda.stream(s_copy):
                h_dst[n % 2].copy_(d_src, non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_copy)
            if each is not None:
                each(n)
            pending.append(e)
            n += 1
            if len(pending) >= 2:
                pending.popleft().synchronize()
                done += 1
                t_done = time.perf_counter()
        s_copy.synchronize()
        return done * chunk / max(t_done - t0, 1e-9) / 1e9

    ceil_gbs = d2h_window(PROBE_S)
    barrier()
    # ---- concurrent file-write ceiling of the box: every rank copies UNIT-sized byte ranges of the clip at once, the
    # way the product does (the MP4 half of extract_segment is a file -> file copy of the unit's samples: plain stores
    # into a recycled mapping of the output, landing.acquire_mapped), and with copy_file_range for comparison.  The
    # synthetic bitstream is uncompressed, ~115 KB per picture, so this path carries far more bytes than a real H.264
    # stream would ---------
    import mmap as _mmap
    unit_bytes = int(idx.nal_offsets[min(UNIT_PICTURES, n_clip - 1)] - idx.nal_offsets[0])
    probe_out = os.path.join(my_dir, "write_probe.bin")
    src_map = np.memmap(raw_path, dtype=np.uint8, mode="r")
    fi = os.open(raw_path, os.O_RDONLY)
    fo = os.open(probe_out, os.O_RDWR | os.O_CREAT, 0o644)
    os.ftruncate(fo, unit_bytes)
    probe_mm = _mmap.mmap(fo, unit_bytes)
    dst_map = np.frombuffer(probe_mm, dtype=np.uint8)

    def copy_once():
        dst_map[:] = src_map[:unit_bytes]

    def copy_in_kernel():
        done = 0
        while done < unit_bytes:
            done += os.copy_file_range(fi, fo, unit_bytes - done, done, done)

    def copy_window(fn, seconds):
        fn()
        barrier()
        tw = time.perf_counter()
        k = 0
        while time.perf_counter() - tw < seconds:
            fn()
            k += 1
        return k * unit_bytes / (time.perf_counter() - tw) / 1e9

    write_gbs = copy_window(copy_once, 0.4)
    try:
        kernel_copy_gbs = copy_window(copy_in_kernel, 0.4)
    except (OSError, AttributeError):
        kernel_copy_gbs = 0.0
    os.close(fi)
    bytes_per_picture_bs = unit_bytes / UNIT_PICTURES
    barrier()
    # ---- the same three transfers AT ONCE, in the workload's byte mix: per picture the path moves frame_bytes of D2H,
    # the bitstream's bytes of H2D and a file copy of the bitstream (read + write) through the same host memory system.
    # This is the platform's ceiling for THIS byte mix, with no kernels and no Python in the way.
    mix_ratio = bytes_per_picture_bs / float(fb + 1032)
    h2d_n = max(1 << 16, int(chunk * mix_ratio) & ~4095)
    d_in = torch.empty(h2d_n, dtype=torch.uint8, device=dev)
    h_in = torch.empty(h2d_n, dtype=torch.uint8, pin_memory=True)
    s_b = torch.cuda.Stream(dev)
    stop = threading.Event()
    copied = [0]                                         # D2H chunks issued so far: the file copy keeps pace with them

    def copier():
        off = k = 0
        n_mix = min(h2d_n, unit_bytes)
        while not stop.is_set():
            if k >= copied[0]:
                time.sleep(0.0002)
                continue
            lo = off % max(1, unit_bytes - n_mix)
            dst_map[:n_mix] = src_map[lo:lo + n_mix]
            off += n_mix
            k += 1

    def each(i):
        with torch.cuda.stream(s_b):
            d_in.copy_(h_in, non_blocking=True)
        copied[0] = i + 1

    th = threading.Thread(target=copier)
    th.start()
    mixed_gbs = d2h_window(PROBE_S, each)
    s_b.synchronize()
    stop.set()
    th.join()
    del dst_map, src_map
    probe_mm.close()
    os.close(fo)
    os.unlink(probe_out)
    del d_src, h_dst, d_in, h_in, s_copy
    barrier()

    # ---- boundaries: shards pulled dynamically by all ranks, merged on rank 0, vs ONE single-GPU pass ---------------
    piece = 240                                          # 8 GOPs per shard: 16 shards over the clip
    pieces = [(a, min(a + piece, n_clip)) for a in range(0, n_clip, piece)]
    counter2 = UnitCounter(store, "vtbench_units_cuts")
    eng_sc = ingest.SegmentIngestor(idx, ingest.IngestOptions(target_height=OUT_H, batch_frames=args.batch_frames,
                                                              keep_frames=False, device=str(dev)))
    mine = []
    while True:
        u = counter2.next()
        if u >= len(pieces):
            break
        a, b = pieces[u]
        res = eng_sc.run(a, b, None)
        mine.append((a, res.sad))
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    cuts_equal = None
    n_cuts = None
    if rank == 0:
        parts = [p for g in gathered for p in g]
        _sad, _scores, cuts_n = shard.merge_and_score(parts, SRC_W, SRC_H, opts.scene_threshold)
        single = eng_sc.run(0, n_clip, None)
        cuts_equal = bool(np.array_equal(cuts_n, single.cuts) and set(cuts_truth) <= set(single.cuts.tolist()))
        n_cuts = int(len(single.cuts))

    diag = None
    if args.diag:
        diag = {}
        for name in ("engine_pinned_ring", "engine_direct_landing"):
            land = landing.acquire(os.path.join(my_dir, "diag.frames"), UNIT_PICTURES * fb) if "landing" in name else None
            eng.run(0, UNIT_PICTURES, ingest.PinnedRing() if land is None else None, landing=land)
            barrier()
            t0 = time.perf_counter()
            for r in range(6):
                a = (r % n_windows) * UNIT_PICTURES
                eng.run(a, a + UNIT_PICTURES, ingest.PinnedRing() if land is None else None, landing=land)
            torch.cuda.synchronize(dev)
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            diag[name] = world * 6 * UNIT_PICTURES / float(dt[0])
            barrier()

    # ---- engine API arm (N=1): SegmentIngestor.run into a pinned ring, no file ------------------------------------
    eng_ms = None
    if world == 1:
        sink = ingest.PinnedRing()
        eng.run(0, min(2 * PASS_FRAMES, n_clip), sink)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n_eng = 0
        for r in range(max(1, min(K, 4))):
            eng.run(0, n_clip, sink)
            n_eng += n_clip
        torch.cuda.synchronize(dev)
        eng_ms = (time.perf_counter() - t0) * 1e3

    # ---- the same pass with the frames handed over on the device (no D2H of frames): what a GPU-resident consumer sees
    dsink_fps = None
    if world == 1:
        eng.run(0, min(2 * PASS_FRAMES, n_clip), None, device_sink=lambda chunk_, k0: None)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n_ds = 0
        for r in range(max(2, min(K, 6))):
            eng.run(0, n_clip, None, device_sink=lambda chunk_, k0: None)
            n_ds += n_clip
        torch.cuda.synchronize(dev)
        dsink_fps = n_ds / (time.perf_counter() - t0)

    stats = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    per_rank = torch.tensor([float(my_units), busy, ceil_gbs, float(e2e_d2h), float(e2e_h2d), write_gbs, mixed_gbs, kernel_copy_gbs], dtype=torch.float64,
                            device=dev)
    all_rank = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_gather(all_rank, per_rank)
    else:
        all_rank = [per_rank]
    dev_ms_max, e2e_ms_max = float(stats[0]), float(stats[1])
    if rank != 0:
        return 0

    peak, peak_kind = load_peaks()
    dominant = max(names, key=lambda n: kern_ms[n])
    alg = {"decode": BYTES_DECODE, "score": BYTES_SCORE, "scale": BYTES_SCALE}
    ach = {n: alg[n] * F / (kern_ms[n] * 1e-3) / 1e9 for n in names}
    traffic, traffic_src = load_traffic()
    dram = {n: (traffic[n] * F / (kern_ms[n] * 1e-3) / 1e9 if traffic and n in traffic else None) for n in names}
    step_pictures = R * F
    value = world * K * step_pictures / (dev_ms_max * 1e-3)
    e2e_pictures = total_units * UNIT_PICTURES
    from video_transformer_b200 import landing as _landing
    landing_stats = _landing.stats()
    e2e_value = e2e_pictures / (e2e_ms_max * 1e-3)
    segs_per_s = value / (720.0 * FPS)                   # shipped plan for 7200 s: 10 segments of 720 s
    ranks = [{"rank": r, "units": int(t[0]), "busy_s": float(t[1]),
              "d2h_gbs": float(t[3]) / float(t[1]) / 1e9 if float(t[1]) > 0 else 0.0,
              "d2h_ceiling_gbs": float(t[2]), "file_copy_ceiling_gbs": float(t[5]),
              "d2h_in_mix_gbs": float(t[6]), "copy_file_range_gbs": float(t[7])} for r, t in enumerate(all_rank)]
    ceiling_fps = sum(r["d2h_ceiling_gbs"] for r in ranks) * 1e9 / (fb + 1032)
    write_fps = sum(r["file_copy_ceiling_gbs"] for r in ranks) * 1e9 / bytes_per_picture_bs
    mixed_fps = sum(r["d2h_in_mix_gbs"] for r in ranks) * 1e9 / (fb + 1032)
    binding = min(ceiling_fps, write_fps) if write_fps > 0 else ceiling_fps
    fused_ms = kern_ms["score"] + kern_ms["scale"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": _config(dw, world, step_pictures, R),
        "timed_region_s": dev_ms_max * 1e-3,
        "segments_per_sec": segs_per_s,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(float(t[4]) for t in all_rank) // max(K, 1)),
                "d2h_bytes_per_step": int(sum(float(t[3]) for t in all_rank) // max(K, 1)),
                "through": "video_segmenter.extract_segment (host MP4 in, faststart MP4 + .frames + .json on /dev/shm out)",
                "pictures": e2e_pictures, "seconds": e2e_ms_max * 1e-3, "unit_pictures": UNIT_PICTURES,
                "landing": landing_mode,
                "mp4_write": "recycled mapping" if landing_stats.get("mapped_files") else "copy_file_range",
                "last_call_ms": {k: round(v * 1e3, 2) for k, v in timings.items()}},
        "e2e_cold": {"value": UNIT_PICTURES / cold_s, "unit": UNIT,
                     "note": "first call of the process: index, plans, pinned staging, and a new landing file "
                             "(allocate + cudaHostRegister, ~4 GB/s on this box) -- later calls recycle it"},
        "e2e_ceiling": {"d2h_gbs_per_rank_concurrent": [r["d2h_ceiling_gbs"] for r in ranks],
                        "d2h_frames_per_s": ceiling_fps,
                        "file_copy_gbs_per_rank_concurrent": [r["file_copy_ceiling_gbs"] for r in ranks],
                        "file_copy_frames_per_s": write_fps, "bitstream_bytes_per_picture": bytes_per_picture_bs,
                        "frames_per_s": binding, "binding": "d2h" if binding == ceiling_fps else "file_copy",
                        "e2e_of_ceiling": e2e_value / binding if binding else None,
                        "mixed_d2h_gbs_per_rank": [r["d2h_in_mix_gbs"] for r in ranks],
                        "mixed_frames_per_s": mixed_fps, "e2e_of_mixed": e2e_value / mixed_fps if mixed_fps else None,
                        "note": "measured right after the timed arm, all ranks at once over the same 0.6 s window: (a) 44 MB chunks device -> pinned "
                                "host (a picture costs frame_bytes + 1032 B of D2H); (b) one unit's samples copied "
                                "into an existing mapping of a /dev/shm file (the stream-copy half of the call, "
                                "landing.acquire_mapped; per_rank.copy_file_range_gbs is the in-kernel copy beside it; "
                                "the synthetic bitstream is uncompressed PCM).  The lower of the two bounds the plugin call on this box.  (c) "
                                "`mixed`: D2H + H2D + file copy at once in the workload's byte mix (per picture: frame "
                                "bytes out, bitstream bytes in, bitstream bytes copied) -- the D2H rate that survives is "
                                "what this path could reach with no kernels and no host code in the way."},
        "per_rank": ranks,
        "cuts_equal_single_gpu": cuts_equal, "cuts": n_cuts,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": ach[dominant], "peak": peak, "unit": "GB/s",
                     "frac": ach[dominant] / peak,
                     "frac_dram": (dram[dominant] / peak if dram[dominant] else None),
                     "traffic": (traffic[dominant] * F if traffic and dominant in traffic else None),
                     "traffic_source": "%s (ncu dram__bytes_read+write per launch of %d pictures)" % (traffic_src, F),
                     "algorithmic_bytes": alg[dominant] * F, "peak_source": peak_kind,
                     "frac_of_nominal_8000": ach[dominant] / 8000.0},
        "kernels": {n: {"ms_per_launch": kern_ms[n], "achieved_gbs": ach[n], "frac": ach[n] / peak,
                        "frac_dram": (dram[n] / peak if dram[n] else None),
                        "alg_bytes_per_picture": alg[n],
                        "dram_bytes_per_picture": (traffic[n] if traffic and n in traffic else None)} for n in names},
        "fused_c2": {"ms_per_launch": fused_ms, "alg_bytes_per_picture": BYTES_FUSED_C2,
                     "frac": BYTES_FUSED_C2 * F / (fused_ms * 1e-3) / 1e9 / peak,
                     "note": "score + scale against the fused-pass roofline (source luma counted once)"},
        "clocks": clocks_dev, "clocks_e2e": clocks_e2e,
    }
    if diag is not None:
        line["diag_frames_per_s"] = diag
    if eng_ms is not None:
        line["e2e_engine"] = {"value": n_eng / (eng_ms * 1e-3), "unit": UNIT,
                              "note": "SegmentIngestor.run: host bitstream in, frames into a 3-slot pinned ring that is "
                                      "overwritten (round 1's e2e); no file, no MP4"}
    if dsink_fps is not None:
        line["e2e_device_sink"] = {"value": dsink_fps, "unit": UNIT,
                                   "note": "host bitstream in (H2D straight from the page-locked file mapping), frames consumed "
                                           "on the device (SegmentIngestor.run(device_sink=...)), scores to host; not the "
                                           "contract's e2e"}
    if world == 1 and args.cpu_sample_frames > 0:
        from oracle import coracle  # noqa: F401  (the CPU baseline leg: the one place bench.py may execute oracle/)
        payload = eng.payload
        arm = CpuArm(raw_path, payload, SRC_W, SRC_H, dw, dh, n_clip, cores)
        cfps, cdt, cframes = arm.run(args.cpu_sample_frames)
        desc = arm.describe()
        arm.close()
        line["cpu_baseline"] = dict({"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
                                     "sample": "%d pictures (the bench clip, wrapped) in %.1f s; one process per core; "
                                               "frames into an in-memory segment buffer (no file)" % (cframes, cdt)},
                                    **desc)
    print(json.dumps(line))
    return 0


def _run_config3(args, rank, world, local, dev, barrier, workdir, torch, dist, video_segmenter):
    """BASELINE.json configs[3]: 64 synthetic 720p clips, per-video dynamic queue over the ranks (atomic claims on the
    file system), every clip probed, planned, cut and ingested through the plugin's functions."""
    from video_transformer_b200 import batch, container, synth
    n_clips, n_pic, w, h = 64, 300, 1280, 720
    clips_dir = os.path.join(workdir, "clips")
    if rank == 0:
        os.makedirs(clips_dir, exist_ok=True)
        bs, _meta = synth.make_testsrc_h264(w, h, n_pic, fps=30, gop=30, cuts=[97, 211])
        raw = os.path.join(clips_dir, "clip.h264")
        open(raw, "wb").write(bs)
        first = os.path.join(clips_dir, "clip_00.mp4")
        container.annexb_to_mp4(raw, first)
        os.unlink(raw)
        for i in range(1, n_clips):
            shutil.copyfile(first, os.path.join(clips_dir, "clip_%02d.mp4" % i))
    barrier()
    vids = [os.path.join(clips_dir, "clip_%02d.mp4" % i) for i in range(n_clips)]
    video_segmenter.configure(target_height=OUT_H, batch_frames=args.batch_frames, device=str(dev), frame_buffers=True)
    # warm-up on a private copy (engine, plans, landing files of this size)
    warm = os.path.join(workdir, "warm_rank%d.mp4" % rank)
    shutil.copyfile(vids[0], warm)
    batch.ingest_video(warm, os.path.join(workdir, "warm_temp_%d" % rank))
    barrier()
    t0 = time.perf_counter()
    def consume(vid, ok):
        if args.consume and ok:
            for fp in video_segmenter.get_segment_dir(vid, os.path.join(workdir, "temp")).glob("*.frames"):
                fp.unlink()

    rep = batch.ingest_batch_dynamic(vids, os.path.join(workdir, "temp"), rank=rank, world=world, on_done=consume)
    torch.cuda.synchronize(dev)
    last_call = {k: round(v * 1e3, 2) for k, v in video_segmenter.LAST_TIMINGS.items()}
    eng_last = video_segmenter._ENGINE_CACHE.get("engine")
    if eng_last is not None:
        last_call["engine_setup"] = eng_last[1].setup_ms
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(rep.pictures), float(rep.segments_done), float(len(rep.processed)), float(len(rep.failed))],
                       dtype=torch.float64, device=dev)
    per = [torch.zeros_like(cnt) for _ in range(world)]
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_gather(per, cnt)
        torch.cuda.synchronize(dev)          # NCCL calls return once enqueued: wait until every rank has arrived, i.e.
                                             # has written its progress file, before rank 0 merges them
    else:
        per = [cnt]
    if rank == 0:
        merged = batch.merge_progress(os.path.join(workdir, "temp"), world)
        pictures = sum(float(t[0]) for t in per)
        segs = sum(float(t[1]) for t in per)
        print(json.dumps({
            "metric": METRIC, "workload": "configs[3]", "value": pictures / float(dt[0]), "unit": UNIT,
            "segments_per_sec": segs / float(dt[0]), "videos_per_sec": n_clips / float(dt[0]), "n_gpus": world,
            "seconds": float(dt[0]), "pictures": int(pictures), "segments": int(segs),
            "videos_per_rank": [int(t[2]) for t in per], "failed": int(sum(float(t[3]) for t in per)),
            "progress_json_processed": len(merged["processed"]),
            "progress_json_missing": sorted(set(Path(v).stem for v in vids) - set(merged["processed"])),
            "last_call_ms": last_call,
            "config": {"workload": "configs[3]: %d synthetic %dx%d@30 clips of %d pictures (I_PCM IDR / GOP 30 + P_Skip), "
                                   "same-height sources are converted NV12 -> YUV420P (not resized) + SAD/hist, one "
                                   "segment each; probe -> budget plan -> manifest -> extract_segment per clip"
                                   % (n_clips, w, h, n_pic),
                       "sharding": "per video, dynamic: ranks claim clips longest-first with O_EXCL files, no collective",
                       "e2e": "wall clock over batch.ingest_batch_dynamic, host MP4 in, MP4 + .frames + .json out",
                       "consumer": ("deletes each clip's .frames after its ingest: landing files are recycled"
                                    if args.consume else "keeps every .frames (26.5 GB): each clip allocates and "
                                    "page-locks a fresh landing file")},
            "data": "synthetic", "dtype": "u8"}))
    return 0


def _config(dw, world, step_pictures=None, passes=None):
    return {"workload": "configs[1]: synthetic 1920x1080@30 H.264 (I_PCM IDR / GOP 30 + P_Skip) -> %dx%d yuv420p "
                        "bicubic + SAD/hist + segment plan" % (dw, OUT_H),
            "step_pictures": step_pictures, "passes_per_step": passes, "pass_pictures": PASS_FRAMES,
            "l2": "inputs larger than L2 (surfaces of one pass = 852 MB; passes cycle through the clip)",
            "sharding": ("one clip; 1 process/GPU pulls GOP-aligned units from a shared counter, no collective on the "
                         "data path") if world > 1 else "single GPU",
            "bitstream": "I_PCM/P_Skip synthetic; entropy decode cost not representative of CABAC content",
            "nvdec": "unavailable on this pool (driver refuses video decode); decode = CUDA PCM-intra kernel"}


if __name__ == "__main__":
    sys.exit(main())

"""The fused ingest pass: bitstream (host) -> decode -> score -> convert/downscale -> segment frame buffers (host).

This is the GPU body behind extract_segment().  In the reference the same work happens inside two ffmpeg child
processes (/root/reference/src/utils/video_segmenter.py:118-154 cut/decode,
/root/reference/src/analyzer/content_analyzer.py:193-211 decode + `scale=-2:360`); here it is three kernels per
batch of pictures, driven from three CUDA streams so that the H2D copy of batch i+1, the kernels of batch i and
the D2H copy of batch i-1 overlap:

    copy-in stream : pinned bitstream bytes            -> HBM
    compute stream : vt_h264_pcm_decode (K0)  NV12 surfaces in HBM
                     vt_sad_hist_u8     (K3)  SAD + histogram on the decoded luma
                     vt_scale_nv12_to_yuv420p (K2) or vt_nv12_to_yuv420p (K1) -> output frames in HBM
    copy-out stream: output frames + per-picture SAD/hist -> pinned host memory -> sink

K4 (scores, cuts, boundaries) runs on the host in float64 from the integer SADs (scene.py).
PyTorch only owns memory, streams and events here.
"""
from __future__ import annotations

import os
from ctypes import c_void_p
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, ops, scene
from ._lib import check, lib
from .container import StreamIndex

_NO_PAYLOAD = np.uint64(0xFFFFFFFFFFFFFFFF)
# Pipeline buffers (device surfaces / outputs, pinned score buffers) of released engines, by geometry: the videos of a
# batch mostly share one geometry, and page-locked allocations are the slow part of building an engine.
_SLOT_POOL: dict = {}


@dataclass
class IngestOptions:
    target_height: int = 720           # downloader.max_resolution (config/config.yaml:75, README sample: 720)
    sws_flags: int = _lib.SWS_BICUBIC  # ffmpeg's scale filter default
    batch_frames: int = 64             # 32-picture launches run the scaler at about half its 256-picture rate
    scene_threshold: float = 0.10
    keep_frames: bool = True           # deliver output frames to the sink (False: scores only, config 3)
    never_upscale: bool = True         # sources at or below the target height are converted, not resized
    # BASELINE.json configs[4]: "1 fps frame sampling + 768x768 RGB resize for upload".  output = "rgb24" delivers
    # packed RGB at rgb_size (stretch, like `scale=W:H`); sample_every = N keeps pictures whose index is a multiple
    # of N (every picture is still decoded and scored, boundaries do not depend on the sampling).
    output: str = "yuv420p"
    rgb_size: tuple | None = None
    sample_every: int = 1
    device: str = "cuda"


@dataclass
class IngestResult:
    first: int
    last: int
    out_width: int
    out_height: int
    frame_bytes: int
    sad: np.ndarray                    # uint64 [n]
    hist: np.ndarray                   # uint32 [n, 256]
    scores: np.ndarray                 # float64 [n]; scores[0] belongs to picture `first`
    cuts: np.ndarray                   # absolute picture indices
    stats: dict = field(default_factory=dict)


class PinnedRing:
    """Sink that keeps the most recent chunks in pinned host memory and hands each to a callback."""

    def __init__(self, callback=None):
        self.callback = callback
        self.bytes = 0
        self.frames = 0

    def __call__(self, chunk: torch.Tensor, first_picture: int) -> None:
        self.bytes += chunk.numel()
        self.frames += chunk.shape[0]
        if self.callback is not None:
            self.callback(chunk, first_picture)


class SegmentIngestor:
    """Decode/score/scale engine for one indexed file on one GPU."""

    def __init__(self, index: StreamIndex, opts: IngestOptions | None = None, host_bytes: np.ndarray | None = None):
        import time as _time
        t_setup = [_time.perf_counter()]
        self.idx = index
        self.opts = opts or IngestOptions()
        self.dev = torch.device(self.opts.device)
        if self.dev.type == "cuda" and self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.host = host_bytes if host_bytes is not None else np.memmap(index.path, dtype=np.uint8, mode="r")
        self._src_map = None             # private mapping of the file, page-locked: H2D straight from the page cache
        n = index.n_frames
        L = lib()
        self.payload = np.zeros(n, np.uint64)
        offs = np.ascontiguousarray(index.nal_offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(index.nal_sizes, dtype=np.uint32)
        if index.kind == "mp4" and not index.extra.get("decodable", True):
            raise _lib.VtError(_lib.VT_ERR_UNSUPPORTED, "video track (%s) is not H.264 with one slice per sample; "
                               "the decode front end of this build handles PCM-intra H.264 only"
                               % index.extra.get("codec"))
        if index.extra.get("payload") is not None:       # container.classify_pcm already ran the slice classifier
            self.payload = index.extra["payload"]
        elif index.kind == "mp4":
            sps = np.frombuffer(index.sps, np.uint8)
            pps = np.frombuffer(index.pps, np.uint8)
            check(L.vt_h264_pcm_layout_ps(self.host.ctypes.data, self.host.size, sps.ctypes.data, sps.size,
                                          pps.ctypes.data, pps.size, offs.ctypes.data, sizes.ctypes.data, n,
                                          self.payload.ctypes.data))
        else:
            check(L.vt_h264_pcm_layout(self.host.ctypes.data, self.host.size, offs.ctypes.data, sizes.ctypes.data, n,
                                       self.payload.ctypes.data))
        t_setup.append(_time.perf_counter())
        self.offs, self.sizes = offs, sizes
        self.keyframes = np.nonzero(index.keyframe)[0].astype(np.int64)
        self.w, self.h = index.width, index.height
        self.pitch = (self.w + 255) // 256 * 256 if self.w > 256 else (self.w + 15) // 16 * 16
        self.rows = self.h + (self.h + 1) // 2
        self.surface_bytes = self.rows * self.pitch
        th = self.opts.target_height
        self.rgb_plan = None
        if self.dev.type == "cuda":
            torch.cuda.set_device(self.dev)               # plans, tables and per-device kernel attributes live on this device
        if self.opts.output not in ("yuv420p", "rgb24") or self.opts.sample_every < 1:
            raise ValueError("output must be 'yuv420p' or 'rgb24', sample_every >= 1")
        if self.opts.output == "rgb24":
            if self.opts.rgb_size:
                self.out_w, self.out_h = (int(v) for v in self.opts.rgb_size)
            elif th and self.h > th:
                self.out_w, self.out_h = ops.scale_width_for_height(self.w, self.h, th), th
            else:
                self.out_w, self.out_h = self.w, self.h
            self.plan = None
            self.rgb_plan = ops.RgbPlan(self.w, self.h, self.out_w, self.out_h, self.opts.sws_flags)
        elif th and (self.h > th or (not self.opts.never_upscale and self.h != th)):
            self.out_h = th
            self.out_w = ops.scale_width_for_height(self.w, self.h, th)
            self.plan = ops.ScalePlan(self.w, self.h, self.out_w, self.out_h, self.opts.sws_flags)
        else:
            self.out_w, self.out_h, self.plan = self.w, self.h, None
        self.frame_bytes = self.out_w * self.out_h + 2 * ((self.out_w + 1) // 2) * ((self.out_h + 1) // 2)
        if self.rgb_plan is not None:
            self.frame_bytes = self.out_w * self.out_h * 3
        t_setup.append(_time.perf_counter())
        B = self.opts.batch_frames
        self.B = B
        max_nal = int(sizes.max()) if n else 0
        # worst case bytes one batch needs on the device: B pictures + their container framing
        self.bs_cap = (max_nal + 64) * min(B, max(1, int(index.keyframe.sum()))) + 4096 * B + 4096
        self.bs_cap = (self.bs_cap + (1 << 20) - 1) >> 20 << 20          # whole MiB: engines of one geometry share buffers
        self.slots = []
        self._pool_key = (str(self.dev), B, self.rows, self.pitch, self.frame_bytes, self.bs_cap, self.opts.sample_every > 1)
        pooled = _SLOT_POOL.get(self._pool_key) or []
        self.n_slots = int(os.environ.get("VT_INGEST_SLOTS", "3"))   # batch i+1 is staged by a helper thread while i computes, i-1 copies out
        self.stage_thread = os.environ.get("VT_INGEST_THREAD", "1") == "1"
        for _ in range(self.n_slots):
            if pooled:
                sl = pooled.pop()
                sl.update(used=False, pending=None, kept=None, land=None, wfut=None, src_lo=0)
                self.slots.append(sl)
                continue
            self.slots.append({
                "bs_host": None, "src_lo": 0,
                "bs_dev": torch.empty(self.bs_cap + 64, dtype=torch.uint8, device=self.dev),
                "surf": torch.empty((B, self.rows, self.pitch), dtype=torch.uint8, device=self.dev),
                "out": torch.empty((B, self.frame_bytes), dtype=torch.uint8, device=self.dev),
                "out_host": None,        # pinned frames, allocated on first use: the direct landing never needs it
                "sel_surf": (torch.empty((B, self.rows, self.pitch), dtype=torch.uint8, device=self.dev)
                             if self.opts.sample_every > 1 else None),
                "sel_idx": torch.empty(B, dtype=torch.int32, device=self.dev),
                "sel_idx_host": torch.empty(B, dtype=torch.int32, pin_memory=True),
                "sad": torch.empty(B, dtype=torch.int64, device=self.dev),
                "hist": torch.empty((B, 256), dtype=torch.int32, device=self.dev),
                "sad_host": torch.empty(B, dtype=torch.int64, pin_memory=True),
                "hist_host": torch.empty((B, 256), dtype=torch.int32, pin_memory=True),
                "ev_in": torch.cuda.Event(), "ev_cmp": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                "used": False, "pending": None, "kept": None, "land": None, "wfut": None,
            })
        t_setup.append(_time.perf_counter())
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        t_setup.append(_time.perf_counter())
        self._register_source()
        t_setup.append(_time.perf_counter())
        self.setup_ms = dict(zip(("layout", "plans", "slots", "streams", "register_source"),
                                 (round((b - a) * 1e3, 3) for a, b in zip(t_setup, t_setup[1:]))))
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._pool = None
        # Staging copies (files that are not page-locked: small ones, and those over VT_INGEST_DIRECT_MAX_GB) are split
        # over a few threads: one thread moves 6-8 GB/s out of the page cache, which bounded every pass that is not
        # D2H-bound at 54 k pictures/s (configs[4] product on a 25 GB source, tools/full_config1.py --config5).  The pool
        # is created on first use.
        self._copy_threads = max(1, int(os.environ.get("VT_INGEST_COPY_THREADS", "4")))
        self._copy_pool = None

    # ------------------------------------------------------------------------------------------------------
    def _register_source(self) -> None:
        """Map the bitstream file privately and page-lock the mapping, so that the copy-in stream reads the page cache
        directly and the staging memcpy (one more pass over the host's memory per batch) disappears.  Off with
        VT_INGEST_DIRECT_H2D=0; files over VT_INGEST_DIRECT_MAX_GB (default 16) and file systems whose mappings cannot be
        pinned keep the staging path -- and so do files under VT_INGEST_DIRECT_MIN_MB (default 128): page-locking costs
        1-4 GB/s up front on this box (33 ms for a 34 MB clip, four times that clip's whole GPU pass), the staging copy
        8 GB/s on a helper thread beside the GPU, so pinning only pays for files that are cut into many segments."""
        import mmap
        if os.environ.get("VT_INGEST_DIRECT_H2D", "1") != "1" or self.dev.type != "cuda":
            return
        try:
            size = os.path.getsize(self.idx.path)
            if size == 0 or size > float(os.environ.get("VT_INGEST_DIRECT_MAX_GB", "16")) * (1 << 30) \
                    or size < float(os.environ.get("VT_INGEST_DIRECT_MIN_MB", "128")) * (1 << 20):
                return
            fd = os.open(self.idx.path, os.O_RDONLY)
            try:
                mm = mmap.mmap(fd, size, flags=mmap.MAP_PRIVATE, prot=mmap.PROT_READ | mmap.PROT_WRITE)
            finally:
                os.close(fd)
        except (OSError, ValueError):
            return
        arr = np.frombuffer(mm, dtype=np.uint8)
        if lib().vt_host_register_source(c_void_p(arr.ctypes.data), size) != 0:
            del arr
            mm.close()
            return
        self._src_map = (mm, arr, arr.ctypes.data)

    def release(self) -> None:
        """Give the pipeline buffers back to the pool (the engine must not run afterwards) and unpin the source."""
        if self.slots:
            torch.cuda.synchronize(self.dev)
            _SLOT_POOL.setdefault(self._pool_key, []).extend(self.slots)
            del _SLOT_POOL[self._pool_key][6:]           # keep at most two engines' worth per geometry
            self.slots = []
        self.close()

    def close(self) -> None:
        for name in ("_pool", "_copy_pool"):
            pool = getattr(self, name, None)
            if pool is not None:
                pool.shutdown(wait=False)
                setattr(self, name, None)
        if self._src_map is not None:
            mm, arr, base = self._src_map
            self._src_map = None
            lib().vt_host_unregister(c_void_p(base))
            del arr
            try:
                mm.close()
            except (BufferError, ValueError):
                pass

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _stage_bitstream(self, slot, b0: int, b1: int):
        """Copy the file bytes pictures [b0,b1) need into the slot's pinned buffer; return payload offsets
        relative to the device copy (or NO_PAYLOAD for pictures that repeat one from before b0)."""
        pay = np.full(b1 - b0, _NO_PAYLOAD, np.uint64)
        idr = [k for k in range(b0, b1) if self.idx.keyframe[k]]
        if not idr:
            return pay, 0
        lo = int(self.offs[idr[0]])
        hi = int(self.offs[idr[-1]]) + int(self.sizes[idr[-1]])
        nbytes = hi - lo
        if nbytes > self.bs_cap:
            raise _lib.VtError(_lib.VT_ERR_NOMEM, "batch bitstream %d B exceeds staging %d B" % (nbytes, self.bs_cap))
        slot["src_lo"] = lo
        if self._src_map is not None:        # the copy engine reads the file's pages itself
            first_idr = idr[0]
            pay[first_idr - b0:] = self.payload[first_idr:b1] - np.uint64(lo)
            return pay, nbytes
        if slot["bs_host"] is None:
            slot["bs_host"] = torch.empty(self.bs_cap, dtype=torch.uint8, pin_memory=True)
        dst = slot["bs_host"].numpy()
        if self._copy_threads > 1 and nbytes >= (4 << 20):
            if self._copy_pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._copy_pool = ThreadPoolExecutor(max_workers=self._copy_threads)
            # PCM-intra streams are uncompressed (115 KB per 1080p picture on average): one thread copies ~8 GB/s out of
            # the page cache, which would bound the device-sink path; numpy releases the GIL, so split the copy
            n = self._copy_threads
            step = (nbytes + n - 1) // n
            futs = [self._copy_pool.submit(np.copyto, dst[a:min(a + step, nbytes)], self.host[lo + a:lo + min(a + step, nbytes)])
                    for a in range(0, nbytes, step)]
            for f in futs:
                f.result()
        else:
            dst[:nbytes] = self.host[lo:hi]
        first_idr = idr[0]
        pay[first_idr - b0:] = self.payload[first_idr:b1] - np.uint64(lo)
        return pay, nbytes

    def kept_pictures(self, first: int, last: int) -> int:
        """How many output frames run(first, last) delivers."""
        if not self.opts.keep_frames:
            return 0
        se = self.opts.sample_every
        return last - first if se == 1 else len(range(-(-first // se) * se, last, se))

    def run(self, first: int, last: int, sink=None, device_sink=None, landing=None, on_primed=None) -> IngestResult:
        """Process pictures [first, last).  Output frames go to sink(chunk_host_tensor, first_picture).

        landing (landing.Landing), if given, is the `.frames` file of the segment: with a registered mapping the copy
        engine writes every batch's frames straight into the file at their final position (no pinned staging, no
        host copy, nothing on the drain path); otherwise batches go through the pinned ring to the landing's writer
        thread.

        device_sink(chunk_device_tensor, first_picture), if given, receives every batch's frames ON THE DEVICE instead:
        it is called with the compute stream current, right after the kernels that produce the chunk were enqueued,
        so whatever it enqueues on the current stream is ordered after them and before the buffers are reused; the
        frames are then not copied to the host at all (scores still are).  This is the hand-off for a consumer that
        lives on the GPU (the model's own input pipeline): the D2H copy is what bounds the host-buffer path.

        on_primed(), if given, is called once the pipeline is full (one batch enqueued per slot) or the pass has nothing
        more to enqueue: the moment from which this thread mostly sleeps on events and other Python threads of the
        caller (the stream copy) can have the interpreter without delaying the GPU."""
        n_total = self.idx.n_frames
        try:
            with torch.cuda.device(self.dev):
                return self._run(first, last, sink, device_sink, landing, n_total, on_primed)
        finally:
            if on_primed is not None:
                on_primed()

    def _run(self, first, last, sink, device_sink, landing, n_total, on_primed=None) -> IngestResult:
        if not (0 <= first < last <= n_total):
            raise ValueError("picture range [%d,%d) outside the stream (%d pictures)" % (first, last, n_total))
        L = lib()
        import time as _time
        host_t = [_time.perf_counter()]
        # lead-in: the score of picture `first` needs mafd of picture first-1, i.e. SAD(first-1, first-2), so
        # decoding starts at the keyframe that picture first-2 depends on ("decode 2 extra
        # lead-in frames per shard" instead of communicating)
        r0 = max(first - 1, 0)               # first picture whose SAD is reported internally
        need = max(first - 2, 0)
        kf = self.keyframes[self.keyframes <= need]
        if kf.size == 0:
            raise _lib.VtError(_lib.VT_ERR_BITSTREAM, "no keyframe at or before picture %d" % need)
        d0 = int(kf[-1])
        n_rep = last - r0
        sad_all = np.zeros(n_rep, np.uint64)
        hist_all = np.zeros((n_rep, 256), np.uint32)
        B = self.B
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        prev_surface = None
        batches = [(b, min(b + B, last)) for b in range(d0, last, B)]
        fb = self.frame_bytes
        direct = landing is not None and landing.direct and device_sink is None
        land_staged = landing is not None and not direct and device_sink is None
        landed = 0                           # frames already given a place in the landing file

        def drain(slot):
            w = slot.get("wfut")
            if w is not None:                # the writer thread still reads this slot's pinned frames
                w.result()
                slot["wfut"] = None
            p = slot["pending"]
            if p is None:
                return
            slot["ev_out"].synchronize()
            b0, b1 = p
            lo, hi = max(b0, r0), b1
            if hi > lo:
                k0 = lo - b0
                sad_all[lo - r0:hi - r0] = slot["sad_host"].numpy()[k0:k0 + hi - lo].view(np.uint64)
                hist_all[lo - r0:hi - r0] = slot["hist_host"].numpy()[k0:k0 + hi - lo].view(np.uint32)
            if land_staged and slot["land"] is not None:
                at, cnt, row0 = slot["land"]
                slot["wfut"] = landing.write_chunk(slot["out_host"][row0:row0 + cnt], at * fb)
                slot["land"] = None
            if sink is not None and self.opts.keep_frames and device_sink is None:
                keep = slot["kept"]
                if keep is None:                              # every picture from max(b0, first) on, in place
                    lo = max(b0, first)
                    if hi > lo:
                        sink(slot["out_host"][lo - b0:hi - b0], lo)
                elif len(keep):                               # sampled pictures, packed from row 0
                    sink(slot["out_host"][:len(keep)], keep[0])
            slot["pending"] = None

        # The bitstream of batch i+1 is copied into its pinned staging buffer by a helper thread (numpy releases the
        # GIL) while this thread enqueues batch i: the host side then keeps up with the D2H copy engine.
        def stage(j):
            sl = self.slots[j % self.n_slots]
            sl["ev_in"].synchronize()        # the previous H2D out of this staging buffer has completed
            return self._stage_bitstream(sl, *batches[j])

        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=1)
        staged = self._pool.submit(stage, 0) if (batches and self.stage_thread) else None
        trace = [] if os.environ.get("VT_INGEST_TRACE") == "1" else None   # tools/trace_unit.py: per-batch stream timeline

        def mark(stream):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e

        host_t.append(_time.perf_counter())
        for i, (b0, b1) in enumerate(batches):
            slot = self.slots[i % self.n_slots]
            nb = b1 - b0
            if i == 1:
                host_t.append(_time.perf_counter())
            if self.stage_thread:
                pay, nbytes = staged.result()
                drain(slot)                  # host may only reuse pinned/device buffers whose copies completed
                if i + 1 < len(batches):
                    staged = self._pool.submit(stage, i + 1)  # only touches that slot's pinned bitstream buffer
            else:
                drain(slot)
                pay, nbytes = self._stage_bitstream(slot, b0, b1)
            with torch.cuda.stream(self.s_in):
                if trace is not None:
                    trace.append([mark(self.s_in)])
                if nbytes:
                    if self._src_map is not None:
                        check(L.vt_copy_to_device_async(c_void_p(slot["bs_dev"].data_ptr()),
                                                        c_void_p(self._src_map[2] + slot["src_lo"]), nbytes,
                                                        c_void_p(self.s_in.cuda_stream)))
                    else:
                        slot["bs_dev"][:nbytes].copy_(slot["bs_host"][:nbytes], non_blocking=True)
                    self.h2d_bytes += nbytes
                slot["ev_in"].record(self.s_in)
                if trace is not None:
                    trace[-1].append(mark(self.s_in))
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(slot["ev_in"])
                if trace is not None:
                    trace[-1].append(mark(self.s_cmp))
                st = c_void_p(self.s_cmp.cuda_stream)
                surf = slot["surf"]
                prev_p = c_void_p(prev_surface.data_ptr()) if prev_surface is not None else None
                # the common case (every picture kept, planar output) is ONE C call per batch: decode -> score -> frames
                fused = self.opts.sample_every == 1 and self.rgb_plan is None
                if fused:
                    check(L.vt_ingest_batch_pcm(self.plan._h if self.plan is not None else None,
                                                c_void_p(slot["bs_dev"].data_ptr()), pay.ctypes.data, nb, self.w, self.h,
                                                prev_p, c_void_p(surf.data_ptr()), self.pitch, self.surface_bytes,
                                                c_void_p(slot["sad"].data_ptr()), c_void_p(slot["hist"].data_ptr()),
                                                c_void_p(slot["out"].data_ptr()) if self.opts.keep_frames else None,
                                                self.frame_bytes, st))
                else:
                    check(L.vt_h264_pcm_decode(c_void_p(slot["bs_dev"].data_ptr()), pay.ctypes.data, nb, self.w, self.h,
                                               prev_p, c_void_p(surf.data_ptr()), self.pitch, self.surface_bytes, st))
                    check(L.vt_sad_hist_u8(c_void_p(surf.data_ptr()), self.pitch, self.surface_bytes, self.w, self.h,
                                           prev_p, nb, c_void_p(slot["sad"].data_ptr()),
                                           c_void_p(slot["hist"].data_ptr()), st))
                kept = None
                src_t, n_conv, out_row0 = surf, nb, 0
                if self.opts.keep_frames and self.opts.sample_every > 1:
                    # sampled output: gather the kept surfaces (K5) and convert only those, packed from row 0
                    kept = [k for k in range(max(b0, first), b1) if k % self.opts.sample_every == 0]
                    n_conv = len(kept)
                    if n_conv:
                        slot["sel_idx_host"][:n_conv] = torch.tensor([k - b0 for k in kept], dtype=torch.int32)
                        slot["sel_idx"][:n_conv].copy_(slot["sel_idx_host"][:n_conv], non_blocking=True)
                        check(L.vt_gather_frames(c_void_p(surf.data_ptr()), self.surface_bytes, self.surface_bytes,
                                                 c_void_p(slot["sel_idx"].data_ptr()), n_conv,
                                                 c_void_p(slot["sel_surf"].data_ptr()), st))
                        src_t = slot["sel_surf"]
                slot["kept"] = kept
                if self.opts.keep_frames and n_conv and not fused:
                    if self.rgb_plan is not None:
                        check(L.vt_scale_nv12_to_rgb24(self.rgb_plan._h, c_void_p(src_t.data_ptr()), self.pitch,
                                                       self.surface_bytes, c_void_p(slot["out"].data_ptr()),
                                                       self.frame_bytes, n_conv, st))
                    elif self.plan is not None:
                        check(L.vt_scale_nv12_to_yuv420p(self.plan._h, c_void_p(src_t.data_ptr()), self.pitch,
                                                         self.surface_bytes, c_void_p(slot["out"].data_ptr()),
                                                         self.frame_bytes, n_conv, st))
                    else:
                        check(L.vt_nv12_to_yuv420p(c_void_p(src_t.data_ptr()), self.pitch, self.surface_bytes, self.w,
                                                   self.h, c_void_p(slot["out"].data_ptr()), self.frame_bytes, n_conv, st))
                if device_sink is not None and self.opts.keep_frames:
                    if kept is not None:
                        if len(kept):
                            device_sink(slot["out"][:len(kept)], kept[0])
                    elif b1 > max(b0, first):
                        device_sink(slot["out"][max(b0, first) - b0:nb], max(b0, first))
                slot["ev_cmp"].record(self.s_cmp)
                if trace is not None:
                    trace[-1].append(mark(self.s_cmp))
                prev_surface = surf[nb - 1]
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot["ev_cmp"])
                if trace is not None:
                    trace[-1].append(mark(self.s_out))
                lo = max(b0, first)
                slot["land"] = None
                cnt = row0 = 0
                if device_sink is None and self.opts.keep_frames:     # else: scores only / handed over on the device
                    if kept is not None:
                        cnt = len(kept)                                # sampled pictures, packed from row 0
                    elif b1 > lo:
                        cnt, row0 = b1 - lo, lo - b0
                if cnt:
                    if direct:                            # straight into the file's registered mapping
                        landing.tensor[landed * fb:(landed + cnt) * fb].view(cnt, fb).copy_(
                            slot["out"][row0:row0 + cnt], non_blocking=True)
                    else:
                        if slot["out_host"] is None:
                            slot["out_host"] = torch.empty((B, fb), dtype=torch.uint8, pin_memory=True)
                        slot["out_host"][row0:row0 + cnt].copy_(slot["out"][row0:row0 + cnt], non_blocking=True)
                        if land_staged:
                            slot["land"] = (landed, cnt, row0)
                    landed += cnt
                    self.d2h_bytes += cnt * fb
                slot["sad_host"][:nb].copy_(slot["sad"][:nb], non_blocking=True)
                slot["hist_host"][:nb].copy_(slot["hist"][:nb], non_blocking=True)
                self.d2h_bytes += nb * (8 + 1024)
                slot["ev_out"].record(self.s_out)
                if trace is not None:
                    trace[-1].append(mark(self.s_out))
            # drain() synchronises this slot's ev_out before the host issues batch i+2 into the same buffers,
            # which orders every device-side reuse (bitstream, surfaces, output) after the copies that read them
            slot["pending"] = (b0, b1)
            if on_primed is not None and i == self.n_slots - 1:
                on_primed()
            if land_staged and i >= 1:
                drain(self.slots[(i - 1) % self.n_slots])    # hand batch i-1 to the writer while batch i runs
        host_t.append(_time.perf_counter())
        for slot in sorted(self.slots, key=lambda sl: sl["pending"][0] if sl["pending"] else -1):
            drain(slot)                      # oldest batch first: the sink sees pictures in order
        for slot in self.slots:
            drain(slot)                      # writer futures of the last batches
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_out)
        host_t.append(_time.perf_counter())
        self.last_host_ms = [round((t - host_t[0]) * 1e3, 3) for t in host_t]   # enter, loop, batch 1, drain, drained
        if trace:
            t0 = trace[0][0]
            self.last_trace = [[t0.elapsed_time(e) for e in row] for row in trace]   # ms: h2d, cmp, d2h start/end
        if r0 == 0:
            sad_all[0] = 0                   # picture 0 has no predecessor
        scores = scene.scene_scores(sad_all, self.w, self.h)
        if first > 0:                        # element 0 was picture first-1: only its mafd was needed
            sad_all, hist_all, scores = sad_all[1:], hist_all[1:], scores[1:]
        cuts = scene.select_cuts(scores, self.opts.scene_threshold, first)
        return IngestResult(first, last, self.out_w, self.out_h, self.frame_bytes, sad_all, hist_all, scores, cuts,
                            {"decoded_from": d0, "batches": len(batches), "h2d_bytes": self.h2d_bytes,
                             "d2h_bytes": self.d2h_bytes, "landed_frames": landed,
                             "landing": ("direct" if direct else "staged" if land_staged else None)})

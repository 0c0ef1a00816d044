"""Decode front end (K0): Annex-B H.264 in host memory -> NV12 surfaces in HBM.

Replaces libavcodec inside the reference's ffmpeg child process
(/root/reference/src/utils/video_segmenter.py:141-154).  NVDEC is the intended engine; this pool's driver
refuses it, so streams must be in the PCM-intra subset that the CUDA kernel in csrc/vt_h264.cu
decodes.  Anything else raises VtError(VT_ERR_UNSUPPORTED): there is no CPU decode path.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from math import gcd

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def nvdec_available() -> tuple[bool, str]:
    n, w, h = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().vt_nvdec_probe(ctypes.byref(n), ctypes.byref(w), ctypes.byref(h))
    if rc == 0:
        return True, "%d NVDEC engine(s), max %dx%d" % (n.value, w.value, h.value)
    return False, (lib().vt_last_error() or b"").decode()


class H264PcmDecoder:
    """Holds one elementary stream: host index (offsets, flags) + the bytes uploaded once to HBM."""

    def __init__(self, bitstream: bytes | np.ndarray, device: torch.device | str = "cuda", upload: bool = True):
        self.host = np.frombuffer(bitstream, np.uint8) if not isinstance(bitstream, np.ndarray) else bitstream
        L = lib()
        info = _lib.StreamInfo()
        check(L.vt_h264_scan(self.host.ctypes.data, self.host.size, ctypes.byref(info), None, None, None, 0))
        n = info.n_frames
        self.offsets = np.zeros(max(n, 1), np.uint64)
        self.sizes = np.zeros(max(n, 1), np.uint32)
        self.flags = np.zeros(max(n, 1), np.uint32)
        check(L.vt_h264_scan(self.host.ctypes.data, self.host.size, ctypes.byref(info), self.offsets.ctypes.data,
                             self.sizes.ctypes.data, self.flags.ctypes.data, n))
        self.info = info
        self.width, self.height, self.n_frames = info.width, info.height, n
        g = gcd(info.fps_num, info.fps_den) or 1
        self.fps_num, self.fps_den = info.fps_num // g, info.fps_den // g
        self.keyframes = np.nonzero(self.flags[:n] & 1)[0]
        if not info.pcm_intra_only:
            raise _lib.VtError(_lib.VT_ERR_UNSUPPORTED,
                               "stream uses H.264 tools outside the PCM-intra subset; NVDEC is required and "
                               "this driver does not expose it")
        self.payload = np.zeros(n, np.uint64)
        check(L.vt_h264_pcm_layout(self.host.ctypes.data, self.host.size, self.offsets.ctypes.data,
                                   self.sizes.ctypes.data, n, self.payload.ctypes.data))
        self.device = torch.device(device)
        self.dev = None
        if upload:
            self.upload()

    def upload(self) -> None:
        # 64 bytes of slack: the kernel reads whole 16-byte groups around each macroblock row
        self.dev = torch.zeros(self.host.size + 64, dtype=torch.uint8, device=self.device)
        import warnings
        with warnings.catch_warnings():          # a read-only bitstream view is fine: it is only copied from
            warnings.simplefilter("ignore", UserWarning)
            self.dev[: self.host.size].copy_(torch.from_numpy(self.host), non_blocking=True)

    @property
    def duration(self) -> float:
        return self.n_frames * self.fps_den / self.fps_num if self.fps_num else 0.0

    def surface_rows(self) -> int:
        return self.height + (self.height + 1) // 2

    def decode(self, first: int, count: int, pitch: int | None = None, prev: torch.Tensor | None = None,
               out: torch.Tensor | None = None) -> torch.Tensor:
        """Decode pictures [first, first+count) into NV12 surfaces -> uint8 CUDA tensor (count, rows, pitch)."""
        if first < 0 or count <= 0 or first + count > self.n_frames:
            raise ValueError("picture range outside the stream")
        pitch = pitch or ((self.width + 15) // 16 * 16)
        rows = self.surface_rows()
        if out is None:
            out = torch.empty((count, rows, pitch), dtype=torch.uint8, device=self.device)
        pay = self.payload[first:first + count].copy()
        # pictures that repeat an IDR from before `first` read from the carried-over surface
        kf = self.keyframes[self.keyframes >= first]
        first_idr = int(kf[0]) if kf.size else first + count
        pay[: max(0, min(count, first_idr - first))] = np.uint64(0xFFFFFFFFFFFFFFFF)
        if first_idr > first and prev is None and self.payload[first] != np.uint64(0xFFFFFFFFFFFFFFFF):
            # no surface carried over: fall back on the referenced IDR's samples, which are still in the stream
            pay[: first_idr - first] = self.payload[first]
        check(lib().vt_h264_pcm_decode(c_void_p(self.dev.data_ptr()), pay.ctypes.data, count, self.width,
                                       self.height, c_void_p(prev.data_ptr()) if prev is not None else None,
                                       c_void_p(out.data_ptr()), pitch, rows * pitch,
                                       c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out


class NvdecSession:
    """K0 proper: an NVDEC session (vt_decode_open / feed / next_surface / release_surface / close).

    Raises VtError(VT_ERR_NVDEC) where the driver exposes no video decode (this pool: see DESIGN.md section 2).
    Surfaces are mapped device memory in the NV12 pitch-linear layout the kernels consume; `surfaces()` yields
    (device pointer, pitch, width, height, surface rows, pts) and releases each surface when the caller asks for the
    next one."""

    H264, HEVC, VP9, AV1 = 4, 8, 9, 11

    def __init__(self, codec: int = 4, max_surfaces: int = 8, stream=None):
        import ctypes
        self._h = ctypes.c_void_p()
        st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else None
        _lib.check(_lib.lib().vt_decode_open(codec, max_surfaces, st, ctypes.byref(self._h)))

    def feed(self, data: bytes, pts: int = 0, end_of_stream: bool = False) -> None:
        import ctypes
        buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data) if data else None
        _lib.check(_lib.lib().vt_decode_feed(self._h, buf, len(data), pts, 1 if end_of_stream else 0))

    def surfaces(self):
        import ctypes
        L = _lib.lib()
        held = None
        while True:
            ptr, pitch, w, h, rows = ctypes.c_uint64(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            pts = ctypes.c_int64()
            rc = _lib.check(L.vt_decode_next_surface(self._h, ctypes.byref(ptr), ctypes.byref(pitch), ctypes.byref(w),
                                                     ctypes.byref(h), ctypes.byref(rows), ctypes.byref(pts)))
            if held is not None:
                _lib.check(L.vt_decode_release_surface(self._h, held))
                held = None
            if rc == 1:
                return
            held = ptr.value
            yield ptr.value, pitch.value, w.value, h.value, rows.value, pts.value

    def close(self) -> None:
        if self._h:
            _lib.lib().vt_decode_close(self._h)
            self._h = None

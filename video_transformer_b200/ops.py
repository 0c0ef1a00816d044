"""Tensor-level wrappers over the C ABI.  torch is only the owner of device memory and streams here."""
from __future__ import annotations

from ctypes import byref, c_int, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import SWS_AREA, SWS_BICUBIC, SWS_BILINEAR, VtError, check, lib  # noqa: F401


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
        raise ValueError("%s must be a contiguous uint8 CUDA tensor" % name)


def nv12_frame_bytes(pitch: int, h: int) -> int:
    return pitch * (h + (h + 1) // 2)


def make_filter(src: int, dst: int, flags: int = SWS_BICUBIC, one: int = 1 << 14):
    """Host-side libswscale-exact filter bank -> (coef[dst, taps] int16, pos[dst] int32, taps)."""
    L = lib()
    cap = max(check(L.vt_sws_max_taps(src, dst, flags)), 4)
    coef = np.zeros((dst, cap), np.int16)
    pos = np.zeros(dst, np.int32)
    taps = c_int(0)
    check(L.vt_sws_make_filter(src, dst, flags, one, coef.ctypes.data, pos.ctypes.data, byref(taps)))
    t = taps.value
    return np.ascontiguousarray(coef.reshape(-1)[: dst * t].reshape(dst, t)), pos, t


def scale_width_for_height(src_w: int, src_h: int, dst_h: int) -> int:
    return check(lib().vt_scale_width_for_height(src_w, src_h, dst_h))


def sad_hist(luma: torch.Tensor, w: int, h: int, pitch: int, frame_stride: int, n_frames: int,
             prev0: torch.Tensor | None = None):
    """K3 on a batch of frames laid out `frame_stride` apart in `luma` (flat uint8 CUDA tensor).

    Returns (sad[n] int64 CUDA tensor holding u64 values, hist[n,256] int32 CUDA tensor holding u32 values).
    """
    _need_cuda(luma, "luma")
    sad = torch.empty(n_frames, dtype=torch.int64, device=luma.device)
    hist = torch.empty((n_frames, 256), dtype=torch.int32, device=luma.device)
    p0 = c_void_p(prev0.data_ptr()) if prev0 is not None else None
    check(lib().vt_sad_hist_u8(c_void_p(luma.data_ptr()), pitch, frame_stride, w, h, p0, n_frames,
                               c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), _stream()))
    return sad, hist


def nv12_to_yuv420p(src: torch.Tensor, w: int, h: int, pitch: int, n_frames: int,
                    src_frame_stride: int | None = None) -> torch.Tensor:
    _need_cuda(src, "src")
    sfs = src_frame_stride or nv12_frame_bytes(pitch, h)
    fb = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
    dst = torch.empty((n_frames, fb), dtype=torch.uint8, device=src.device)
    check(lib().vt_nv12_to_yuv420p(c_void_p(src.data_ptr()), pitch, sfs, w, h, c_void_p(dst.data_ptr()), fb,
                                   n_frames, _stream()))
    return dst


def nv12_to_rgb24(src: torch.Tensor, w: int, h: int, pitch: int, n_frames: int,
                  src_frame_stride: int | None = None) -> torch.Tensor:
    _need_cuda(src, "src")
    sfs = src_frame_stride or nv12_frame_bytes(pitch, h)
    dst = torch.empty((n_frames, h, w, 3), dtype=torch.uint8, device=src.device)
    check(lib().vt_nv12_to_rgb24(c_void_p(src.data_ptr()), pitch, sfs, w, h, c_void_p(dst.data_ptr()), w * h * 3,
                                 n_frames, _stream()))
    return dst


def gather_frames(src: torch.Tensor, frame_stride: int, frame_bytes: int, index: torch.Tensor | None, count: int,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    _need_cuda(src, "src")
    if out is None:
        out = torch.empty((count, frame_bytes), dtype=torch.uint8, device=src.device)
    ip = c_void_p(index.data_ptr()) if index is not None else None
    check(lib().vt_gather_frames(c_void_p(src.data_ptr()), frame_stride, frame_bytes, ip, count,
                                 c_void_p(out.data_ptr()), _stream()))
    return out


class ScalePlan:
    """Device-resident libswscale-exact filter banks for one (src size -> dst size, flags)."""

    def __init__(self, sw: int, sh: int, dw: int, dh: int, flags: int = SWS_BICUBIC):
        self.sw, self.sh, self.dw, self.dh, self.flags = sw, sh, dw, dh, flags
        self.cdw, self.cdh = (dw + 1) // 2, (dh + 1) // 2
        self._h = c_void_p()
        check(lib().vt_scale_plan_create(sw, sh, dw, dh, flags, byref(self._h)))

    def close(self) -> None:
        if self._h:
            lib().vt_scale_plan_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def stream_info(self, chroma: bool = False) -> dict:
        """Which kernel scale_nv12 uses for a plane kind (streaming TMA kernel or the generic two-pass one)."""
        arr = (c_int * 8)()
        check(lib().vt_scale_plan_stream_info(self._h, 1 if chroma else 0, arr))
        keys = ("streaming", "dp2a_pairs", "v_taps", "cols_per_lane", "rows_out", "tile_w", "tile_h", "warp_smem")
        return dict(zip(keys, [int(v) for v in arr]))

    @property
    def out_frame_bytes(self) -> int:
        return self.dw * self.dh + 2 * self.cdw * self.cdh

    def scale_plane(self, src: torch.Tensor, chroma: bool = False) -> torch.Tensor:
        """One planar 8-bit plane (H x W uint8 CUDA tensor) through the generic kernels."""
        _need_cuda(src, "src")
        dw, dh = (self.cdw, self.cdh) if chroma else (self.dw, self.dh)
        dst = torch.empty((dh, dw), dtype=torch.uint8, device=src.device)
        check(lib().vt_scale_plane_u8(self._h, 1 if chroma else 0, c_void_p(src.data_ptr()), src.stride(0),
                                      c_void_p(dst.data_ptr()), dw, _stream()))
        return dst

    def scale_nv12(self, src: torch.Tensor, pitch: int, n_frames: int, src_frame_stride: int | None = None,
                   out: torch.Tensor | None = None) -> torch.Tensor:
        _need_cuda(src, "src")
        sfs = src_frame_stride or nv12_frame_bytes(pitch, self.sh)
        if out is None:
            out = torch.empty((n_frames, self.out_frame_bytes), dtype=torch.uint8, device=src.device)
        check(lib().vt_scale_nv12_to_yuv420p(self._h, c_void_p(src.data_ptr()), pitch, sfs,
                                             c_void_p(out.data_ptr()), self.out_frame_bytes, n_frames, _stream()))
        return out


    @property
    def fuses_score(self) -> bool:
        """True when scale_score_nv12 runs SAD/histogram inside the luma pass of the scaler (one read of the source)."""
        return bool(lib().vt_scale_plan_fuses_score(self._h))

    def scale_score_nv12(self, src: torch.Tensor, pitch: int, n_frames: int, src_frame_stride: int | None = None,
                         prev0: torch.Tensor | None = None, out: torch.Tensor | None = None):
        """K2 + K3: (frames, sad[n] int64, hist[n,256] int32) of a batch of NV12 surfaces."""
        _need_cuda(src, "src")
        sfs = src_frame_stride or nv12_frame_bytes(pitch, self.sh)
        if out is None:
            out = torch.empty((n_frames, self.out_frame_bytes), dtype=torch.uint8, device=src.device)
        sad = torch.empty(n_frames, dtype=torch.int64, device=src.device)
        hist = torch.empty((n_frames, 256), dtype=torch.int32, device=src.device)
        p0 = c_void_p(prev0.data_ptr()) if prev0 is not None else None
        check(lib().vt_scale_score_nv12_to_yuv420p(self._h, c_void_p(src.data_ptr()), pitch, sfs, p0,
                                                   c_void_p(out.data_ptr()), self.out_frame_bytes, n_frames,
                                                   c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), _stream()))
        return out, sad, hist


class RgbPlan:
    """Device-resident filter banks for NV12 -> RGB24 at (dw, dh): `ffmpeg -vf scale=dw:dh -pix_fmt rgb24` semantics."""

    def __init__(self, sw: int, sh: int, dw: int, dh: int, flags: int = SWS_BICUBIC):
        self.sw, self.sh, self.dw, self.dh, self.flags = sw, sh, dw, dh, flags
        self._h = c_void_p()
        check(lib().vt_rgb_plan_create(sw, sh, dw, dh, flags, byref(self._h)))

    def close(self) -> None:
        if self._h:
            lib().vt_rgb_plan_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def scale_nv12(self, src: torch.Tensor, pitch: int, n_frames: int, src_frame_stride: int | None = None,
                   out: torch.Tensor | None = None) -> torch.Tensor:
        _need_cuda(src, "src")
        sfs = src_frame_stride or nv12_frame_bytes(pitch, self.sh)
        if out is None:
            out = torch.empty((n_frames, self.dh, self.dw, 3), dtype=torch.uint8, device=src.device)
        check(lib().vt_scale_nv12_to_rgb24(self._h, c_void_p(src.data_ptr()), pitch, sfs, c_void_p(out.data_ptr()),
                                           self.dw * self.dh * 3, n_frames, _stream()))
        return out


class JpegPlan:
    """Baseline-JPEG encoder for planar YUV420P pictures of one size (the Motion-JPEG samples of the upload reducer)."""

    def __init__(self, w: int, h: int, quality: int = 75, expand_range: bool = True):
        self.w, self.h, self.quality, self.expand_range = w, h, quality, expand_range
        self._h = c_void_p()
        check(lib().vt_jpeg_plan_create(w, h, quality, 1 if expand_range else 0, byref(self._h)))
        self.frame_bytes_in = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
        self.max_frame_bytes = int(lib().vt_jpeg_max_frame_bytes(self._h))

    def close(self) -> None:
        if self._h:
            lib().vt_jpeg_plan_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def encode(self, frames: torch.Tensor, out: torch.Tensor | None = None):
        """frames: [n, frame_bytes_in] uint8 CUDA tensor -> (packed bytes CUDA tensor, offsets int64 CUDA tensor [n+1],
        status int32 CUDA tensor [1]).  Asynchronous on the current stream; offsets[n] is the number of valid bytes."""
        _need_cuda(frames, "frames")
        n = frames.shape[0]
        if out is None:
            out = torch.empty(n * max(self.frame_bytes_in // 2, 65536), dtype=torch.uint8, device=frames.device)
        offsets = torch.empty(n + 1, dtype=torch.int64, device=frames.device)
        status = torch.empty(1, dtype=torch.int32, device=frames.device)
        check(lib().vt_jpeg_encode_yuv420p(self._h, c_void_p(frames.data_ptr()), frames.stride(0), n,
                                           c_void_p(out.data_ptr()), out.numel(), c_void_p(offsets.data_ptr()),
                                           c_void_p(status.data_ptr()), _stream()))
        return out, offsets, status

    def encode_to_host(self, frames: torch.Tensor) -> list[bytes]:
        """Synchronous convenience: one `bytes` per picture.  Raises when the encoder refused (status != 0)."""
        out, offsets, status = self.encode(frames)
        off = offsets.cpu().numpy()
        st = int(status.cpu()[0])
        if st != 0:
            raise VtError(_lib.VT_ERR_UNSUPPORTED, "JPEG encoder refused the batch (status %d: %s)" % (
                st, "an MCU row did not compress below its raw size" if st == 1 else "output buffer too small"))
        host = out[: int(off[-1])].cpu().numpy()
        return [host[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]

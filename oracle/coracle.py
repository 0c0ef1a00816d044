"""ctypes front end of oracle/libvtoracle.so (the C restatement in vt_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

BILINEAR, BICUBIC, AREA = 2, 4, 0x20


def build() -> str:
    so = os.path.join(_HERE, "libvtoracle.so")
    src = os.path.join(_HERE, "vt_oracle.c")
    src2 = os.path.join(_HERE, "jpeg_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(src2)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        u8p = ctypes.c_void_p
        L.vto_sws_max_taps.argtypes = [ctypes.c_int] * 3
        L.vto_sws_make_filter.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, u8p,
                                          ctypes.POINTER(ctypes.c_int)]
        L.vto_sws_scale_plane.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.vto_sad_hist.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, u8p]
        L.vto_sad_hist.restype = None
        L.vto_sad_hist_fast.argtypes = L.vto_sad_hist.argtypes
        L.vto_sad_hist_fast.restype = None
        L.vto_nv12_to_yuv420p.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, u8p, u8p]
        L.vto_nv12_to_yuv420p.restype = None
        L.vto_pcm_picture_to_yuv420p.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, u8p, u8p]
        L.vto_pcm_picture_to_yuv420p.restype = None
        L.vto_yuv_to_rgb24.argtypes = [u8p, ctypes.c_int, u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       u8p, ctypes.c_int, ctypes.c_int]
        L.vtj_encode.argtypes = [u8p, ctypes.c_int, u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, u8p, ctypes.c_size_t]
        L.vtj_encode.restype = ctypes.c_size_t
        L.vtj_quant_table.argtypes = [ctypes.c_int, ctypes.c_int, u8p]
        L.vtj_quant_table.restype = None
        L.vtj_block_coefficients.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p,
                                             ctypes.c_int, u8p]
        L.vtj_block_coefficients.restype = None
        _lib = L
    return _lib


def make_filter(src_w: int, dst_w: int, flags: int = BICUBIC, one: int = 1 << 14):
    L = lib()
    cap = max(L.vto_sws_max_taps(src_w, dst_w, flags), 4)
    coef = np.zeros((dst_w, cap), np.int16)
    pos = np.zeros(dst_w, np.int32)
    taps = ctypes.c_int(0)
    rc = L.vto_sws_make_filter(src_w, dst_w, flags, one, coef.ctypes.data, pos.ctypes.data, ctypes.byref(taps))
    if rc:
        raise RuntimeError("vto_sws_make_filter failed")
    t = taps.value
    return np.ascontiguousarray(coef.reshape(-1)[: dst_w * t].reshape(dst_w, t)), pos, t


def scale_plane(src: np.ndarray, dw: int, dh: int, flags: int = BICUBIC) -> np.ndarray:
    src = np.ascontiguousarray(src)
    sh, sw = src.shape
    dst = np.zeros((dh, dw), np.uint8)
    rc = lib().vto_sws_scale_plane(src.ctypes.data, sw, sh, src.strides[0], dst.ctypes.data, dw, dh, dw, flags)
    if rc:
        raise RuntimeError("vto_sws_scale_plane failed")
    return dst


def scale_yuv420p(y, u, v, dw, dh, flags=BICUBIC):
    cw, ch = -(-dw // 2), -(-dh // 2)
    return scale_plane(y, dw, dh, flags), scale_plane(u, cw, ch, flags), scale_plane(v, cw, ch, flags)


def sad_hist(cur: np.ndarray, prev: np.ndarray | None, fast: bool = False):
    cur = np.ascontiguousarray(cur)
    h, w = cur.shape
    sad = np.zeros(1, np.uint64)
    hist = np.zeros(256, np.uint32)
    if prev is not None:
        prev = np.ascontiguousarray(prev)
    fn = lib().vto_sad_hist_fast if fast else lib().vto_sad_hist
    fn(cur.ctypes.data, cur.strides[0], prev.ctypes.data if prev is not None else None,
       prev.strides[0] if prev is not None else 0, w, h, sad.ctypes.data, hist.ctypes.data)
    return int(sad[0]), hist


def nv12_to_yuv420p(nv12: np.ndarray, w: int, h: int, pitch: int):
    """nv12: flat u8 buffer, Y plane h*pitch followed by UV plane ceil(h/2)*pitch."""
    nv12 = np.ascontiguousarray(nv12)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = np.zeros((h, w), np.uint8)
    u = np.zeros((ch, cw), np.uint8)
    v = np.zeros((ch, cw), np.uint8)
    base = nv12.ctypes.data
    lib().vto_nv12_to_yuv420p(base, base + h * pitch, pitch, w, h, y.ctypes.data, u.ctypes.data, v.ctypes.data)
    return y, u, v


def pcm_picture_to_yuv420p(buf: np.ndarray, payload_off: int, w: int, h: int):
    """One I_PCM picture whose macroblock 0 starts at buf[payload_off] -> (Y, U, V)."""
    mb_w, mb_h = (w + 15) // 16, (h + 15) // 16
    y = np.empty((h, w), np.uint8)
    u = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    v = np.empty_like(u)
    lib().vto_pcm_picture_to_yuv420p(buf.ctypes.data + int(payload_off), mb_w, mb_h, w, h, y.ctypes.data,
                                     u.ctypes.data, v.ctypes.data)
    return y, u, v


def nv12_to_rgb24(nv12: np.ndarray, w: int, h: int, pitch: int, dw: int | None = None, dh: int | None = None) -> np.ndarray:
    """NV12 surface (flat u8, Y rows then UV rows of `pitch` bytes) -> RGB24 (dh, dw, 3), swscale bicubic semantics."""
    nv12 = np.ascontiguousarray(nv12)
    dw, dh = dw or w, dh or h
    out = np.zeros((dh, dw, 3), np.uint8)
    base = nv12.ctypes.data
    rc = lib().vto_yuv_to_rgb24(base, pitch, base + h * pitch, base + h * pitch + 1, pitch, 2, w, h, out.ctypes.data, dw, dh)
    if rc:
        raise RuntimeError("vto_yuv_to_rgb24 failed")
    return out


def jpeg_encode(y: np.ndarray, u: np.ndarray | None = None, v: np.ndarray | None = None, quality: int = 75,
                restart_interval: int = 0, expand_range: bool = False) -> bytes:
    """Baseline JPEG of a grey (y only) or YCbCr 4:2:0 picture (jpeg_oracle.c)."""
    y = np.ascontiguousarray(y)
    h, w = y.shape
    cap = 4 * w * h + 4096
    out = np.zeros(cap, np.uint8)
    if u is not None:
        u, v = np.ascontiguousarray(u), np.ascontiguousarray(v)
        n = lib().vtj_encode(y.ctypes.data, y.strides[0], u.ctypes.data, v.ctypes.data, u.strides[0], w, h, quality,
                             restart_interval, int(expand_range), out.ctypes.data, cap)
    else:
        n = lib().vtj_encode(y.ctypes.data, y.strides[0], None, None, 0, w, h, quality, restart_interval,
                             int(expand_range), out.ctypes.data, cap)
    if n == 0:
        raise RuntimeError("vtj_encode overflow")
    return out[:n].tobytes()


def jpeg_block(plane: np.ndarray, bx: int, by: int, quality: int, chroma: bool, expand: int = 0) -> np.ndarray:
    """Quantised coefficients (zigzag order) of one 8x8 block."""
    plane = np.ascontiguousarray(plane)
    h, w = plane.shape
    q = np.zeros(64, np.uint8)
    lib().vtj_quant_table(quality, int(chroma), q.ctypes.data)
    zz = np.zeros(64, np.int16)
    lib().vtj_block_coefficients(plane.ctypes.data, plane.strides[0], w, h, bx, by, q.ctypes.data, expand, zz.ctypes.data)
    return zz

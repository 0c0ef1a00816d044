"""Oracle (F): ctypes binding to the libswscale the image carries (OpenCV wheel, FFmpeg 8.0.1, libswscale 9.1.100).

TEST INFRASTRUCTURE ONLY.  The reference shells out to `ffmpeg -vf scale=-2:360`
(/root/reference/src/analyzer/content_analyzer.py:193-211), i.e. libswscale with the scale filter's
default flags (SWS_BICUBIC).  There is no ffmpeg binary in the image, so this module calls the same
library in-process.  It is used to (1) pin oracle/vt_oracle.c (our C restatement of the swscale
arithmetic), (2) generate tests/golden/sws_*.npz.  Nothing in video_transformer_b200/ imports it.
"""
from __future__ import annotations

import ctypes
import glob
import os

import numpy as np

SWS_FAST_BILINEAR = 1
SWS_BILINEAR = 2
SWS_BICUBIC = 4
SWS_POINT = 0x10
SWS_AREA = 0x20
SWS_FULL_CHR_H_INT = 0x2000
SWS_FULL_CHR_H_INP = 0x4000
SWS_ACCURATE_RND = 0x40000
SWS_BITEXACT = 0x80000

PIX_YUV420P = 0
PIX_RGB24 = 2
PIX_GRAY8 = 8
PIX_NV12 = 23

_libs = None


def available() -> bool:
    try:
        _load()
        return True
    except Exception:  # noqa: BLE001
        return False


def _load():
    global _libs
    if _libs is not None:
        return _libs
    import cv2  # the wheel that vendors FFmpeg

    d = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
    avu = ctypes.CDLL(glob.glob(os.path.join(d, "libavutil*"))[0], mode=ctypes.RTLD_GLOBAL)
    sws = ctypes.CDLL(glob.glob(os.path.join(d, "libswscale*"))[0])
    sws.sws_getContext.restype = ctypes.c_void_p
    sws.sws_getContext.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    sws.sws_scale.restype = ctypes.c_int
    sws.sws_scale.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p, ctypes.c_void_p]
    sws.sws_freeContext.argtypes = [ctypes.c_void_p]
    sws.swscale_version.restype = ctypes.c_uint
    _libs = (sws, avu)
    return _libs


def version() -> str:
    sws, _ = _load()
    v = sws.swscale_version()
    return "%d.%d.%d" % (v >> 16, (v >> 8) & 255, v & 255)


def _scale(src_planes, src_fmt, sw, sh, dst_shapes, dst_fmt, dw, dh, flags):
    sws, _ = _load()
    ctx = sws.sws_getContext(sw, sh, src_fmt, dw, dh, dst_fmt, flags, None, None, None)
    if not ctx:
        raise RuntimeError("sws_getContext failed")
    try:
        srcs = [np.ascontiguousarray(p) for p in src_planes]
        # swscale SIMD may read a few bytes past a row end: give every plane slack.
        dsts = [np.zeros((h, w + 64), np.uint8) for (h, w) in dst_shapes]
        sp = (ctypes.c_void_p * 4)(*[s.ctypes.data for s in srcs] + [None] * (4 - len(srcs)))
        ss = (ctypes.c_int * 4)(*[s.strides[0] for s in srcs] + [0] * (4 - len(srcs)))
        dp = (ctypes.c_void_p * 4)(*[d.ctypes.data for d in dsts] + [None] * (4 - len(dsts)))
        ds = (ctypes.c_int * 4)(*[d.strides[0] for d in dsts] + [0] * (4 - len(dsts)))
        rc = sws.sws_scale(ctx, sp, ss, 0, sh, dp, ds)
        if rc != dh:
            raise RuntimeError("sws_scale returned %d, expected %d" % (rc, dh))
        return [np.ascontiguousarray(d[:, :w]) for d, (h, w) in zip(dsts, dst_shapes)]
    finally:
        sws.sws_freeContext(ctx)


def _pad(p):
    """Copy a plane into a buffer with 64 B of row slack (SIMD over-read safety)."""
    h, w = p.shape
    buf = np.zeros((h + 1, w + 64), np.uint8)
    buf[:h, :w] = p
    return buf[:h]


def scale_yuv420p(y, u, v, dw, dh, flags=SWS_BICUBIC):
    sh, sw = y.shape
    cw, ch = -(-dw // 2), -(-dh // 2)
    return _scale([_pad(y), _pad(u), _pad(v)], PIX_YUV420P, sw, sh, [(dh, dw), (ch, cw), (ch, cw)],
                  PIX_YUV420P, dw, dh, flags)


def scale_gray(y, dw, dh, flags=SWS_BICUBIC):
    sh, sw = y.shape
    return _scale([_pad(y)], PIX_GRAY8, sw, sh, [(dh, dw)], PIX_GRAY8, dw, dh, flags)[0]


def yuv420p_to_rgb24(y, u, v, dw=None, dh=None, flags=SWS_BICUBIC):
    sh, sw = y.shape
    dw = dw or sw
    dh = dh or sh
    out = _scale([_pad(y), _pad(u), _pad(v)], PIX_YUV420P, sw, sh, [(dh, dw * 3)], PIX_RGB24, dw, dh, flags)[0]
    return out.reshape(dh, dw, 3)


def nv12_to_rgb24(y, uv, dw=None, dh=None, flags=SWS_BICUBIC):
    sh, sw = y.shape
    dw = dw or sw
    dh = dh or sh
    out = _scale([_pad(y), _pad(uv)], PIX_NV12, sw, sh, [(dh, dw * 3)], PIX_RGB24, dw, dh, flags)[0]
    return out.reshape(dh, dw, 3)

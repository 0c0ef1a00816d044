"""ctypes binding of libvtseg.so (C ABI: include/vtseg.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VT_LIB") or os.path.join(_HERE, "libvtseg.so")   # VT_LIB: measurement builds only

VT_OK = 0
VT_ERR_INVALID, VT_ERR_CUDA, VT_ERR_UNSUPPORTED, VT_ERR_BITSTREAM, VT_ERR_NOMEM, VT_ERR_NVDEC = -1, -2, -3, -4, -5, -6
SWS_BILINEAR, SWS_BICUBIC, SWS_AREA = 2, 4, 0x20


class VtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libvtseg error %d: %s" % (code, msg))
        self.code = code


class StreamInfo(ctypes.Structure):
    _fields_ = [("codec", c_int), ("width", c_int), ("height", c_int), ("coded_width", c_int),
                ("coded_height", c_int), ("fps_num", c_int), ("fps_den", c_int), ("n_frames", c_int),
                ("n_idr", c_int), ("pcm_intra_only", c_int)]


# name -> (restype, argtypes); this table is also what tests use to check the exported symbols.
SIGNATURES = {
    "vt_version": (c_int, []),
    "vt_last_error": (c_char_p, []),
    "vt_launch_count": (c_uint64, []),
    "vt_sws_max_taps": (c_int, [c_int, c_int, c_int]),
    "vt_sws_make_filter": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_int)]),
    "vt_scale_width_for_height": (c_int, [c_int, c_int, c_int]),
    "vt_scale_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "vt_scale_plan_destroy": (None, [c_void_p]),
    "vt_scale_plan_stream_info": (c_int, [c_void_p, c_int, c_void_p]),
    "vt_scale_plane_u8": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "vt_scale_nv12_to_yuv420p": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_size_t, c_int, c_void_p]),
    "vt_scale_plan_fuses_score": (c_int, [c_void_p]),
    "vt_scale_score_nv12_to_yuv420p": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_size_t, c_int,
                                               c_void_p, c_void_p, c_void_p]),
    "vt_nv12_to_yuv420p": (c_int, [c_void_p, c_int, c_size_t, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "vt_nv12_to_rgb24": (c_int, [c_void_p, c_int, c_size_t, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "vt_rgb_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "vt_rgb_plan_destroy": (None, [c_void_p]),
    "vt_scale_nv12_to_rgb24": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_size_t, c_int, c_void_p]),
    "vt_sad_hist_u8": (c_int, [c_void_p, c_int, c_size_t, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vt_gather_frames": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_int, c_void_p, c_void_p]),
    "vt_jpeg_plan_create": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "vt_jpeg_plan_destroy": (None, [c_void_p]),
    "vt_jpeg_max_frame_bytes": (c_size_t, [c_void_p]),
    "vt_jpeg_header": (c_int, [c_void_p, c_void_p, c_size_t, POINTER(c_size_t)]),
    "vt_jpeg_encode_yuv420p": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                       c_void_p]),
    "vt_host_register": (c_int, [c_void_p, c_size_t]),
    "vt_host_unregister": (c_int, [c_void_p]),
    "vt_host_register_source": (c_int, [c_void_p, c_size_t]),
    "vt_copy_to_device_async": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "vt_copy_to_host_async": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "vt_h264_scan": (c_int, [c_void_p, c_size_t, POINTER(StreamInfo), c_void_p, c_void_p, c_void_p, c_int]),
    "vt_h264_pcm_layout": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "vt_h264_pcm_layout_ps": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p,
                                      c_int, c_void_p]),
    "vt_h264_pcm_decode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_size_t,
                                   c_void_p]),
    "vt_ingest_batch_pcm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_size_t,
                                    c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vt_nvdec_probe": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "vt_decode_open": (c_int, [c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "vt_decode_feed": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.c_int64, c_int]),
    "vt_decode_next_surface": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                       POINTER(c_int), POINTER(ctypes.c_int64)]),
    "vt_decode_release_surface": (c_int, [c_void_p, c_uint64]),
    "vt_decode_close": (None, [c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libvtseg.so.  Raises (never falls back) when it is absent or lacks a declared symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VtError(VT_ERR_INVALID, "%s not built: run `python -m video_transformer_b200.build` "
                          "(or __graft_entry__.build())" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export what vtseg.h declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise VtError(rc, (lib().vt_last_error() or b"").decode("utf-8", "replace"))
    return rc

"""K0: the synthetic H.264 writer against libavcodec (CPU), and the CUDA PCM-intra decoder against both."""
import ctypes

import numpy as np
import pytest

from video_transformer_b200 import _lib, synth


def _scan(bs: bytes, cap: int = 4096):
    L = _lib.lib()
    info = _lib.StreamInfo()
    buf = np.frombuffer(bs, np.uint8)
    offs = np.zeros(cap, np.uint64)
    sizes = np.zeros(cap, np.uint32)
    flags = np.zeros(cap, np.uint32)
    _lib.check(L.vt_h264_scan(buf.ctypes.data, len(bs), ctypes.byref(info), offs.ctypes.data, sizes.ctypes.data,
                              flags.ctypes.data, cap))
    n = info.n_frames
    return info, offs[:n], sizes[:n], flags[:n]


def _expected_frames(w, h, meta):
    scene, ref, out = 0, None, []
    cuts = set(meta["cuts"])
    for k in range(meta["n_frames"]):
        if k in cuts:
            scene += 1
        if k in meta["idr_frames"]:
            y, u, v = synth.testsrc_frame(w, h, k, scene)
            ref = (np.maximum(y, 1), np.maximum(u, 1), np.maximum(v, 1))
        out.append(ref)
    return out


@pytest.mark.parametrize("w,h", [(320, 240), (1920, 1080), (854, 480)])
def test_writer_decodes_bit_exact_with_libavcodec(tmp_path, w, h):
    cv2 = pytest.importorskip("cv2")
    bs, meta = synth.make_testsrc_h264(w, h, 9, fps=30, gop=4, cuts=[6])
    p = tmp_path / "clip.h264"
    p.write_bytes(bs)
    cap = cv2.VideoCapture(str(p), cv2.CAP_FFMPEG)
    assert cap.isOpened()
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    exp = _expected_frames(w, h, meta)
    k = 0
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert np.array_equal(np.asarray(fr).reshape(-1)[: w * h].reshape(h, w), exp[k][0]), k
        k += 1
    assert k == 9


def test_scan_reports_stream_facts(vtlib):
    bs, meta = synth.make_testsrc_h264(1920, 1080, 10, fps=30, gop=5, cuts=[7])
    info, offs, sizes, flags = _scan(bs)
    assert (info.width, info.height, info.coded_width, info.coded_height) == (1920, 1080, 1920, 1088)
    assert info.fps_num * 1 == 30 * info.fps_den
    assert info.n_frames == 10 and info.n_idr == 3 and info.pcm_intra_only == 1
    assert [int(f) & 1 for f in flags] == [1 if k in meta["idr_frames"] else 0 for k in range(10)]


def test_scan_rejects_garbage(vtlib):
    L = _lib.lib()
    info = _lib.StreamInfo()
    junk = np.arange(1, 200, dtype=np.uint8)
    assert L.vt_h264_scan(junk.ctypes.data, junk.size, ctypes.byref(info), None, None, None, 0) == _lib.VT_ERR_BITSTREAM


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,pitch", [(320, 240, 320), (1920, 1080, 2048), (854, 480, 1024), (1280, 720, 1280)])
def test_cuda_pcm_decode_bit_exact(cuda, w, h, pitch):
    import torch
    from video_transformer_b200 import decode
    bs, meta = synth.make_testsrc_h264(w, h, 9, fps=30, gop=4, cuts=[6])
    dec = decode.H264PcmDecoder(bs)
    surf = dec.decode(0, 9, pitch=pitch)            # (9, rows, pitch) uint8 CUDA
    got = surf.cpu().numpy()
    exp = _expected_frames(w, h, meta)
    ch = h // 2
    for k in range(9):
        y, u, v = exp[k]
        assert np.array_equal(got[k, :h, :w], y), k
        assert np.array_equal(got[k, h:h + ch, 0:w:2], u), k
        assert np.array_equal(got[k, h:h + ch, 1:w:2], v), k
    # a batch that starts on a skip picture needs the carried-over surface
    tail = dec.decode(5, 4, pitch=pitch, prev=surf[4]).cpu().numpy()
    assert np.array_equal(tail[:, :, :w], got[5:9, :, :w])   # pitch padding is never written

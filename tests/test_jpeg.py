"""Baseline-JPEG encoder of the upload reducer (rows a8 / f4).

Oracle chain: oracle/jpeg_oracle.c restates the Independent JPEG Group's algorithm (integer islow DCT, quantiser rounding,
quality scaling, Annex-K Huffman coding).  It is PINNED against the real library: for grey pictures its entropy-coded
scan, its DQT and its DHT segments are byte-identical to cv2.imencode's (libjpeg-turbo) at every quality tried, with and
without restart intervals.  The CUDA encoder must then equal the oracle byte for byte on 4:2:0 colour pictures, and
libjpeg must decode its output to the source picture within the quantiser's error.
"""
import numpy as np
import pytest

from video_transformer_b200 import synth

cv2 = pytest.importorskip("cv2")


def _segments(b: bytes):
    i, out = 2, []
    while True:
        assert b[i] == 0xFF, i
        m, ln = b[i + 1], (b[i + 2] << 8) | b[i + 3]
        out.append((m, b[i + 4:i + 2 + ln]))
        i += 2 + ln
        if m == 0xDA:
            return out, b[i:]


def _grey(w, h, kind, seed=0):
    if kind == "noise":
        return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    y = synth.testsrc_frame((w + 1) // 2 * 2 + 2, (h + 1) // 2 * 2 + 2, 5 + seed, 1 + seed)[0]
    return np.ascontiguousarray(y[:h, :w])


@pytest.mark.parametrize("w,h", [(64, 48), (200, 120), (333, 77), (640, 360), (8, 8), (17, 9)])
@pytest.mark.parametrize("kind", ["testsrc", "noise"])
def test_oracle_scan_is_byte_identical_to_libjpeg(oracle_c, w, h, kind):
    img = _grey(w, h, kind)
    for q in (5, 25, 50, 75, 90, 100):
        for ri in (0, (w + 7) // 8):
            params = [cv2.IMWRITE_JPEG_QUALITY, q] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, ri] if ri else [])
            ok, ref = cv2.imencode(".jpg", img, params)
            assert ok
            rs, rscan = _segments(ref.tobytes())
            ms, mscan = _segments(oracle_c.jpeg_encode(img, quality=q, restart_interval=ri))
            assert mscan == rscan, (q, ri)
            rd, md = dict(rs), dict(ms)
            assert md[0xDB] == rd[0xDB] and md[0xC0] == rd[0xC0]
            assert b"".join(p for m, p in ms if m == 0xC4) == b"".join(p for m, p in rs if m == 0xC4)


def test_oracle_colour_tables_equal_libjpeg(oracle_c):
    """The chrominance quantiser and Huffman tables: compare with a colour JPEG written by the library."""
    img = np.random.default_rng(3).integers(0, 256, (32, 32, 3), dtype=np.uint8)
    for q in (30, 75, 95):
        ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        rs, _ = _segments(ref.tobytes())
        y = img[:, :, 0]
        ms, _ = _segments(oracle_c.jpeg_encode(y, y[::2, ::2].copy(), y[::2, ::2].copy(), quality=q))
        assert [p for m, p in ms if m == 0xDB] == [p for m, p in rs if m == 0xDB]
        assert sorted(p for m, p in ms if m == 0xC4) == sorted(p for m, p in rs if m == 0xC4)


def _expanded(y, u, v):
    yy = np.clip(((y.astype(np.int64) - 16) * 19077 + 8192) >> 14, 0, 255)
    uu = np.clip((((u.astype(np.int64) - 128) * 18652 + 8192) >> 14) + 128, 0, 255)
    vv = np.clip((((v.astype(np.int64) - 128) * 18652 + 8192) >> 14) + 128, 0, 255)
    return yy, uu, vv


def test_oracle_colour_picture_decodes_to_the_source(oracle_c):
    """libjpeg decodes the oracle's 4:2:0 stream; luma comes back within the quantiser's error of the (range-expanded)
    source."""
    w, h = 640, 360
    yy, xx = np.mgrid[0:h, 0:w]
    y = (60 + (xx * 140) // (w - 1)).astype(np.uint8)                   # smooth, and inside the RGB gamut (the check goes
    u = np.full((h // 2, w // 2), 120, np.uint8)                        # through libjpeg's YCbCr -> BGR and back)
    v = (128 + 12 * np.sin(yy[::2, ::2] / 40.0)).astype(np.uint8)
    jpg = oracle_c.jpeg_encode(y, u, v, quality=90, restart_interval=(w + 15) // 16, expand_range=True)
    img = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
    assert img is not None and img.shape == (h, w, 3)
    got = cv2.cvtColor(img, cv2.COLOR_BGR2YCrCb)
    ey, eu, ev = _expanded(y, u, v)
    assert np.abs(got[:, :, 0].astype(int) - ey).max() <= 4
    assert np.abs(got[::2, ::2, 2].astype(int) - eu).max() <= 4 and np.abs(got[::2, ::2, 1].astype(int) - ev).max() <= 4


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n", [(640, 360, 5), (320, 240, 3), (1280, 720, 2), (322, 182, 3), (66, 34, 4), (16, 16, 2)])
def test_cuda_encoder_equals_the_oracle_byte_for_byte(cuda, oracle_c, w, h, n):
    import torch
    from video_transformer_b200 import ops
    cw, ch = (w + 1) // 2, (h + 1) // 2
    rng = np.random.default_rng(w * 7 + h)
    pics = []
    for k in range(n):
        if k % 3 == 2:                                                   # mid-amplitude noise: long codes, many 0xFF bytes
            y = rng.integers(60, 200, (h, w), dtype=np.uint8)
            u = rng.integers(100, 160, (ch, cw), dtype=np.uint8)
            v = rng.integers(100, 160, (ch, cw), dtype=np.uint8)
        else:
            sy, su, sv = synth.testsrc_frame((w + 1) // 2 * 2, (h + 1) // 2 * 2, 3 + k, k)
            y, u, v = sy[:h, :w].copy(), su[:ch, :cw].copy(), sv[:ch, :cw].copy()
        pics.append((y, u, v))
    flat = np.stack([np.concatenate([p.reshape(-1) for p in pic]) for pic in pics])
    dev = torch.from_numpy(flat).cuda()
    for q, expand in ((75, True), (30, False), (92, True)):
        plan = ops.JpegPlan(w, h, q, expand)
        got = plan.encode_to_host(dev)
        for k, (y, u, v) in enumerate(pics):
            exp = oracle_c.jpeg_encode(y, u, v, quality=q, restart_interval=(w + 15) // 16, expand_range=expand)
            assert got[k] == exp, (q, expand, k, len(got[k]), len(exp))
            img = cv2.imdecode(np.frombuffer(got[k], np.uint8), cv2.IMREAD_COLOR)
            assert img is not None and img.shape == (h, w, 3)
        plan.close()


@pytest.mark.gpu
def test_cuda_encoder_refuses_rows_that_do_not_compress(cuda):
    """Quality 100 on full-range noise needs more bits than the raw samples: the encoder reports it (status 1) instead
    of writing a damaged stream."""
    import torch
    from video_transformer_b200 import _lib, ops
    w, h = 64, 32
    flat = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (2, w * h * 3 // 2), dtype=np.uint8)).cuda()
    plan = ops.JpegPlan(w, h, 100, False)
    with pytest.raises(_lib.VtError):
        plan.encode_to_host(flat)
    plan.close()

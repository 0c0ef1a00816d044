"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for integer/byte work (SAD, histogram, NV12 re-layout, scaler on the bit-exact swscale path).
"""
import numpy as np
import pytest
import torch

from video_transformer_b200 import ops

pytestmark = pytest.mark.gpu


def _nv12_batch(rng, n, w, h, pitch, lo=0, hi=256):
    rows = h + (h + 1) // 2
    buf = rng.integers(lo, hi, (n, rows, pitch), dtype=np.uint8)
    return buf


@pytest.mark.parametrize("w,h,pitch,n", [(1280, 720, 1280, 3), (1920, 1080, 2048, 2), (3840, 2160, 3840, 2),
                                         (1000, 562, 1024, 3), (333, 77, 397, 4), (16, 16, 16, 2),
                                         # widths whose rows share leftover passes (k rows merged), heights that cut a group
                                         (1280, 723, 1280, 2), (1920, 1083, 1920, 2), (768, 37, 768, 3),
                                         (528, 61, 560, 2), (3840, 53, 3840, 2), (1040, 7, 1040, 2),
                                         (854, 480, 896, 3), (1366, 49, 1376, 2), (2001, 19, 2016, 2),   # merged AND ragged
                                         (2560, 1440, 2560, 2), (1024, 47, 1024, 2), (512, 100, 512, 3),  # multiples of 512
                                         (8640, 24, 8704, 2),      # rows wider than 8160: counters flush mid-row
                                         (7680, 4320, 7680, 1)])   # 8K: the largest picture NVDEC would hand over
def test_sad_hist_matches_oracle(cuda, oracle_c, w, h, pitch, n):
    rng = np.random.default_rng(w * 7 + h)
    luma = rng.integers(0, 256, (n, h, pitch), dtype=np.uint8)
    luma[0, : h // 2] = 77            # flat area: worst case for atomics-based histograms
    prev0 = rng.integers(0, 256, (h, pitch), dtype=np.uint8)
    d = torch.from_numpy(luma).to(cuda)
    sad, hist = ops.sad_hist(d.view(-1), w, h, pitch, h * pitch, n, prev0=torch.from_numpy(prev0).to(cuda))
    sad = sad.cpu().numpy().astype(np.uint64)
    hist = hist.cpu().numpy().astype(np.uint32)
    for f in range(n):
        prev = prev0 if f == 0 else luma[f - 1]
        s, hh = oracle_c.sad_hist(luma[f, :, :w], prev[:, :w])
        assert int(sad[f]) == s, (f, int(sad[f]), s)
        assert np.array_equal(hist[f], hh), f
        assert int(hist[f].sum()) == w * h


def test_sad_first_frame_without_prev_is_zero(cuda, oracle_c):
    rng = np.random.default_rng(5)
    luma = rng.integers(0, 256, (2, 720, 1280), dtype=np.uint8)
    sad, hist = ops.sad_hist(torch.from_numpy(luma).to(cuda).view(-1), 1280, 720, 1280, 720 * 1280, 2)
    assert int(sad[0]) == 0
    assert int(sad[1]) == oracle_c.sad_hist(luma[1], luma[0])[0]
    assert np.array_equal(hist[0].cpu().numpy().astype(np.uint32), oracle_c.sad_hist(luma[0], None)[1])


@pytest.mark.parametrize("w,h,pitch", [(1280, 720, 1280), (1920, 1080, 2048), (854, 480, 896), (322, 182, 322)])
def test_nv12_to_yuv420p_exact(cuda, oracle_c, w, h, pitch):
    rng = np.random.default_rng(w + h)
    n = 2
    buf = _nv12_batch(rng, n, w, h, pitch)
    out = ops.nv12_to_yuv420p(torch.from_numpy(buf).to(cuda).view(-1), w, h, pitch, n).cpu().numpy()
    cw, ch = (w + 1) // 2, (h + 1) // 2
    for f in range(n):
        y, u, v = oracle_c.nv12_to_yuv420p(buf[f].reshape(-1), w, h, pitch)
        assert np.array_equal(out[f, : w * h].reshape(h, w), y)
        assert np.array_equal(out[f, w * h: w * h + cw * ch].reshape(ch, cw), u)
        assert np.array_equal(out[f, w * h + cw * ch:].reshape(ch, cw), v)


@pytest.mark.parametrize("flags", [ops.SWS_BICUBIC, ops.SWS_BILINEAR, ops.SWS_AREA])
@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (1280, 720, 640, 360), (1280, 720, 768, 768),
                                         (641, 363, 322, 182), (100, 100, 37, 53)])
def test_scale_plane_generic_bit_exact(cuda, oracle_c, flags, sw, sh, dw, dh):
    rng = np.random.default_rng(sw + dh + flags)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    plan = ops.ScalePlan(sw, sh, dw, dh, flags)
    got = plan.scale_plane(torch.from_numpy(src).to(cuda)).cpu().numpy()
    exp = oracle_c.scale_plane(src, dw, dh, flags)
    assert np.array_equal(got, exp), np.abs(got.astype(int) - exp.astype(int)).max()


@pytest.mark.parametrize("sw,sh,pitch,dw,dh,flags", [
    (1920, 1080, 2048, 1280, 720, ops.SWS_BICUBIC),
    (1280, 720, 1280, 640, 360, ops.SWS_BICUBIC),
    (3840, 2160, 3840, 1280, 720, ops.SWS_BICUBIC),
    (1920, 1080, 1920, 1280, 720, ops.SWS_AREA),
    (1920, 1080, 1920, 1280, 720, ops.SWS_BILINEAR),
    (854, 480, 896, 640, 360, ops.SWS_BICUBIC),
    (1280, 720, 1280, 768, 768, ops.SWS_BICUBIC),
    (1920, 1080, 1920, 1000, 562, ops.SWS_BICUBIC),     # ragged right edge: the last strip slides left
    (1920, 1080, 2048, 854, 480, ops.SWS_BICUBIC),      # ratio 2.25: non-periodic phases, 10 horizontal taps
    (1280, 720, 1280, 854, 480, ops.SWS_BICUBIC),
    (3840, 2160, 3840, 1920, 1080, ops.SWS_BICUBIC),    # 1080 output rows: the vertical table spans two launches
    (1920, 1080, 1920, 1280, 720, ops.SWS_BICUBIC | 0),
    (640, 480, 640, 320, 240, ops.SWS_AREA),
    # exact 3:2 / 2:1 / 3:1 with strips that do not tile the width (adjacent-column layout, last strip slid left)
    (1440, 1080, 1536, 960, 720, ops.SWS_BICUBIC),
    (960, 540, 1024, 640, 360, ops.SWS_BICUBIC),
    (1620, 1080, 1664, 1080, 720, ops.SWS_BICUBIC),
    (1920, 1080, 1920, 960, 540, ops.SWS_BICUBIC),
    (1920, 1080, 2048, 640, 360, ops.SWS_BICUBIC),
    (2880, 1620, 3072, 960, 540, ops.SWS_BICUBIC),
    (810, 540, 896, 540, 360, ops.SWS_BICUBIC),         # 3:2 but the width is not a multiple of 8: pair layout
    (960, 540, 1024, 640, 360, ops.SWS_BILINEAR),
    # odd chroma widths (427, 683, 213): the pair kernel with byte stores, the last strip starting on an odd column
    (1920, 1080, 1920, 1366, 768, ops.SWS_BICUBIC),
    (3840, 2160, 3840, 854, 480, ops.SWS_BICUBIC),
    (1280, 720, 1280, 854, 480, ops.SWS_BILINEAR),
    # large ratios: 17-32 taps run on the dp2a two-pass kernels (12 / 16 coefficient pairs, 24 / 32 vertical taps) ...
    (3840, 2160, 3840, 640, 360, ops.SWS_BICUBIC),      # 6:1, the upload reducer's shape for 4K sources
    (1920, 1080, 2048, 320, 180, ops.SWS_BICUBIC),
    (1920, 1080, 1920, 284, 160, ops.SWS_AREA),
    (2560, 1440, 2560, 426, 240, ops.SWS_BILINEAR),
    # ... and more than 32 taps on the general two-pass kernels
    (3840, 2160, 3840, 426, 240, ops.SWS_BICUBIC),
    # planes the pair kernel does not take: narrower than a strip -> fast two-pass kernels
    # (both batched over the pictures)
    (322, 182, 336, 160, 90, ops.SWS_BICUBIC),
    (640, 360, 640, 200, 112, ops.SWS_BILINEAR),
    (1920, 1080, 1920, 426, 240, ops.SWS_BICUBIC),
    (1280, 720, 1280, 1280, 360, ops.SWS_BICUBIC),      # horizontal taps of 1, vertical 2:1
])
def test_scale_nv12_to_yuv420p_bit_exact(cuda, oracle_c, sw, sh, pitch, dw, dh, flags):
    rng = np.random.default_rng(sw * 3 + dw)
    n = 2 if sw * sh > 640 * 360 else 35              # small shapes: more pictures than one two-pass chunk (32)
    buf = _nv12_batch(rng, n, sw, sh, pitch)
    plan = ops.ScalePlan(sw, sh, dw, dh, flags)
    out = plan.scale_nv12(torch.from_numpy(buf).to(cuda).view(-1), pitch, n).cpu().numpy()
    cw, ch = (dw + 1) // 2, (dh + 1) // 2
    for f in range(n):
        y, u, v = oracle_c.nv12_to_yuv420p(buf[f].reshape(-1), sw, sh, pitch)
        ey, eu, ev = oracle_c.scale_yuv420p(y, u, v, dw, dh, flags)
        gy = out[f, : dw * dh].reshape(dh, dw)
        gu = out[f, dw * dh: dw * dh + cw * ch].reshape(ch, cw)
        gv = out[f, dw * dh + cw * ch:].reshape(ch, cw)
        assert np.array_equal(gy, ey), ("Y", np.abs(gy.astype(int) - ey.astype(int)).max())
        assert np.array_equal(gu, eu), ("U", np.abs(gu.astype(int) - eu.astype(int)).max())
        assert np.array_equal(gv, ev), ("V", np.abs(gv.astype(int) - ev.astype(int)).max())


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (1280, 720, 640, 360), (3840, 2160, 1280, 720),
                                         (1920, 1080, 854, 480)])      # the downloader's default height: 427-wide chroma
def test_headline_shapes_take_the_streaming_kernel(cuda, sw, sh, dw, dh):
    plan = ops.ScalePlan(sw, sh, dw, dh, ops.SWS_BICUBIC)
    for chroma in (False, True):
        info = plan.stream_info(chroma)
        assert info["streaming"] == 1, info
        assert info["tile_w"] <= 1024 and info["tile_w"] % 16 == 0


def test_streaming_and_generic_kernels_agree_on_a_batch(cuda, oracle_c):
    # 7 frames, structured content (flat areas, ramps, noise) so that clipping paths are exercised
    from video_transformer_b200 import synth
    sw, sh, pitch, dw, dh = 1920, 1080, 2048, 1280, 720
    n = 7
    rng = np.random.default_rng(11)
    buf = np.zeros((n, sh + sh // 2, pitch), np.uint8)
    for f in range(n):
        y, u, v = synth.testsrc_frame(sw, sh, 31 * f, f)
        if f % 3 == 2:
            y = rng.integers(0, 256, y.shape, dtype=np.uint8)     # full-range noise: over/undershoot clipping
        buf[f] = synth.planar_to_nv12(y, u, v, pitch).reshape(sh + sh // 2, pitch)
    plan = ops.ScalePlan(sw, sh, dw, dh, ops.SWS_BICUBIC)
    d = torch.from_numpy(buf).to(cuda)
    out = plan.scale_nv12(d.view(-1), pitch, n).cpu().numpy()
    for f in range(n):
        gy = plan.scale_plane(d[f, :sh, :sw].contiguous()).cpu().numpy()
        assert np.array_equal(out[f, : dw * dh].reshape(dh, dw), gy), f
        y, u, v = oracle_c.nv12_to_yuv420p(buf[f].reshape(-1), sw, sh, pitch)
        ey, eu, ev = oracle_c.scale_yuv420p(y, u, v, dw, dh, oracle_c.BICUBIC)
        exp = np.concatenate([ey.reshape(-1), eu.reshape(-1), ev.reshape(-1)])
        assert np.array_equal(out[f], exp), f


def test_gather_frames(cuda):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (10, 4096 + 64), dtype=np.uint8)
    idx = np.array([7, 0, 3, 3, 9], np.int32)
    d = torch.from_numpy(src).to(cuda)
    out = ops.gather_frames(d.view(-1), 4096 + 64, 4096, torch.from_numpy(idx).to(cuda), len(idx)).cpu().numpy()
    assert np.array_equal(out, src[idx, :4096])
    out2 = ops.gather_frames(d.view(-1)[4160 * 2:], 4160, 1000, None, 3).cpu().numpy()
    assert np.array_equal(out2, src[2:5, :1000])


@pytest.mark.parametrize("sw,sh,pitch,dw,dh", [(1280, 720, 1280, 768, 768), (1920, 1080, 2048, 768, 768),
                                               (640, 360, 640, 640, 360), (322, 182, 336, 160, 90),
                                               (1920, 1080, 1920, 1280, 720), (854, 480, 896, 768, 768),
                                               (640, 360, 640, 1280, 720),          # upscale: 4-tap banks
                                               (1280, 720, 1280, 700, 394),         # ragged tiles (700 = 10 x 64 + 60)
                                               (1280, 720, 1283, 768, 768),         # pitch not a multiple of 4: general kernels
                                               (3840, 2160, 3840, 768, 768)])       # 20 horizontal taps: general kernels
@pytest.mark.parametrize("path", ["fused", "split"])
def test_nv12_to_rgb24_scaled_bit_exact(cuda, oracle_c, sw, sh, pitch, dw, dh, path, monkeypatch):
    """K1b / config 5: NV12 -> RGB24 with scaling equals the oracle (itself bit-exact against libswscale).  `fused` is
    the default dispatch (the one-pass tile kernel where the plan qualifies), `split` forces the three-launch path."""
    if path == "split":
        monkeypatch.setenv("VT_RGB_KERNEL", "split")
    else:
        monkeypatch.delenv("VT_RGB_KERNEL", raising=False)
    rng = np.random.default_rng(sw + dw)
    n = 3
    buf = _nv12_batch(rng, n, sw, sh, pitch)
    buf[1, : sh // 2] = 235                      # flat bright area next to noise: exercises the clamps
    plan = ops.RgbPlan(sw, sh, dw, dh)
    launches = ops.lib().vt_launch_count()
    out = plan.scale_nv12(torch.from_numpy(buf).to(cuda).view(-1), pitch, n).cpu().numpy()
    launches = ops.lib().vt_launch_count() - launches
    if path == "fused" and pitch % 4 == 0 and sw <= 1920:
        assert launches == 1                     # one pass: no int16 planes in HBM
    for f in range(n):
        exp = oracle_c.nv12_to_rgb24(buf[f].reshape(-1), sw, sh, pitch, dw, dh)
        assert np.array_equal(out[f], exp), (f, np.abs(out[f].astype(int) - exp.astype(int)).max())


def test_nv12_to_rgb24_same_size_matches_committed_libswscale_output(cuda, oracle_c):
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_rgb_vectors.npz"))
    for name in ("a", "b", "c", "e"):
        sw, sh, pitch, dw, dh = (int(v) for v in gold[name + "_dims"])
        src = torch.from_numpy(np.ascontiguousarray(gold[name + "_nv12"])).to(cuda)
        if (sw, sh) == (dw, dh):
            got = ops.nv12_to_rgb24(src.view(-1), sw, sh, pitch, 1).cpu().numpy()[0]
        else:
            got = ops.RgbPlan(sw, sh, dw, dh).scale_nv12(src.view(-1), pitch, 1).cpu().numpy()[0]
        assert np.array_equal(got, gold[name + "_rgb"]), name


def test_scaler_properties_at_full_config2_size(cuda):
    """BASELINE.json configs[1] shape, 64 pictures per launch: properties that need no oracle.
    (1) a flat picture stays flat (every filter row sums to one and the rounding terms cancel);
    (2) a picture's output does not depend on its position in the batch (segment / strip scheduling);
    (3) histogram mass and SAD symmetry on the same batch."""
    sw, sh, pitch, dw, dh = 1920, 1080, 2048, 1280, 720
    n = 64
    g = torch.Generator(device="cpu").manual_seed(7)
    base = torch.randint(0, 256, (sh + sh // 2, pitch), dtype=torch.uint8, generator=g)
    buf = torch.empty((n, sh + sh // 2, pitch), dtype=torch.uint8)
    buf[:] = base
    for f, v in ((3, 0), (17, 255), (40, 77)):
        buf[f] = v
    buf[50] = torch.randint(0, 256, (sh + sh // 2, pitch), dtype=torch.uint8, generator=g)
    d = buf.to(cuda)
    plan = ops.ScalePlan(sw, sh, dw, dh, ops.SWS_BICUBIC)
    assert plan.stream_info(False)["streaming"] == 1
    out = plan.scale_nv12(d.view(-1), pitch, n)
    for f, v in ((3, 0), (17, 255), (40, 77)):
        assert bool((out[f] == v).all()), (f, v)
    same = [f for f in range(n) if f not in (3, 17, 40, 50)]
    ref = out[same[0]]
    for f in same[1:]:
        assert torch.equal(out[f], ref), f
    assert not torch.equal(out[50], ref)
    sad, hist = ops.sad_hist(d.view(-1), sw, sh, pitch, (sh + sh // 2) * pitch, n)
    assert bool((hist.to(torch.int64).sum(dim=1) == sw * sh).all())
    assert int(sad[1]) == 0 and int(sad[0]) == 0                 # identical neighbours
    pair = torch.stack([buf[50], buf[49]]).to(cuda)             # SAD(a, b) == SAD(b, a)
    s_ab, _ = ops.sad_hist(pair.view(-1), sw, sh, pitch, (sh + sh // 2) * pitch, 2)
    assert int(s_ab[1]) == int(sad[50]) and int(sad[51]) == int(sad[50])


def test_bad_arguments_return_error_codes_not_crashes(cuda, vtlib):
    from ctypes import c_void_p
    from video_transformer_b200 import _lib
    t = torch.zeros(4096, dtype=torch.uint8, device=cuda)
    sad = torch.zeros(1, dtype=torch.int64, device=cuda)
    hist = torch.zeros(256, dtype=torch.int32, device=cuda)
    p = c_void_p(t.data_ptr())
    assert vtlib.vt_sad_hist_u8(p, 8, 64, 16, 4, None, 1, c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), None) == _lib.VT_ERR_INVALID
    assert vtlib.vt_sad_hist_u8(p, 16, 64, 16, 4, None, 0, c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), None) == _lib.VT_ERR_INVALID
    assert vtlib.vt_nv12_to_yuv420p(None, 16, 64, 16, 4, p, 96, 1, None) == _lib.VT_ERR_INVALID
    assert vtlib.vt_gather_frames(p, 64, 0, None, 1, p, None) == _lib.VT_ERR_INVALID
    h = c_void_p()
    from ctypes import byref
    assert vtlib.vt_scale_plan_create(1, 1, 1, 1, 4, byref(h)) == _lib.VT_ERR_INVALID
    assert vtlib.vt_rgb_plan_create(64, 48, 33, 24, 4, byref(h)) == _lib.VT_ERR_UNSUPPORTED       # odd output width
    assert b"odd output width" in vtlib.vt_last_error()
    plan = ops.ScalePlan(64, 48, 32, 24)
    assert vtlib.vt_scale_nv12_to_yuv420p(plan._h, p, 32, 64 * 72, p, 32 * 24 * 3 // 2, 1, None) == _lib.VT_ERR_INVALID   # pitch < width


@pytest.mark.parametrize("offset", [4, 2])
def test_scale_into_a_destination_that_is_not_8_byte_aligned(cuda, oracle_c, offset):
    """The adjacent-column layout stores 8 bytes per lane; a destination aligned to 4 (or 2) bytes must take the pair
    layout (or the general kernels) and still be bit-exact."""
    sw, sh, pitch, dw, dh, n = 1920, 1080, 1920, 1280, 720, 2
    rng = np.random.default_rng(99)
    buf = _nv12_batch(rng, n, sw, sh, pitch)
    plan = ops.ScalePlan(sw, sh, dw, dh, ops.SWS_BICUBIC)
    fb = plan.out_frame_bytes
    raw = torch.zeros(n * fb + 64, dtype=torch.uint8, device=cuda)
    out = raw[offset:offset + n * fb].view(n, fb)
    plan.scale_nv12(torch.from_numpy(buf).to(cuda).view(-1), pitch, n, out=out)
    got = out.cpu().numpy()
    for f in range(n):
        y, u, v = oracle_c.nv12_to_yuv420p(buf[f].reshape(-1), sw, sh, pitch)
        ey, eu, ev = oracle_c.scale_yuv420p(y, u, v, dw, dh, oracle_c.BICUBIC)
        assert np.array_equal(got[f], np.concatenate([ey.reshape(-1), eu.reshape(-1), ev.reshape(-1)])), f
    assert int(raw[:offset].sum()) == 0 and int(raw[offset + n * fb:].sum()) == 0    # nothing written outside


@pytest.mark.parametrize("sw,sh,pitch,dw,dh,fused", [(1920, 1080, 2048, 1280, 720, True), (1920, 1080, 1920, 1280, 720, True),
                                                     (3840, 2160, 3840, 2560, 1440, True), (768, 432, 768, 512, 288, True),
                                                     (960, 540, 1024, 640, 360, False),      # 640 is not a whole number of strips
                                                     (1280, 720, 1280, 640, 360, False)])    # 2:1: separate kernels
@pytest.mark.parametrize("opt_in", [True, False])
def test_fused_scale_and_score_equals_the_separate_kernels(cuda, oracle_c, sw, sh, pitch, dw, dh, fused, opt_in, monkeypatch):
    """K3 inside K2's luma pass (opt-in, VT_FUSED_SCORE=1): frames, SADs and histograms equal the two standalone
    kernels bit for bit (and the oracle), with and without a picture preceding the batch; every source pixel is counted
    exactly once.  Without the opt-in the same call runs the kernels one after the other."""
    if opt_in:
        monkeypatch.setenv("VT_FUSED_SCORE", "1")
    else:
        monkeypatch.delenv("VT_FUSED_SCORE", raising=False)
    rng = np.random.default_rng(sw * 3 + dh)
    n = 5
    buf = _nv12_batch(rng, n, sw, sh, pitch)
    buf[2, : sh // 3] = 200                              # a flat region: every lane hits the same counter row
    buf[3] = buf[2]                                      # identical pictures: SAD 0
    prev = _nv12_batch(rng, 1, sw, sh, pitch)[0]
    plan = ops.ScalePlan(sw, sh, dw, dh)
    assert plan.fuses_score == fused
    d = torch.from_numpy(buf).to(cuda)
    dprev = torch.from_numpy(prev).to(cuda)
    rows = sh + sh // 2
    ref_out = plan.scale_nv12(d.view(-1), pitch, n).cpu().numpy()
    for p0 in (None, dprev):
        ref_sad, ref_hist = ops.sad_hist(d.view(-1), sw, sh, pitch, rows * pitch, n, prev0=None if p0 is None else p0.view(-1))
        before = ops.lib().vt_launch_count()
        out, sad, hist = plan.scale_score_nv12(d.view(-1), pitch, n, prev0=None if p0 is None else p0.view(-1))
        launches = ops.lib().vt_launch_count() - before
        assert launches == (2 if fused and opt_in else 3)   # luma(+score) and chroma / score, luma, chroma
        assert np.array_equal(out.cpu().numpy(), ref_out)
        assert np.array_equal(sad.cpu().numpy(), ref_sad.cpu().numpy())
        assert np.array_equal(hist.cpu().numpy(), ref_hist.cpu().numpy())
        assert int(hist.cpu().numpy().view(np.uint32).sum()) == n * sw * sh
    y0 = buf[0, :sh, :sw]
    s, h = oracle_c.sad_hist(y0, prev[:sh, :sw])
    assert int(sad.cpu()[0]) == s and np.array_equal(hist.cpu().numpy()[0].view(np.uint32), h)
    assert int(sad.cpu()[3]) == 0

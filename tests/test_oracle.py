"""Pins the oracle itself: oracle/vt_oracle.c against libswscale (live, when the image has it) and against the
committed libswscale outputs in tests/golden/sws_vectors.npz; SAD/hist/NV12 against numpy."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_vectors.npz"))
MODES = {"bicubic": 4, "bilinear": 2, "area": 0x20}


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
@pytest.mark.parametrize("mode", list(MODES))
def test_c_oracle_equals_committed_libswscale_bitexact_output(oracle_c, name, mode):
    sw, sh, dw, dh = (int(v) for v in GOLD[name + "_dims"])
    y, u, v = GOLD[name + "_src_y"], GOLD[name + "_src_u"], GOLD[name + "_src_v"]
    gy, gu, gv = oracle_c.scale_yuv420p(y, u, v, dw, dh, MODES[mode])
    assert np.array_equal(gy, GOLD["%s_%s_y" % (name, mode)])
    assert np.array_equal(gu, GOLD["%s_%s_u" % (name, mode)])
    assert np.array_equal(gv, GOLD["%s_%s_v" % (name, mode)])
    # distance to ffmpeg's default (SIMD, non-bitexact) path: the +-1 LSB budget of the north star
    d = np.abs(gy.astype(int) - GOLD["%s_%s_default_y" % (name, mode)].astype(int))
    assert d.max() <= 1
    mse = float((d.astype(np.float64) ** 2).mean())
    assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 50.0


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (1280, 720, 640, 360), (3840, 2160, 1280, 720),
                                         (1280, 720, 768, 768), (641, 363, 322, 182), (854, 480, 640, 360)])
@pytest.mark.parametrize("mode", list(MODES))
def test_c_oracle_equals_live_libswscale(oracle_c, sw, sh, dw, dh, mode):
    from oracle import ffsws
    if not ffsws.available():
        pytest.skip("libswscale not present in this image")
    rng = np.random.default_rng(sw + dh)
    y = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    exp = ffsws.scale_gray(y, dw, dh, MODES[mode] | ffsws.SWS_ACCURATE_RND | ffsws.SWS_BITEXACT)
    assert np.array_equal(oracle_c.scale_plane(y, dw, dh, MODES[mode]), exp)


def test_product_filter_banks_equal_oracle(oracle_c, vtlib):
    from video_transformer_b200 import ops
    for s, d in [(1920, 1280), (1080, 720), (3840, 1280), (1280, 640), (1280, 768), (720, 768), (100, 37), (33, 16)]:
        for fl in MODES.values():
            for one in (1 << 14, 1 << 12):
                a, b = ops.make_filter(s, d, fl, one), oracle_c.make_filter(s, d, fl, one)
                assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (s, d, fl, one)
                assert (a[0].astype(np.int64).sum(axis=1) == one).all()      # every row sums to `one`
                assert (a[1] >= 0).all() and (a[1] + a[2] <= max(s, a[2])).all()


def test_scale_width_rule(vtlib):
    from video_transformer_b200 import ops
    assert ops.scale_width_for_height(1920, 1080, 720) == 1280
    assert ops.scale_width_for_height(1920, 1080, 360) == 640
    assert ops.scale_width_for_height(3840, 2160, 720) == 1280
    assert ops.scale_width_for_height(854, 480, 360) == 640       # 640.5 -> rounds half away from zero /2*2
    assert ops.scale_width_for_height(1000, 562, 360) == 640
    assert ops.scale_width_for_height(720, 1280, 360) == 202      # portrait: 202.5 -> 101.25 pairs -> 202


def test_sad_hist_oracle_against_numpy(oracle_c):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    b = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    s, h = oracle_c.sad_hist(a, b)
    assert s == int(np.abs(a.astype(np.int64) - b).sum())
    assert np.array_equal(h, np.bincount(a.reshape(-1), minlength=256).astype(np.uint32))
    assert oracle_c.sad_hist(a, None)[0] == 0


def test_nv12_oracle_against_numpy(oracle_c):
    rng = np.random.default_rng(2)
    w, h, pitch = 66, 34, 80
    buf = rng.integers(0, 256, (h + h // 2, pitch), dtype=np.uint8)
    y, u, v = oracle_c.nv12_to_yuv420p(buf.reshape(-1), w, h, pitch)
    assert np.array_equal(y, buf[:h, :w])
    assert np.array_equal(u, buf[h:, 0:w:2]) and np.array_equal(v, buf[h:, 1:w:2])


def test_scene_oracle_and_product_agree():
    from oracle import scene_oracle
    from video_transformer_b200 import scene
    rng = np.random.default_rng(4)
    sad = rng.integers(0, 1280 * 720 * 60, 500).astype(np.uint64)
    sad[::37] = 1280 * 720 * 200
    a = scene.scene_scores(sad, 1280, 720)
    b = scene_oracle.scene_scores(sad.tolist(), 1280, 720)
    assert [float(x).hex() for x in a] == [float(x).hex() for x in b]
    for thr in (0.0, 0.10, 0.40):
        assert scene.select_cuts(a, thr).tolist() == scene_oracle.select_cuts(b, thr)
    kf = np.arange(0, 18000, 30)
    for (s, e) in [(0.0, 500.0), (460.0, 600.0), (12.3456, 47.001), (599.99, 600.0), (3.0004, 3.0339), (700.0, 800.0)]:
        for sc in (False, True):
            assert scene.frames_for_window(s, e, 18000, 30, 1, kf, sc) == \
                scene_oracle.frames_for_window(s, e, 18000, 30, 1, kf, sc)
            assert scene.frames_for_window(s, e, 17982, 30000, 1001, kf, sc) == \
                scene_oracle.frames_for_window(s, e, 17982, 30000, 1001, kf, sc)


RGB_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_rgb_vectors.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_rgb_oracle_equals_committed_libswscale_output(oracle_c, name):
    """vto_yuv_to_rgb24 (K1b / config 5) is bit-exact against libswscale's nv12 -> rgb24, default and bit-exact flags."""
    sw, sh, pitch, dw, dh = (int(v) for v in RGB_GOLD[name + "_dims"])
    got = oracle_c.nv12_to_rgb24(RGB_GOLD[name + "_nv12"].reshape(-1), sw, sh, pitch, dw, dh)
    assert np.array_equal(got, RGB_GOLD[name + "_rgb"])
    assert np.array_equal(got, RGB_GOLD[name + "_rgb_bitexact"])


@pytest.mark.parametrize("sw,sh,dw,dh", [(1280, 720, 768, 768), (1920, 1080, 768, 768), (320, 240, 320, 240),
                                         (641, 363, 322, 182)])
def test_rgb_oracle_equals_live_libswscale(oracle_c, sw, sh, dw, dh):
    from oracle import ffsws
    if not ffsws.available():
        pytest.skip("libswscale not present in this image")
    rng = np.random.default_rng(sw + dh)
    pitch = (sw + 15) // 16 * 16
    buf = rng.integers(0, 256, (sh + (sh + 1) // 2, pitch), dtype=np.uint8)
    exp = ffsws.nv12_to_rgb24(buf[:sh, :sw], buf[sh:, : 2 * ((sw + 1) // 2)], dw, dh, ffsws.SWS_BICUBIC)
    assert np.array_equal(oracle_c.nv12_to_rgb24(buf.reshape(-1), sw, sh, pitch, dw, dh), exp)

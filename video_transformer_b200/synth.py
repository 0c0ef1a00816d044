"""Synthetic input for the ingest path: a seeded `testsrc`-like picture generator and an H.264 writer.

BASELINE.json quotes the metric on "synthetic ffmpeg testsrc video".  The image has no ffmpeg binary, no
lavfi and no H.264 encoder, so both are restated here:

* `testsrc_frame` -- deterministic YUV420 picture: colour bars, a luma ramp, a seeded noise rectangle that
  changes only at scene cuts, a moving box and a frame-counter strip.
* `H264PcmWriter` -- a from-scratch Annex-B H.264 writer (Constrained Baseline, CAVLC): every IDR picture
  is made of I_PCM macroblocks (raw samples), every other picture is a single mb_skip_run (repeats the
  previous picture).  Any conforming decoder reconstructs these bit-exactly; entropy decoding cost is NOT
  representative of real CABAC/CAVLC content, and every report that uses these streams says so.

Samples are clipped to [1, 255] so that no emulation-prevention byte is ever needed inside slice data
(a 0x00 sample after the "0D 00" macroblock header could form a start code).
"""
from __future__ import annotations


import numpy as np

_BARS_Y = (235, 210, 170, 145, 106, 81, 41, 16)
_BARS_U = (128, 16, 166, 54, 202, 90, 240, 128)
_BARS_V = (128, 146, 16, 34, 222, 240, 110, 128)


def scene_cut_frames(n_frames: int, fps: float, seed: int = 42, mean_gap_s: float = 8.0, min_gap_s: float = 2.0):
    """Ground-truth cut list: sorted frame indices, spacing >= min_gap_s, deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    cuts, t = [], 0.0
    total = n_frames / fps
    while True:
        t += min_gap_s + rng.exponential(max(mean_gap_s - min_gap_s, 0.1))
        if t >= total:
            break
        k = int(round(t * fps))
        if 0 < k < n_frames and (not cuts or k - cuts[-1] >= int(min_gap_s * fps)):
            cuts.append(k)
    return cuts


def testsrc_frame(w: int, h: int, k: int, scene: int = 0):
    """Frame k of scene `scene` as planar (Y[h,w], U[h/2,w/2], V[h/2,w/2]) uint8, limited range."""
    cw, ch = w // 2, h // 2
    x = np.arange(w)
    bar = np.minimum(x * 8 // w, 7)
    bar = (bar + scene) % 8
    y = np.empty((h, w), np.uint8)
    half = h // 2
    y[:half] = np.asarray(_BARS_Y, np.uint8)[bar][None, :]
    ramp = (16 + (x * 219) // max(w - 1, 1)).astype(np.uint8)
    if scene & 1:
        ramp = ramp[::-1]
    y[half:] = ramp[None, :]
    cbar = bar[::2][:cw]
    u = np.empty((ch, cw), np.uint8)
    v = np.empty((ch, cw), np.uint8)
    u[: ch // 2] = np.asarray(_BARS_U, np.uint8)[cbar][None, :]
    v[: ch // 2] = np.asarray(_BARS_V, np.uint8)[cbar][None, :]
    u[ch // 2:] = 128
    v[ch // 2:] = 128
    # seeded noise rectangle: content depends on the scene only
    rng = np.random.default_rng(1234 + scene)
    nh, nw = max(h // 4, 2) & ~1, max(w // 4, 2) & ~1
    y0, x0 = (h // 8) & ~1, (w // 2 + w // 8) & ~1
    x0 = min(x0, w - nw)
    y[y0:y0 + nh, x0:x0 + nw] = rng.integers(16, 236, (nh, nw), dtype=np.uint8)
    u[y0 // 2:(y0 + nh) // 2, x0 // 2:(x0 + nw) // 2] = rng.integers(16, 241, (nh // 2, nw // 2), dtype=np.uint8)
    v[y0 // 2:(y0 + nh) // 2, x0 // 2:(x0 + nw) // 2] = rng.integers(16, 241, (nh // 2, nw // 2), dtype=np.uint8)
    # moving white box and a frame-counter strip (bits of k as 8-px cells)
    bs = min(64, h // 4, w // 4) & ~1
    bx = ((k * 4) % max(w - bs, 1)) & ~1
    by = (h - h // 4) & ~1
    y[by:by + bs, bx:bx + bs] = 235
    for bit in range(min(24, w // 8)):
        y[0:8, bit * 8:bit * 8 + 8] = 235 if (k >> bit) & 1 else 16
    return y, u, v


def planar_to_nv12(y, u, v, pitch: int | None = None) -> np.ndarray:
    """Planar YUV420 -> flat NV12 buffer (Y rows then interleaved UV rows, `pitch` bytes per row)."""
    h, w = y.shape
    pitch = pitch or w
    ch = (h + 1) // 2
    out = np.zeros((h + ch, pitch), np.uint8)
    out[:h, :w] = y
    out[h:h + u.shape[0], 0:2 * u.shape[1]:2] = u
    out[h:h + v.shape[0], 1:2 * v.shape[1]:2] = v
    return out.reshape(-1)


class _BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, n: int, v: int):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def ue(self, v: int):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v: int):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_zero(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def trailing(self):
        self.bits.append(1)
        self.align_zero()

    def tobytes(self) -> bytes:
        assert len(self.bits) % 8 == 0
        b = np.packbits(np.asarray(self.bits, np.uint8))
        return b.tobytes()


def _escape(rbsp: bytes) -> bytes:
    """Insert emulation prevention bytes (only ever needed for the small header NALs here)."""
    out = bytearray()
    zeros = 0
    for c in rbsp:
        if zeros >= 2 and c <= 3:
            out.append(3)
            zeros = 0
        out.append(c)
        zeros = zeros + 1 if c == 0 else 0
    return bytes(out)


_START = b"\x00\x00\x00\x01"


class H264PcmWriter:
    """Annex-B H.264 elementary stream of I_PCM IDR pictures and all-skip P pictures."""

    def __init__(self, width: int, height: int, fps_num: int = 30, fps_den: int = 1):
        if width % 2 or height % 2:
            raise ValueError("even dimensions required")
        self.w, self.h = width, height
        self.mb_w, self.mb_h = (width + 15) // 16, (height + 15) // 16
        self.fps_num, self.fps_den = fps_num, fps_den
        self.frame_num = 0
        self.idr_count = 0
        self.log2_max_frame_num = 8
        self._sps = self._make_sps()
        self._pps = self._make_pps()

    # -- parameter sets ------------------------------------------------------------------------------------
    def _make_sps(self) -> bytes:
        b = _BitWriter()
        b.u(8, 66)            # profile_idc: Baseline
        b.u(8, 0xC0)          # constraint_set0/1 (Constrained Baseline)
        mbs = self.mb_w * self.mb_h
        level = 31 if mbs <= 3600 else (40 if mbs <= 8192 else (51 if mbs <= 36864 else 52))
        b.u(8, level)
        b.ue(0)               # sps id
        b.ue(self.log2_max_frame_num - 4)
        b.ue(2)               # pic_order_cnt_type 2: output order == decode order
        b.ue(1)               # max_num_ref_frames
        b.u(1, 0)             # gaps_in_frame_num_value_allowed_flag
        b.ue(self.mb_w - 1)
        b.ue(self.mb_h - 1)
        b.u(1, 1)             # frame_mbs_only_flag
        b.u(1, 1)             # direct_8x8_inference_flag
        cr, cb = (self.mb_w * 16 - self.w) // 2, (self.mb_h * 16 - self.h) // 2
        if cr or cb:
            b.u(1, 1); b.ue(0); b.ue(cr); b.ue(0); b.ue(cb)
        else:
            b.u(1, 0)
        b.u(1, 1)             # vui_parameters_present_flag
        b.u(1, 0); b.u(1, 0); b.u(1, 0); b.u(1, 0)   # aspect, overscan, video_signal_type, chroma_loc
        b.u(1, 1)             # timing_info_present_flag
        b.u(32, self.fps_den); b.u(32, 2 * self.fps_num); b.u(1, 1)
        b.u(1, 0); b.u(1, 0)  # nal/vcl hrd
        b.u(1, 0)             # pic_struct_present_flag
        b.u(1, 1)             # bitstream_restriction_flag
        b.u(1, 1); b.ue(0); b.ue(0); b.ue(10); b.ue(10)
        b.ue(0)               # max_num_reorder_frames
        b.ue(1)               # max_dec_frame_buffering
        b.trailing()
        return _START + b"\x67" + _escape(b.tobytes())

    def _make_pps(self) -> bytes:
        b = _BitWriter()
        b.ue(0); b.ue(0)
        b.u(1, 0)             # entropy_coding_mode_flag: CAVLC
        b.u(1, 0)
        b.ue(0)               # one slice group
        b.ue(0); b.ue(0)
        b.u(1, 0); b.u(2, 0)
        b.se(0); b.se(0); b.se(0)
        b.u(1, 1)             # deblocking_filter_control_present_flag
        b.u(1, 0); b.u(1, 0)
        b.trailing()
        return _START + b"\x68" + _escape(b.tobytes())

    # -- pictures ------------------------------------------------------------------------------------------
    def _mb_samples(self, y, u, v) -> np.ndarray:
        """(n_mb, 384) uint8 in macroblock raster order, samples clipped to >= 1, padded by edge replication."""
        H, W = self.mb_h * 16, self.mb_w * 16
        if y.shape != (H, W):
            y = np.pad(y, ((0, H - y.shape[0]), (0, W - y.shape[1])), mode="edge")
            u = np.pad(u, ((0, H // 2 - u.shape[0]), (0, W // 2 - u.shape[1])), mode="edge")
            v = np.pad(v, ((0, H // 2 - v.shape[0]), (0, W // 2 - v.shape[1])), mode="edge")
        ym = y.reshape(self.mb_h, 16, self.mb_w, 16).transpose(0, 2, 1, 3).reshape(-1, 256)
        um = u.reshape(self.mb_h, 8, self.mb_w, 8).transpose(0, 2, 1, 3).reshape(-1, 64)
        vm = v.reshape(self.mb_h, 8, self.mb_w, 8).transpose(0, 2, 1, 3).reshape(-1, 64)
        return np.maximum(np.concatenate([ym, um, vm], axis=1), 1)

    def idr(self, y: np.ndarray, u: np.ndarray, v: np.ndarray, with_params: bool = True) -> bytes:
        b = _BitWriter()
        b.ue(0)               # first_mb_in_slice
        b.ue(7)               # slice_type I (all slices)
        b.ue(0)               # pps id
        b.u(self.log2_max_frame_num, 0)
        b.ue(self.idr_count & 0xFFFF)
        b.u(1, 0); b.u(1, 0)  # dec_ref_pic_marking (IDR)
        b.se(0)               # slice_qp_delta
        b.ue(1)               # disable_deblocking_filter_idc
        b.ue(25)              # mb_type I_PCM
        b.align_zero()
        head = b"\x65" + b.tobytes()
        mbs = self._mb_samples(y, u, v)
        body = np.empty((mbs.shape[0], 386), np.uint8)
        body[:, 0] = 0x0D     # ue(25) = 000011010, then 7 alignment zeros
        body[:, 1] = 0x00
        body[:, 2:] = mbs
        self.idr_count += 1
        self.frame_num = 1
        prefix = (self._sps + self._pps if with_params else b"") + _START
        # offset, inside the returned bytes, of macroblock 0's first sample (what vt_h264_pcm_layout reports as the
        # picture's payload): the generator knows it without parsing anything
        self.last_payload_offset = len(prefix) + len(head)
        self.last_nal_offset = len(prefix)
        out = prefix + head + body.reshape(-1)[2:].tobytes() + b"\x80"
        return out

    def skip(self) -> bytes:
        b = _BitWriter()
        b.ue(0); b.ue(5); b.ue(0)
        b.u(self.log2_max_frame_num, self.frame_num % (1 << self.log2_max_frame_num))
        b.u(1, 0)             # num_ref_idx_active_override_flag
        b.u(1, 0)             # ref_pic_list_modification_flag_l0
        b.u(1, 0)             # adaptive_ref_pic_marking_mode_flag
        b.se(0)
        b.ue(1)
        b.ue(self.mb_w * self.mb_h)   # mb_skip_run
        b.trailing()
        self.frame_num += 1
        return _START + b"\x41" + _escape(b.tobytes())


def make_testsrc_h264(w: int, h: int, n_frames: int, fps: int = 30, gop: int = 30, cuts=None, seed: int = 42):
    """Whole clip as bytes + metadata.  A new IDR is written at every GOP start and at every scene cut; the
    picture content changes at cuts (scene index) and at GOP starts (moving box / counter advance)."""
    wr = H264PcmWriter(w, h, fps, 1)
    cuts = scene_cut_frames(n_frames, fps, seed) if cuts is None else list(cuts)
    cutset = set(cuts)
    chunks, scene, idr_frames = [], 0, []
    for k in range(n_frames):
        if k in cutset:
            scene += 1
        if k % gop == 0 or k in cutset:
            chunks.append(wr.idr(*testsrc_frame(w, h, k, scene)))
            idr_frames.append(k)
        else:
            chunks.append(wr.skip())
    return b"".join(chunks), {"width": w, "height": h, "fps": fps, "n_frames": n_frames, "cuts": cuts,
                              "idr_frames": idr_frames, "gop": gop}

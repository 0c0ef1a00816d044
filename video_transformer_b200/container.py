"""Container layer: ISO-BMFF (MP4) mux/demux for H.264 video and raw Annex-B probing.  Host-side byte I/O.

What it replaces in the reference (all inside ffmpeg/ffprobe child processes):
  * ffprobe `format=duration`                         /root/reference/src/utils/video_utils.py:9-27
  * `-ss S -i IN -t D -movflags +faststart -c copy`   /root/reference/src/utils/video_segmenter.py:118-137
    (input seek to the keyframe at or before S, copy samples, moov before mdat)
The writer produces: ftyp, moov (mvhd, one video trak with avc1/avcC, stts, stss, stsc, stsz, co64), mdat.
Samples are AVCC (4-byte length + NAL); parameter sets live in avcC.
"""
from __future__ import annotations

import ctypes
import struct
from dataclasses import dataclass, field
from math import gcd
from pathlib import Path

import numpy as np


@dataclass
class StreamIndex:
    kind: str                      # "mp4" or "h264"
    path: Path
    width: int
    height: int
    fps_num: int
    fps_den: int
    n_frames: int
    duration: float                # seconds, what ffprobe's format=duration would print
    nal_offsets: np.ndarray        # uint64: file offset of each picture's slice NAL header byte
    nal_sizes: np.ndarray          # uint32
    keyframe: np.ndarray           # bool per picture (IDR)
    sps: bytes = b""               # NAL bytes without start code / length
    pps: bytes = b""
    extra: dict = field(default_factory=dict)

    def pts(self, k: int) -> float:
        """Presentation time of picture k: CFR, one multiply then one divide."""
        return k * self.fps_den / self.fps_num


def _box(kind: bytes, payload: bytes) -> bytes:
    return struct.pack(">I4s", 8 + len(payload), kind) + payload


def _full(kind: bytes, version: int, flags: int, payload: bytes) -> bytes:
    return _box(kind, struct.pack(">I", (version << 24) | flags) + payload)


_MATRIX = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)


def write_mp4(path: str | Path, *, sps: bytes, pps: bytes, samples, width: int, height: int, fps_num: int,
              fps_den: int, keyframes) -> None:
    """Write an H.264 video-only MP4 with the moov box first (the `+faststart` layout).

    samples: sequence of NAL byte strings (one slice NAL per picture; no start codes);
    keyframes: iterable of booleans, one per sample.
    """
    sizes = [4 + len(s) for s in samples]
    n = len(sizes)
    g = gcd(fps_num, fps_den) or 1
    timescale, delta = fps_num // g, fps_den // g
    media_dur = n * delta
    movie_dur = (media_dur * 1000 + timescale // 2) // timescale  # mvhd timescale 1000, as ffmpeg writes
    sync = [i + 1 for i, k in enumerate(keyframes) if k]

    avcc = struct.pack(">BBBBBB", 1, sps[1], sps[2], sps[3], 0xFF, 0xE1) + struct.pack(">H", len(sps)) + sps + \
        struct.pack(">BH", 1, len(pps)) + pps
    avc1 = struct.pack(">6xH", 1) + struct.pack(">HHIII", 0, 0, 0, 0, 0) + struct.pack(">HH", width, height) + \
        struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + bytes(32) + struct.pack(">Hh", 0x18, -1) + \
        _box(b"avcC", avcc)
    stsd = _full(b"stsd", 0, 0, struct.pack(">I", 1) + _box(b"avc1", avc1))
    stts = _full(b"stts", 0, 0, struct.pack(">III", 1, n, delta))
    stss = _full(b"stss", 0, 0, struct.pack(">I", len(sync)) + struct.pack(">%dI" % len(sync), *sync))
    stsc = _full(b"stsc", 0, 0, struct.pack(">IIII", 1, 1, 1, 1))
    stsz = _full(b"stsz", 0, 0, struct.pack(">II", 0, n) + np.asarray(sizes, ">u4").tobytes())

    def build_moov(first_sample_offset: int) -> bytes:
        offs = (np.uint64(first_sample_offset) +
                np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)).astype(">u8")
        co64 = _full(b"co64", 0, 0, struct.pack(">I", n) + offs.tobytes())
        stbl = _box(b"stbl", stsd + stts + stss + stsc + stsz + co64)
        dinf = _box(b"dinf", _full(b"dref", 0, 0, struct.pack(">I", 1) + _full(b"url ", 0, 1, b"")))
        minf = _box(b"minf", _full(b"vmhd", 0, 1, bytes(8)) + dinf + stbl)
        mdhd = _full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, media_dur, 0x55C4, 0))
        hdlr = _full(b"hdlr", 0, 0, struct.pack(">I4s12x", 0, b"vide") + b"VideoHandler\x00")
        mdia = _box(b"mdia", mdhd + hdlr + minf)
        tkhd = _full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, movie_dur) + bytes(8) +
                     struct.pack(">hhhH", 0, 0, 0, 0) + _MATRIX + struct.pack(">II", width << 16, height << 16))
        trak = _box(b"trak", tkhd + mdia)
        mvhd = _full(b"mvhd", 0, 0, struct.pack(">IIIIIH", 0, 0, 1000, movie_dur, 0x10000, 0x0100) + bytes(10) +
                     _MATRIX + bytes(24) + struct.pack(">I", 2))
        return _box(b"moov", mvhd + trak)

    ftyp = _box(b"ftyp", b"isom" + struct.pack(">I", 0x200) + b"isomiso2avc1mp41")
    moov_len = len(build_moov(0))
    mdat_size = sum(sizes)
    header = ftyp + build_moov(len(ftyp) + moov_len + 16)
    path = Path(path)
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack(">I4sQ", 1, b"mdat", 16 + mdat_size))
        for s in samples:
            f.write(struct.pack(">I", len(s)))
            f.write(s)


def _iter_boxes(buf, start: int, end: int):
    pos = start
    while pos + 8 <= end:
        size, kind = struct.unpack_from(">I4s", buf, pos)
        head = 8
        if size == 1:
            size = struct.unpack_from(">Q", buf, pos + 8)[0]
            head = 16
        elif size == 0:
            size = end - pos
        if size < head or pos + size > end:
            return
        yield kind, pos + head, pos + size
        pos += size


def _find(buf, start, end, kind):
    for k, s, e in _iter_boxes(buf, start, end):
        if k == kind:
            return s, e
    return None


def _sps_size(sps: bytes) -> tuple[int, int]:
    """Display size from an SPS NAL via the library's parser (vt_h264_scan on a tiny Annex-B blob)."""
    from . import _lib
    blob = np.frombuffer(b"\x00\x00\x00\x01" + sps + b"\x00\x00\x00\x01\x68\xce\x38\x80", np.uint8)
    info = _lib.StreamInfo()
    _lib.check(_lib.lib().vt_h264_scan(blob.ctypes.data, blob.size, ctypes.byref(info), None, None, None, 0))
    return info.width, info.height


def probe_mp4(path: Path) -> StreamIndex | None:
    data = np.memmap(path, dtype=np.uint8, mode="r")
    n_bytes = data.size
    moov = None
    pos = 0
    # walk top-level boxes reading only their headers (mdat can be many GB)
    while pos + 8 <= n_bytes:
        size, kind = struct.unpack(">I4s", bytes(data[pos:pos + 8]))
        head = 8
        if size == 1:
            size = struct.unpack(">Q", bytes(data[pos + 8:pos + 16]))[0]
            head = 16
        elif size == 0:
            size = n_bytes - pos
        if size < head:
            return None
        if kind == b"moov":
            moov = bytes(data[pos + head:pos + size])
            break
        pos += size
    if moov is None:
        return None
    mvhd = _find(moov, 0, len(moov), b"mvhd")
    if mvhd is None:
        return None
    ver = moov[mvhd[0]]
    if ver == 1:
        ts, dur = struct.unpack_from(">IQ", moov, mvhd[0] + 20)
    else:
        ts, dur = struct.unpack_from(">II", moov, mvhd[0] + 12)
    movie_duration = dur / ts if ts else 0.0
    for kind, s, e in _iter_boxes(moov, 0, len(moov)):
        if kind != b"trak":
            continue
        mdia = _find(moov, s, e, b"mdia")
        if mdia is None:
            continue
        hdlr = _find(moov, mdia[0], mdia[1], b"hdlr")
        if hdlr is None or moov[hdlr[0] + 8:hdlr[0] + 12] != b"vide":
            continue
        mdhd = _find(moov, mdia[0], mdia[1], b"mdhd")
        mver = moov[mdhd[0]]
        m_ts = struct.unpack_from(">I", moov, mdhd[0] + (20 if mver == 1 else 12))[0]
        minf = _find(moov, mdia[0], mdia[1], b"minf")
        stbl = _find(moov, minf[0], minf[1], b"stbl")
        stsd = _find(moov, stbl[0], stbl[1], b"stsd")
        entry = stsd[0] + 8
        esize, ekind = struct.unpack_from(">I4s", moov, entry)
        if ekind != b"avc1":
            return None
        width, height = struct.unpack_from(">HH", moov, entry + 8 + 24)
        avcc = _find(moov, entry + 8 + 78, entry + esize, b"avcC")
        a = avcc[0]
        nal_len_size = (moov[a + 4] & 3) + 1
        n_sps = moov[a + 5] & 31
        p = a + 6
        sps = b""
        for _ in range(n_sps):
            ln = struct.unpack_from(">H", moov, p)[0]
            sps = sps or moov[p + 2:p + 2 + ln]
            p += 2 + ln
        n_pps = moov[p]
        p += 1
        pps = b""
        for _ in range(n_pps):
            ln = struct.unpack_from(">H", moov, p)[0]
            pps = pps or moov[p + 2:p + 2 + ln]
            p += 2 + ln
        stts = _find(moov, stbl[0], stbl[1], b"stts")
        n_tt = struct.unpack_from(">I", moov, stts[0] + 4)[0]
        tt = np.frombuffer(moov, ">u4", 2 * n_tt, stts[0] + 8).reshape(-1, 2)
        stsz = _find(moov, stbl[0], stbl[1], b"stsz")
        fixed, n = struct.unpack_from(">II", moov, stsz[0] + 4)
        sizes = np.full(n, fixed, np.uint64) if fixed else np.frombuffer(moov, ">u4", n, stsz[0] + 12).astype(np.uint64)
        stsc = _find(moov, stbl[0], stbl[1], b"stsc")
        n_sc = struct.unpack_from(">I", moov, stsc[0] + 4)[0]
        sc = np.frombuffer(moov, ">u4", 3 * n_sc, stsc[0] + 8).reshape(-1, 3)
        co = _find(moov, stbl[0], stbl[1], b"co64")
        if co is not None:
            n_co = struct.unpack_from(">I", moov, co[0] + 4)[0]
            chunk_off = np.frombuffer(moov, ">u8", n_co, co[0] + 8).astype(np.uint64)
        else:
            co = _find(moov, stbl[0], stbl[1], b"stco")
            n_co = struct.unpack_from(">I", moov, co[0] + 4)[0]
            chunk_off = np.frombuffer(moov, ">u4", n_co, co[0] + 8).astype(np.uint64)
        # samples per chunk -> sample offsets
        per_chunk = np.zeros(n_co, np.int64)
        for i in range(n_sc):
            first = int(sc[i, 0]) - 1
            last = int(sc[i + 1, 0]) - 1 if i + 1 < n_sc else n_co
            per_chunk[first:last] = int(sc[i, 1])
        offs = np.zeros(n, np.uint64)
        k = 0
        for c in range(n_co):
            o = int(chunk_off[c])
            for _ in range(int(per_chunk[c])):
                if k >= n:
                    break
                offs[k] = o
                o += int(sizes[k])
                k += 1
        stss = _find(moov, stbl[0], stbl[1], b"stss")
        key = np.zeros(n, bool)
        if stss is None:
            key[:] = True
        else:
            n_ss = struct.unpack_from(">I", moov, stss[0] + 4)[0]
            key[np.frombuffer(moov, ">u4", n_ss, stss[0] + 8).astype(np.int64) - 1] = True
        if n_tt == 0 or n == 0:
            return None
        delta = int(tt[0, 1])
        g = gcd(int(m_ts), delta) or 1
        try:
            dw, dh = _sps_size(sps)
        except Exception:  # noqa: BLE001
            dw, dh = width, height
        return StreamIndex("mp4", Path(path), dw, dh, int(m_ts) // g, delta // g, n, movie_duration,
                           offs + np.uint64(nal_len_size), (sizes - nal_len_size).astype(np.uint32), key, sps, pps,
                           {"cfr": bool(n_tt == 1), "nal_length_size": nal_len_size})
    return None


def probe_h264(path: Path) -> StreamIndex | None:
    from . import _lib
    data = np.memmap(path, dtype=np.uint8, mode="r")
    L = _lib.lib()
    info = _lib.StreamInfo()
    if L.vt_h264_scan(data.ctypes.data, data.size, ctypes.byref(info), None, None, None, 0) != 0:
        return None
    n = info.n_frames
    offs = np.zeros(max(n, 1), np.uint64)
    sizes = np.zeros(max(n, 1), np.uint32)
    flags = np.zeros(max(n, 1), np.uint32)
    if L.vt_h264_scan(data.ctypes.data, data.size, ctypes.byref(info), offs.ctypes.data, sizes.ctypes.data,
                      flags.ctypes.data, n) != 0 or n == 0 or info.fps_num <= 0:
        return None
    g = gcd(info.fps_num, info.fps_den) or 1
    fn, fd = info.fps_num // g, info.fps_den // g
    return StreamIndex("h264", Path(path), info.width, info.height, fn, fd, n, n * fd / fn, offs[:n], sizes[:n],
                       (flags[:n] & 1).astype(bool), extra={"pcm_intra_only": bool(info.pcm_intra_only)})


def probe(path: Path) -> StreamIndex | None:
    """Index a media file.  Returns None when the file is not a container this layer reads."""
    path = Path(path)
    if not path.is_file() or path.stat().st_size < 16:
        return None
    with open(path, "rb") as f:
        head = f.read(12)
    if head[4:8] in (b"ftyp", b"moov", b"free", b"mdat", b"styp"):
        return probe_mp4(path)
    if head[:4] == b"\x00\x00\x00\x01" or head[:3] == b"\x00\x00\x01":
        return probe_h264(path)
    return None


def annexb_to_mp4(src_h264: str | Path, dst_mp4: str | Path) -> StreamIndex:
    """Wrap a raw Annex-B H.264 file into MP4 (used to build synthetic .mp4 inputs)."""
    idx = probe_h264(Path(src_h264))
    if idx is None:
        raise ValueError("not an Annex-B H.264 stream: %s" % src_h264)
    data = np.memmap(src_h264, dtype=np.uint8, mode="r")
    sps, pps = find_parameter_sets(data)
    samples = [bytes(data[int(o):int(o) + int(s)]) for o, s in zip(idx.nal_offsets, idx.nal_sizes)]
    write_mp4(dst_mp4, sps=sps, pps=pps, samples=samples, width=idx.width, height=idx.height, fps_num=idx.fps_num,
              fps_den=idx.fps_den, keyframes=idx.keyframe)
    return probe_mp4(Path(dst_mp4))


def find_parameter_sets(data: np.ndarray, limit: int = 1 << 16) -> tuple[bytes, bytes]:
    """First SPS and PPS NAL of an Annex-B buffer (searched in its first `limit` bytes)."""
    head = bytes(data[:limit])
    sps = pps = b""
    pos = 0
    while True:
        i = head.find(b"\x00\x00\x01", pos)
        if i < 0:
            break
        j = head.find(b"\x00\x00\x01", i + 3)
        end = j if j >= 0 else len(head)
        nal = head[i + 3:end].rstrip(b"\x00")
        if nal:
            t = nal[0] & 31
            if t == 7 and not sps:
                sps = nal
            elif t == 8 and not pps:
                pps = nal
            elif t in (1, 5):
                break
        pos = i + 3
        if sps and pps:
            break
    if not sps or not pps:
        raise ValueError("no SPS/PPS found")
    return sps, pps

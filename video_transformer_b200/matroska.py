"""Matroska / WebM without ffprobe / ffmpeg: container duration, and a demuxer that indexes the blocks of every track into
the track model of isobmff.py so that Matroska sources are stream-copied into MP4 segments (EBML walk; host-side byte I/O).

Replaces `ffprobe -show_entries format=duration` (/root/reference/src/utils/video_utils.py:7-38) for `.mkv`/`.webm`
inputs: libavformat's matroska demuxer reports Segment/Info/Duration (a float in TimestampScale units, default
1 ms) as the container duration.  When a muxer could not seek back to fill Duration in (live streams), the time of
the last block is used: last Cluster timestamp + the largest block offset inside it.
"""
from __future__ import annotations

import os
import struct
from pathlib import Path

ID_EBML = 0x1A45DFA3
ID_SEGMENT = 0x18538067
ID_INFO = 0x1549A966
ID_TIMESTAMP_SCALE = 0x2AD7B1
ID_DURATION = 0x4489
ID_CLUSTER = 0x1F43B675
ID_CLUSTER_TIMESTAMP = 0xE7
ID_SIMPLE_BLOCK = 0xA3
ID_BLOCK_GROUP = 0xA0
ID_BLOCK = 0xA1
UNKNOWN_SIZE = -1


def _read_id(buf: bytes, pos: int):
    if pos >= len(buf):
        return None
    b0 = buf[pos]
    if b0 == 0:
        return None
    n = 8 - b0.bit_length() + 1
    if n > 4 or pos + n > len(buf):
        return None
    return int.from_bytes(buf[pos:pos + n], "big"), pos + n


def _read_size(buf: bytes, pos: int):
    if pos >= len(buf):
        return None
    b0 = buf[pos]
    if b0 == 0:
        return None
    n = 8 - b0.bit_length() + 1
    if pos + n > len(buf):
        return None
    val = int.from_bytes(buf[pos:pos + n], "big") & ((1 << (7 * n)) - 1)
    if val == (1 << (7 * n)) - 1:
        val = UNKNOWN_SIZE
    return val, pos + n


class _Reader:
    """Element headers read straight from the file (clusters can be many GB; only headers are touched)."""

    def __init__(self, path: Path):
        self.f = open(path, "rb")
        self.size = os.path.getsize(path)

    def header(self, pos: int):
        """(id, payload_start, payload_size) of the element at pos, or None."""
        self.f.seek(pos)
        buf = self.f.read(16)
        r = _read_id(buf, 0)
        if r is None:
            return None
        eid, p = r
        r = _read_size(buf, p)
        if r is None:
            return None
        sz, p = r
        return eid, pos + p, sz

    def read(self, pos: int, n: int) -> bytes:
        self.f.seek(pos)
        return self.f.read(n)

    def close(self):
        self.f.close()


def _children(buf: bytes):
    pos = 0
    while pos < len(buf):
        r = _read_id(buf, pos)
        if r is None:
            return
        eid, p = r
        r = _read_size(buf, p)
        if r is None:
            return
        sz, p = r
        if sz == UNKNOWN_SIZE or p + sz > len(buf):
            return
        yield eid, buf[p:p + sz]
        pos = p + sz


def _last_block_time(rd: _Reader, clusters, scale_ns: int) -> float:
    """Seconds of the last block of the last cluster (fallback when Info has no Duration)."""
    if not clusters:
        return 0.0
    start, size = clusters[-1]
    end = rd.size if size == UNKNOWN_SIZE else min(start + size, rd.size)
    pos = start
    cluster_ts = 0
    best = 0
    while pos < end:
        h = rd.header(pos)
        if h is None:
            break
        eid, p, sz = h
        if sz == UNKNOWN_SIZE:
            break
        if eid == ID_CLUSTER_TIMESTAMP:
            cluster_ts = int.from_bytes(rd.read(p, sz), "big")
        elif eid == ID_SIMPLE_BLOCK or eid == ID_BLOCK_GROUP:
            body = rd.read(p, min(sz, 32))
            if eid == ID_BLOCK_GROUP:
                r = _read_id(body, 0)
                if r and r[0] == ID_BLOCK:
                    r2 = _read_size(body, r[1])
                    body = body[r2[1]:] if r2 else b""
                else:
                    body = b""
            r = _read_size(body, 0)             # track number (vint)
            if r and r[1] + 2 <= len(body):
                rel = struct.unpack_from(">h", body, r[1])[0]
                best = max(best, rel)
        pos = p + sz
    return (cluster_ts + best) * scale_ns / 1e9


def duration_seconds(path: str | Path) -> float:
    """Container duration of a Matroska/WebM file in seconds; 0.0 when the file cannot be read as such."""
    path = Path(path)
    try:
        rd = _Reader(path)
    except OSError:
        return 0.0
    try:
        h = rd.header(0)
        if h is None or h[0] != ID_EBML or h[2] == UNKNOWN_SIZE:
            return 0.0
        pos = h[1] + h[2]
        h = rd.header(pos)
        if h is None or h[0] != ID_SEGMENT:
            return 0.0
        seg_start, seg_size = h[1], h[2]
        seg_end = rd.size if seg_size == UNKNOWN_SIZE else min(seg_start + seg_size, rd.size)
        pos = seg_start
        scale_ns = 1000000
        duration = None
        clusters = []
        while pos < seg_end:
            h = rd.header(pos)
            if h is None:
                break
            eid, p, sz = h
            if eid == ID_INFO and sz != UNKNOWN_SIZE:
                for cid, body in _children(rd.read(p, sz)):
                    if cid == ID_TIMESTAMP_SCALE:
                        scale_ns = int.from_bytes(body, "big") or 1000000
                    elif cid == ID_DURATION:
                        if len(body) == 4:
                            duration = struct.unpack(">f", body)[0]
                        elif len(body) == 8:
                            duration = struct.unpack(">d", body)[0]
                if duration is not None:
                    break
            if eid == ID_CLUSTER:
                clusters.append((p, sz))
                if sz == UNKNOWN_SIZE:
                    break                        # live stream: cannot skip ahead without scanning
            if sz == UNKNOWN_SIZE:
                break
            pos = p + sz
        if duration is not None and duration > 0:
            # libavformat: duration * timestamp_scale (ns) rescaled to microseconds
            return int(duration * scale_ns / 1000.0 + 0.5) / 1e6 if duration * scale_ns >= 1000 else 0.0
        return _last_block_time(rd, clusters, scale_ns)
    except (OSError, struct.error, ValueError):
        return 0.0
    finally:
        rd.close()


# ---- demux: Matroska / WebM -> the same track model the ISO-BMFF cutter works on ---------------------------------------
# The reference's `ffmpeg -ss S -i IN.webm -t D -c copy OUT.mp4` (/root/reference/src/utils/video_segmenter.py:118-137)
# re-wraps Matroska streams into MP4 without decoding.  read_movie() indexes the blocks of every track (offsets, sizes,
# timestamps, keyframe flags) and builds the MP4 sample description each codec needs from its CodecPrivate, so that
# isobmff.cut_movie treats the file like any MP4.  Codecs that have an MP4 mapping here: V_VP9, V_AV1,
# V_MPEG4/ISO/AVC, V_MPEGH/ISO/HEVC, V_MPEG4/ISO/ASP, V_VP8; A_OPUS, A_AAC, A_MPEG/L3, A_FLAC.  Tracks with other codecs (Vorbis ...) have
# no MP4 sample entry and are left out (ffmpeg refuses them in MP4 as well).
ID_TRACKS = 0x1654AE6B
ID_TRACK_ENTRY = 0xAE
ID_TRACK_NUMBER = 0xD7
ID_TRACK_TYPE = 0x83
ID_CODEC_ID = 0x86
ID_CODEC_PRIVATE = 0x63A2
ID_DEFAULT_DURATION = 0x23E383
ID_VIDEO = 0xE0
ID_PIXEL_WIDTH = 0xB0
ID_PIXEL_HEIGHT = 0xBA
ID_AUDIO = 0xE1
ID_SAMPLING_FREQ = 0xB5
ID_CHANNELS = 0x9F
ID_BLOCK_DURATION = 0x9B
ID_REFERENCE_BLOCK = 0xFB


def _uint(b: bytes) -> int:
    return int.from_bytes(b, "big")


def _float(b: bytes) -> float:
    return struct.unpack(">f", b)[0] if len(b) == 4 else struct.unpack(">d", b)[0] if len(b) == 8 else 0.0


def _descr(tag: int, payload: bytes) -> bytes:
    """MPEG-4 descriptor with a one-byte-per-7-bits length (as many bytes as needed)."""
    n = len(payload)
    size = bytes([n & 0x7F])
    n >>= 7
    while n:
        size = bytes([0x80 | (n & 0x7F)]) + size
        n >>= 7
    return bytes([tag]) + size + payload


def _esds(object_type: int, stream_type: int, dsi: bytes) -> bytes:
    from .isobmff import full_box
    dcd = bytes([object_type, (stream_type << 2) | 1]) + bytes(3) + struct.pack(">II", 0, 0) + \
        (_descr(5, dsi) if dsi else b"")         # MP3 has no decoder-specific info
    es = struct.pack(">HB", 0, 0) + _descr(4, dcd) + _descr(6, b"\x02")
    return full_box(b"esds", 0, 0, _descr(3, es))


def _video_entry(codec: bytes, width: int, height: int, extra: bytes) -> bytes:
    from .isobmff import box
    return box(codec, struct.pack(">6xH", 1) + bytes(16) + struct.pack(">HH", width, height) +
               struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + bytes(32) + struct.pack(">Hh", 0x18, -1) + extra)


def _audio_entry(codec: bytes, channels: int, rate: int, extra: bytes) -> bytes:
    from .isobmff import box
    return box(codec, struct.pack(">6xH", 1) + struct.pack(">HHIHHHHI", 0, 0, 0, channels, 16, 0, 0,
                                                           (min(rate, 65535) << 16) & 0xFFFFFFFF) + extra)


def _sample_entry(codec_id: str, private: bytes, width: int, height: int, channels: int, rate: float):
    """(fourcc, sample entry box) of a Matroska track, or None when the codec has no MP4 mapping here."""
    from .isobmff import box, full_box
    if codec_id == "V_VP9":
        # vpcC: profile 0, level unknown (10), 8 bit 4:2:0, limited range, unspecified colour; no init data
        vpcc = full_box(b"vpcC", 1, 0, bytes([0, 10, (8 << 4) | (1 << 1) | 0, 2, 2, 2]) + struct.pack(">H", 0))
        return b"vp09", _video_entry(b"vp09", width, height, vpcc)
    if codec_id == "V_VP8":
        vpcc = full_box(b"vpcC", 1, 0, bytes([0, 10, (8 << 4) | (1 << 1) | 0, 2, 2, 2]) + struct.pack(">H", 0))
        return b"vp08", _video_entry(b"vp08", width, height, vpcc)
    if codec_id == "V_AV1" and private:
        return b"av01", _video_entry(b"av01", width, height, box(b"av1C", private))
    if codec_id == "V_MPEG4/ISO/AVC" and private:
        return b"avc1", _video_entry(b"avc1", width, height, box(b"avcC", private))
    if codec_id == "V_MPEGH/ISO/HEVC" and private:
        return b"hvc1", _video_entry(b"hvc1", width, height, box(b"hvcC", private))
    if codec_id == "V_MPEG4/ISO/ASP":
        return b"mp4v", _video_entry(b"mp4v", width, height, _esds(0x20, 4, private))
    if codec_id == "A_OPUS" and len(private) >= 19 and private[:8] == b"OpusHead":
        ch = private[9]
        pre_skip, in_rate, gain = struct.unpack_from("<HIh", private, 10)
        family = private[18]
        dops = bytes([0, ch]) + struct.pack(">HIh", pre_skip, in_rate, gain) + bytes([family]) + \
            (private[19:19 + 2 + ch] if family else b"")
        return b"Opus", _audio_entry(b"Opus", ch, 48000, box(b"dOps", dops))
    if codec_id == "A_AAC" and private:
        return b"mp4a", _audio_entry(b"mp4a", channels or 2, int(rate or 48000), _esds(0x40, 5, private))
    if codec_id == "A_MPEG/L3":                      # MPEG-1/2 layer III: `mp4a` with object type 0x6B, no decoder config
        return b"mp4a", _audio_entry(b"mp4a", channels or 2, int(rate or 44100), _esds(0x6B, 5, b""))
    if codec_id == "A_FLAC" and private[:4] == b"fLaC" and len(private) >= 4 + 4 + 34:
        # dfLa: the metadata blocks of the CodecPrivate (STREAMINFO first) without the `fLaC` marker
        return b"fLaC", _audio_entry(b"fLaC", channels or 2, int(rate or 48000), full_box(b"dfLa", 0, 0, private[4:]))
    return None


def build_track(track_id: int, video: bool, codec: bytes, entry_box: bytes, timescale: int, width: int, height: int,
                samples: list, default_delta: int = 0):
    """isobmff.Track from a demuxed sample list [(file offset, size, presentation time, keyframe)] in decode order
    (shared by the Matroska and FLV readers): decode times are the sorted presentation times, the difference becomes
    `ctts`, the first decode time an empty edit."""
    import numpy as np
    from . import isobmff
    matrix = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)
    arr = np.asarray([(o, s, ts) for o, s, ts, _k in samples], np.int64)
    pts = arr[:, 2]
    dts = np.sort(pts)                                  # samples are in decode order; sorted pts are valid decode times
    cts_off = pts - dts
    if cts_off.min() < 0:                               # keep offsets non-negative (the edit list removes the shift)
        shift = int(-cts_off.min())
        cts_off = cts_off + shift
    else:
        shift = 0
    deltas = np.diff(dts)
    last = default_delta if default_delta else (int(deltas[-1]) if deltas.size else 1)
    deltas = np.append(deltas, max(last, 1)).astype(np.int64)
    dts0 = int(dts[0])
    tkhd = struct.pack(">I", 3) + struct.pack(">IIIII", 0, 0, track_id, 0, 0) + bytes(8) + \
        struct.pack(">hhhH", 0, 0, 0x0100 if not video else 0, 0) + matrix + \
        struct.pack(">II", (width << 16) if video else 0, (height << 16) if video else 0)
    mdhd = struct.pack(">I", 0) + struct.pack(">IIIIHH", 0, 0, timescale, 0, 0x55C4, 0)
    hdlr = struct.pack(">I", 0) + struct.pack(">I4s12x", 0, b"vide" if video else b"soun") + \
        (b"VideoHandler\x00" if video else b"SoundHandler\x00")
    dinf = isobmff.box(b"dinf", isobmff.full_box(b"dref", 0, 0, struct.pack(">I", 1) + isobmff.full_box(b"url ", 0, 1, b"")))
    mh = isobmff.full_box(b"vmhd", 0, 1, bytes(8)) if video else isobmff.full_box(b"smhd", 0, 0, bytes(4))
    stsd = isobmff.full_box(b"stsd", 0, 0, struct.pack(">I", 1) + entry_box)
    key = np.asarray([k for _o, _s, _t, k in samples], bool)
    reorder = bool((cts_off != cts_off[0]).any())
    # media time 0 = the first sample's decode time; presentation = (cts - shift) + dts0 on the file's timeline
    edits = []
    if dts0 > 0:
        edits.append((dts0, -1, 0x10000))
    edits.append((0, shift, 0x10000))
    return isobmff.Track(track_id, b"vide" if video else b"soun", codec, timescale, int(deltas.sum()),
                         width if video else 0, height if video else 0, tkhd, mdhd, hdlr, mh + dinf, stsd,
                         arr[:, 1].astype(np.uint64), arr[:, 0].astype(np.uint64), (dts - dts0).astype(np.int64), deltas,
                         cts_off.astype(np.int64) if (reorder or shift) else None, key, not bool(key.all()), edits,
                         (24, len(stsd)))


def read_movie(path: str | Path):
    """Index a Matroska/WebM file into an isobmff.Movie (tracks with sample tables and MP4 sample descriptions).
    Raises isobmff.BmffError when the file cannot be read as Matroska or holds no track with an MP4 mapping."""
    import numpy as np
    from . import isobmff
    path = Path(path)
    rd = _Reader(path)
    try:
        h = rd.header(0)
        if h is None or h[0] != ID_EBML or h[2] == UNKNOWN_SIZE:
            raise isobmff.BmffError("not an EBML file")
        h = rd.header(h[1] + h[2])
        if h is None or h[0] != ID_SEGMENT:
            raise isobmff.BmffError("no Segment")
        seg_start, seg_size = h[1], h[2]
        seg_end = rd.size if seg_size == UNKNOWN_SIZE else min(seg_start + seg_size, rd.size)
        scale_ns, duration = 1000000, 0.0
        tracks = {}                                     # number -> dict
        blocks = {}                                     # number -> list of (offset, size, time, key)
        pos = seg_start
        while pos < seg_end:
            h = rd.header(pos)
            if h is None:
                break
            eid, p, sz = h
            if sz == UNKNOWN_SIZE and eid != ID_CLUSTER:
                break
            if eid == ID_INFO:
                for cid, body in _children(rd.read(p, sz)):
                    if cid == ID_TIMESTAMP_SCALE:
                        scale_ns = _uint(body) or 1000000
                    elif cid == ID_DURATION:
                        duration = _float(body)
            elif eid == ID_TRACKS:
                for cid, body in _children(rd.read(p, sz)):
                    if cid != ID_TRACK_ENTRY:
                        continue
                    t = {"number": 0, "type": 0, "codec": "", "private": b"", "default_ns": 0, "w": 0, "h": 0, "ch": 0,
                         "rate": 0.0}
                    for k, v in _children(body):
                        if k == ID_TRACK_NUMBER:
                            t["number"] = _uint(v)
                        elif k == ID_TRACK_TYPE:
                            t["type"] = _uint(v)
                        elif k == ID_CODEC_ID:
                            t["codec"] = v.rstrip(b"\x00").decode("ascii", "replace")
                        elif k == ID_CODEC_PRIVATE:
                            t["private"] = bytes(v)
                        elif k == ID_DEFAULT_DURATION:
                            t["default_ns"] = _uint(v)
                        elif k == ID_VIDEO:
                            for k2, v2 in _children(v):
                                if k2 == ID_PIXEL_WIDTH:
                                    t["w"] = _uint(v2)
                                elif k2 == ID_PIXEL_HEIGHT:
                                    t["h"] = _uint(v2)
                        elif k == ID_AUDIO:
                            for k2, v2 in _children(v):
                                if k2 == ID_SAMPLING_FREQ:
                                    t["rate"] = _float(v2)
                                elif k2 == ID_CHANNELS:
                                    t["ch"] = _uint(v2)
                    tracks[t["number"]] = t
                    blocks[t["number"]] = []
            elif eid == ID_CLUSTER:
                end = rd.size if sz == UNKNOWN_SIZE else min(p + sz, rd.size)
                cts = 0
                q = p
                while q < end:
                    hh = rd.header(q)
                    if hh is None:
                        break
                    e2, p2, s2 = hh
                    if s2 == UNKNOWN_SIZE:
                        break
                    if e2 in (ID_CLUSTER, ID_TRACKS, ID_INFO) and sz == UNKNOWN_SIZE:
                        break                               # the unknown-size cluster ended: a sibling begins here
                    if e2 == ID_CLUSTER_TIMESTAMP:
                        cts = _uint(rd.read(p2, s2))
                    elif e2 == ID_SIMPLE_BLOCK or e2 == ID_BLOCK_GROUP:
                        bp, bs, key = p2, s2, True
                        if e2 == ID_BLOCK_GROUP:
                            bp = bs = -1
                            key = True
                            g = p2
                            while g < p2 + s2:
                                gh = rd.header(g)
                                if gh is None or gh[2] == UNKNOWN_SIZE:
                                    break
                                if gh[0] == ID_BLOCK:
                                    bp, bs = gh[1], gh[2]
                                elif gh[0] == ID_REFERENCE_BLOCK:
                                    key = False
                                g = gh[1] + gh[2]
                        if bp >= 0 and bs >= 4:
                            head = rd.read(bp, min(bs, 12))
                            r = _read_size(head, 0)
                            if r is not None and r[1] + 3 <= len(head):
                                num, hp = r
                                rel = struct.unpack_from(">h", head, hp)[0]
                                flags = head[hp + 2]
                                if flags & 0x06:
                                    raise isobmff.BmffError("laced Matroska blocks are not supported")
                                if e2 == ID_SIMPLE_BLOCK:
                                    key = bool(flags & 0x80)
                                if num in blocks:
                                    blocks[num].append((bp + hp + 3, bs - hp - 3, cts + rel, key))
                    q = p2 + s2
                pos = q if sz == UNKNOWN_SIZE else p + sz
                continue
            pos = p + sz
    finally:
        rd.close()
    timescale = max(1, int(round(1e9 / scale_ns)))          # ticks per second of the block timestamps (usually 1000)
    movie = isobmff.Movie(path, timescale, int(round(duration)) if duration > 0 else 0)
    next_id = 1
    for num in sorted(tracks):
        t, bl = tracks[num], blocks[num]
        if not bl or t["type"] not in (1, 2):
            continue
        entry = _sample_entry(t["codec"], t["private"], t["w"], t["h"], t["ch"], t["rate"])
        if entry is None:
            continue
        codec, entry_box = entry
        default = int(round(t["default_ns"] / scale_ns)) if t["default_ns"] else 0
        tr = build_track(next_id, t["type"] == 1, codec, entry_box, timescale, t["w"], t["h"], bl, default)
        movie.tracks.append(tr)
        next_id += 1
    if not movie.tracks:
        raise isobmff.BmffError("no Matroska track with an MP4 mapping")
    if not movie.duration:
        movie.duration = max(int(t_.dts[-1] + t_.deltas[-1]) for t_ in movie.tracks)
    return movie

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

from . import isobmff


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


def _sps_size(sps: bytes) -> tuple[int, int]:
    """Display size from an SPS NAL via the library's parser (vt_h264_scan on a tiny Annex-B blob)."""
    from . import _lib
    blob = np.frombuffer(b"\x00\x00\x00\x01" + sps + b"\x00\x00\x00\x01\x68\xce\x38\x80", np.uint8)
    info = _lib.StreamInfo()
    _lib.check(_lib.lib().vt_h264_scan(blob.ctypes.data, blob.size, ctypes.byref(info), None, None, None, 0))
    return info.width, info.height


def _parse_avcc(stsd: bytes, lo: int, hi: int):
    """(nal_length_size, first SPS, first PPS) from the avcC box of an avc1/avc3 sample entry, or None."""
    # VisualSampleEntry: 8 bytes SampleEntry + 70 bytes visual fields, then child boxes
    at = isobmff.find_box(stsd, lo + 78, hi, b"avcC")
    if at is None:
        return None
    a, e = at
    if e - a < 7:
        return None
    nls = (stsd[a + 4] & 3) + 1
    n_sps = stsd[a + 5] & 31
    p = a + 6
    sps = pps = b""
    for _ in range(n_sps):
        if p + 2 > e:
            return None
        ln = struct.unpack_from(">H", stsd, p)[0]
        sps = sps or stsd[p + 2:p + 2 + ln]
        p += 2 + ln
    if p >= e:
        return nls, sps, pps
    n_pps = stsd[p]
    p += 1
    for _ in range(n_pps):
        if p + 2 > e:
            break
        ln = struct.unpack_from(">H", stsd, p)[0]
        pps = pps or stsd[p + 2:p + 2 + ln]
        p += 2 + ln
    return nls, sps, pps


def _walk_avc_samples(data: np.ndarray, offsets: np.ndarray, sizes: np.ndarray, nls: int, max_nals: int = 64):
    """Walk the length-prefixed NAL units of every sample at once (one numpy pass per NAL position).

    Returns (first_vcl_offset uint64 [n], first_vcl_size uint32 [n], vcl_count int32 [n], ok bool [n]);
    first_vcl_offset points at the NAL header byte of the sample's first slice NAL (types 1..5).  ok is False for
    samples whose framing runs past the sample or that hold more than max_nals units."""
    n = offsets.size
    pos = offsets.astype(np.int64).copy()
    end = pos + sizes.astype(np.int64)
    limit = int(data.size)
    first_off = np.zeros(n, np.uint64)
    first_size = np.zeros(n, np.uint32)
    count = np.zeros(n, np.int32)
    ok = np.ones(n, bool)
    active = np.nonzero(pos + nls < end)[0]
    ok[(sizes.astype(np.int64) <= nls)] = False
    for _ in range(max_nals):
        if active.size == 0:
            break
        p = pos[active]
        bad = p + nls + 1 > limit
        if bad.any():
            ok[active[bad]] = False
            active, p = active[~bad], p[~bad]
            if active.size == 0:
                break
        ln = np.zeros(active.size, np.int64)
        for b in range(nls):
            ln = (ln << 8) | data[p + b].astype(np.int64)
        hdr = data[p + nls]
        typ = hdr & 31
        nal_end = p + nls + ln
        broken = (ln <= 0) | (nal_end > end[active])
        ok[active[broken]] = False
        vcl = (~broken) & (typ >= 1) & (typ <= 5)
        fresh = vcl & (count[active] == 0)
        first_off[active[fresh]] = (p[fresh] + nls).astype(np.uint64)
        first_size[active[fresh]] = ln[fresh].astype(np.uint32)
        count[active[vcl]] += 1
        pos[active] = nal_end
        go = (~broken) & (nal_end + nls < end[active])
        active = active[go]
    if active.size:
        ok[active] = False
    return first_off, first_size, count, ok


def index_from_movie(movie: "isobmff.Movie") -> StreamIndex | None:
    """StreamIndex of the first video track of a parsed ISO-BMFF movie (any codec).

    For H.264 (`avc1`/`avc3`) the samples' NAL framing is walked so that `nal_offsets` points at each picture's first
    slice NAL; other codecs get whole-sample offsets and `extra["decodable"] = False` (the pixel pass cannot run, the
    stream copy can)."""
    t = movie.video_track()
    if t is None:
        return None
    n = t.n
    mts = movie.timescale or 1000
    vals, counts = np.unique(t.deltas, return_counts=True)
    delta = int(vals[np.argmax(counts)]) if vals.size else 0
    if delta <= 0 or t.timescale <= 0:
        delta, fps_n = 1, 0
    else:
        fps_n = t.timescale
    g = gcd(fps_n, delta) or 1
    empty_s, media_time = t.edit_shift(mts)
    first_cts = int(t.cts_off[0]) if t.cts_off is not None else 0
    cfr = bool(vals.size == 1 and (t.cts_off is None or (t.cts_off == first_cts).all())
               and empty_s == 0.0 and media_time == first_cts)
    extra = {"cfr": cfr, "codec": t.codec.decode("latin-1"), "movie": movie, "track_id": t.track_id,
             "decodable": False, "times": None if cfr else t.pres_times(mts),
             "tracks": [(x.handler.decode("latin-1"), x.codec.decode("latin-1")) for x in movie.tracks]}
    sps = pps = b""
    nal_off, nal_size = t.offsets.astype(np.uint64), np.minimum(t.sizes, 0xFFFFFFFF).astype(np.uint32)
    width, height = t.width, t.height
    if t.codec in (b"avc1", b"avc3"):
        cfg = _parse_avcc(t.stsd, *t.entry_payload)
        if cfg is not None:
            nls, sps, pps = cfg
            extra["nal_length_size"] = nls
            try:
                data = np.memmap(movie.path, dtype=np.uint8, mode="r")
                f_off, f_size, vcl, ok = _walk_avc_samples(data, t.offsets, t.sizes, nls)
                extra["single_slice"] = bool(ok.all() and (vcl == 1).all())
                if extra["single_slice"] and sps and pps:
                    nal_off, nal_size = f_off, f_size
                    extra["decodable"] = True        # framing is what the PCM-intra decoder indexes; slices are
                                                     # classified later (classify_pcm), which may still refuse
            except (OSError, ValueError):
                extra["single_slice"] = False
            if sps:
                try:
                    width, height = _sps_size(sps)
                except Exception:  # noqa: BLE001 - no library / odd SPS: keep the sample entry's size
                    pass
    return StreamIndex("mp4", Path(movie.path), int(width), int(height), fps_n // g, delta // g, n,
                       movie.duration_seconds(), nal_off, nal_size, t.sync.copy(), bytes(sps), bytes(pps), extra)


def probe_mp4(path: Path) -> StreamIndex | None:
    try:
        movie = isobmff.read_movie(path)
    except (isobmff.BmffError, struct.error, OSError, IndexError):
        return None
    return index_from_movie(movie)


def classify_pcm(idx: StreamIndex) -> bool:
    """True when every picture of `idx` is one the PCM-intra decoder (vt_h264_pcm_decode) handles: the result is cached
    in idx.extra["pcm_intra_only"].  Unknown or unparsable streams are False (fail closed)."""
    cached = idx.extra.get("pcm_intra_only")
    if cached is not None:
        return bool(cached)
    ok = False
    if idx.kind == "mp4" and idx.extra.get("decodable") and idx.n_frames:
        from . import _lib
        try:
            host = np.memmap(idx.path, dtype=np.uint8, mode="r")
            pay = np.zeros(idx.n_frames, np.uint64)
            sps = np.frombuffer(idx.sps, np.uint8)
            pps = np.frombuffer(idx.pps, np.uint8)
            offs = np.ascontiguousarray(idx.nal_offsets, dtype=np.uint64)
            sizes = np.ascontiguousarray(idx.nal_sizes, dtype=np.uint32)
            rc = _lib.lib().vt_h264_pcm_layout_ps(host.ctypes.data, host.size, sps.ctypes.data, sps.size,
                                                  pps.ctypes.data, pps.size, offs.ctypes.data, sizes.ctypes.data,
                                                  idx.n_frames, pay.ctypes.data)
            ok = rc == 0
            if ok:
                idx.extra["payload"] = pay
        except (OSError, ValueError):
            ok = False
    idx.extra["pcm_intra_only"] = ok
    return ok


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
    """Index a media file.  Returns None when the file is not a container this layer can cut (ISO-BMFF or
    Matroska/WebM or FLV with a video track that has an MP4 mapping, or a raw Annex-B H.264 stream);
    `container_duration` still knows AVI."""
    path = Path(path)
    if not path.is_file() or path.stat().st_size < 16:
        return None
    with open(path, "rb") as f:
        head = f.read(12)
    if head[4:8] in (b"ftyp", b"moov", b"free", b"mdat", b"styp", b"wide", b"skip", b"pnot"):
        return probe_mp4(path)
    if head[:4] == b"\x00\x00\x00\x01" or head[:3] == b"\x00\x00\x01":
        return probe_h264(path)
    if head[:4] == b"\x1a\x45\xdf\xa3":
        from . import matroska
        try:
            movie = matroska.read_movie(path)
        except (isobmff.BmffError, struct.error, OSError, IndexError, ValueError):
            return None
        idx = index_from_movie(movie)
        if idx is not None:
            idx.duration = matroska.duration_seconds(path) or idx.duration
            idx.extra["container"] = "matroska"
        return idx
    if head[:3] == b"FLV" and head[3] == 1:
        from . import flv
        try:
            movie = flv.read_movie(path)
        except (isobmff.BmffError, struct.error, OSError, IndexError, ValueError):
            return None
        idx = index_from_movie(movie)
        if idx is not None:
            idx.duration = _flv_duration(path) or idx.duration
            idx.extra["container"] = "flv"
        return idx
    return None


def container_duration(path: Path) -> float:
    """Seconds that `ffprobe -show_entries format=duration` prints for the file (0.0 when unknown).

    ISO-BMFF: the movie header (any codec, audio-only files included); Matroska/WebM: Segment Info Duration x
    TimestampScale; AVI: the longest stream header; raw Annex-B H.264: picture count over the VUI frame rate."""
    path = Path(path)
    if not path.is_file() or path.stat().st_size < 12:
        return 0.0
    with open(path, "rb") as f:
        head = f.read(12)
    if head[:4] == b"\x1a\x45\xdf\xa3":
        from . import matroska
        return matroska.duration_seconds(path)
    if head[:4] == b"RIFF" and head[8:12] == b"AVI ":
        return _avi_duration(path)
    if head[:3] == b"FLV" and head[3] == 1:
        return _flv_duration(path)
    if head[4:8] in (b"ftyp", b"moov", b"free", b"mdat", b"styp", b"wide", b"skip", b"pnot"):
        try:
            return isobmff.read_movie(path).duration_seconds()
        except (isobmff.BmffError, struct.error, OSError, IndexError):
            return 0.0
    idx = probe(path)
    return float(idx.duration) if idx is not None else 0.0


def _avi_duration(path: Path) -> float:
    """Longest stream of an AVI file: strh dwLength * dwScale / dwRate (what libavformat's avi demuxer derives),
    falling back to avih dwTotalFrames * dwMicroSecPerFrame."""
    with open(path, "rb") as f:
        buf = f.read(1 << 16)
    best = 0.0
    pos = buf.find(b"avih")
    fallback = 0.0
    if pos >= 0 and pos + 8 + 20 <= len(buf):
        us_per_frame, = struct.unpack_from("<I", buf, pos + 8)
        total, = struct.unpack_from("<I", buf, pos + 8 + 16)
        fallback = us_per_frame * total / 1e6
    pos = 0
    while True:
        pos = buf.find(b"strh", pos)
        if pos < 0 or pos + 8 + 36 > len(buf):
            break
        scale, rate, _start, length = struct.unpack_from("<IIII", buf, pos + 8 + 20)
        if rate:
            best = max(best, length * scale / rate)
        pos += 4
    return best if best > 0 else fallback


def _flv_duration(path: Path) -> float:
    """FLV (the downloader's `best[height<=N]` fallback can deliver it, src/downloader/video_downloader.py:56): the
    `duration` number of the onMetaData script tag, which is what libavformat's flv demuxer reports; files without it
    get the timestamp of their last tag (found through the trailing PreviousTagSize), as the demuxer does by seeking."""
    size = path.stat().st_size
    with open(path, "rb") as f:
        buf = f.read(min(size, 1 << 16))
        at = buf.find(b"onMetaData")
        if at >= 0:
            k = buf.find(b"\x00\x08duration\x00", at)           # AMF0: key length, key, type 0 (number), float64
            if k >= 0 and k + 19 <= len(buf):
                d, = struct.unpack_from(">d", buf, k + 11)
                if d == d and 0 < d < 1e9:
                    return float(d)
        if size < 13 + 15:
            return 0.0
        f.seek(size - 4)
        prev, = struct.unpack(">I", f.read(4))
        if prev < 11 or prev + 4 > size - 9:
            return 0.0
        f.seek(size - 4 - prev)
        tag = f.read(11)
    if len(tag) < 11 or tag[0] & 0x1F not in (8, 9, 18):
        return 0.0
    ms = (tag[7] << 24) | (tag[4] << 16) | (tag[5] << 8) | tag[6]
    return ms / 1000.0


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

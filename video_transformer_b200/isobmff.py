"""Codec-agnostic ISO-BMFF (MP4 / MOV / 3GP / M4A) reader and all-track stream-copy cutter.  Host-side byte I/O.

What it replaces in the reference (both run inside ffmpeg/ffprobe child processes there):
  * ffprobe `format=duration`                          /root/reference/src/utils/video_utils.py:7-38
  * `ffmpeg -ss S -i IN -t D -movflags +faststart -c copy OUT`
                                                       /root/reference/src/utils/video_segmenter.py:118-137
A stream copy never looks inside a sample: every track is carried with its sample description (`stsd`) copied
verbatim, its sample tables re-cut to the selected range and its sample bytes copied as they are, so the cut works
for any codec (H.264 with several NALs per sample, HEVC, MPEG-4 part 2, VP9, AV1, AAC, PCM ...), with B-frame
reordering (`ctts`), edit lists (`edts/elst`), 32- and 64-bit chunk offsets and any chunking.

Selection rule (derived from the command line above): s = round(start, 3),
d = round(end - start, 3); the reference track (first video track) keeps the samples whose presentation time is in
[s, s + d), extended back to the last sync sample at or before the first of them; every other track keeps the
samples that overlap the reference track's kept time span.  The movie box is written before the media data
(`+faststart`).  Fragmented sources (`moof`/`traf`/`trun`) are indexed like any other; the output is never fragmented.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

AUDIO_PREROLL_SAMPLES = {b"mp4a": 1, b"Opus": 4, b".mp3": 1, b"ac-3": 1, b"ec-3": 1}

CONTAINER_BOXES = (b"moov", b"trak", b"edts", b"mdia", b"minf", b"dinf", b"stbl", b"mvex", b"moof", b"traf")


class BmffError(ValueError):
    pass


# ---- box walking ------------------------------------------------------------------------------------------------
def iter_boxes(buf, start: int, end: int):
    """Yield (type, payload_start, box_end) for the boxes of buf[start:end]; stops at the first malformed header."""
    pos = start
    while pos + 8 <= end:
        size, kind = struct.unpack_from(">I4s", buf, pos)
        head = 8
        if size == 1:
            if pos + 16 > end:
                return
            size = struct.unpack_from(">Q", buf, pos + 8)[0]
            head = 16
        elif size == 0:
            size = end - pos
        if size < head or pos + size > end:
            return
        yield kind, pos + head, pos + size
        pos += size


def find_box(buf, start: int, end: int, kind: bytes):
    for k, s, e in iter_boxes(buf, start, end):
        if k == kind:
            return s, e
    return None


def box(kind: bytes, payload: bytes) -> bytes:
    n = 8 + len(payload)
    if n >= 1 << 32:
        return struct.pack(">I4sQ", 1, kind, n + 8) + payload
    return struct.pack(">I4s", n, kind) + payload


def full_box(kind: bytes, version: int, flags: int, payload: bytes) -> bytes:
    return box(kind, struct.pack(">I", (version << 24) | flags) + payload)


def top_level(path: Path):
    """[(type, box_start, payload_start, box_end)] of a file's top-level boxes, reading headers only."""
    out = []
    size_file = os.path.getsize(path)
    with open(path, "rb") as f:
        pos = 0
        while pos + 8 <= size_file:
            f.seek(pos)
            hdr = f.read(16)
            if len(hdr) < 8:
                break
            size, kind = struct.unpack_from(">I4s", hdr, 0)
            head = 8
            if size == 1:
                if len(hdr) < 16:
                    break
                size = struct.unpack_from(">Q", hdr, 8)[0]
                head = 16
            elif size == 0:
                size = size_file - pos
            if size < head:
                break
            end = min(pos + size, size_file)      # a truncated last box (e.g. mdat of an interrupted download)
            out.append((kind, pos, pos + head, end))
            pos += size
    return out


# ---- data model -------------------------------------------------------------------------------------------------
@dataclass
class Track:
    track_id: int
    handler: bytes                  # b"vide", b"soun", b"text", ...
    codec: bytes                    # fourcc of the first sample entry (b"avc1", b"hev1", b"mp4v", b"vp09", b"mp4a", ...)
    timescale: int
    media_duration: int             # mdhd duration (media timescale)
    width: int                      # video: from the sample entry; else 0
    height: int
    tkhd: bytes                     # whole box payload incl. version/flags (copied, duration patched on write)
    mdhd: bytes
    hdlr: bytes
    minf_other: bytes               # the boxes of minf except stbl (vmhd/smhd/nmhd, dinf), as whole boxes
    stsd: bytes                     # whole stsd box, copied verbatim
    sizes: np.ndarray               # uint64 [n]
    offsets: np.ndarray             # uint64 [n] file offsets of the samples
    dts: np.ndarray                 # int64 [n] decode times (media timescale), dts[0] = 0
    deltas: np.ndarray              # int64 [n] sample durations
    cts_off: np.ndarray | None      # int64 [n] composition offsets (None without ctts)
    sync: np.ndarray                # bool [n]
    has_stss: bool
    edits: list                     # [(segment_duration (movie ts), media_time, rate_16_16)]
    entry_payload: tuple = (0, 0)   # (start, end) of the first sample entry's payload inside `stsd`
    desc: np.ndarray | None = None  # int64 [n] sample description index (1-based) of every sample; None = all use entry 1
    unit: tuple | None = None       # (bytes, duration) of ONE sample when the arrays above describe GROUPS of equal
                                    # samples (uniform tracks with millions of samples, e.g. PCM audio: see _parse_track)

    @property
    def n(self) -> int:
        return int(self.sizes.size)

    def edit_shift(self, movie_timescale: int):
        """(empty duration in seconds, media_time of the first normal edit)."""
        empty = 0
        media_time = 0
        for seg_dur, mt, _rate in self.edits:
            if mt < 0:
                empty += seg_dur
            else:
                media_time = mt
                break
        return (empty / movie_timescale if movie_timescale else 0.0), media_time

    def pres_times(self, movie_timescale: int) -> np.ndarray:
        """Presentation time (seconds, float64) of every sample, decode order."""
        empty_s, mt = self.edit_shift(movie_timescale)
        cts = self.dts if self.cts_off is None else self.dts + self.cts_off
        return (cts - mt).astype(np.float64) / float(self.timescale) + empty_s


@dataclass
class Movie:
    path: Path
    timescale: int
    duration: int                   # mvhd duration (movie timescale); 0 when unknown
    tracks: list = field(default_factory=list)
    ftyp: bytes = b""               # whole ftyp box (copied)
    fragmented: bool = False
    fragment_duration_s: float = 0.0

    def duration_seconds(self) -> float:
        """What `ffprobe -show_entries format=duration` prints: libavformat keeps the movie header's duration
        (rescaled to microseconds); a fragmented file without it gets the longest track."""
        if self.duration and self.timescale:
            return _us(self.duration, self.timescale)
        if self.fragment_duration_s > 0:
            return self.fragment_duration_s
        best = 0.0
        for t in self.tracks:
            if t.timescale and t.n:
                best = max(best, _us(int(t.dts[-1] + t.deltas[-1]), t.timescale))
        return best

    def video_track(self):
        for t in self.tracks:
            if t.handler == b"vide" and t.n:
                return t
        return None


def _us(value: int, timescale: int) -> float:
    """value/timescale rounded to microseconds the way av_rescale does (nearest, half away from zero)."""
    return ((value * 1000000 + timescale // 2) // timescale) / 1e6


# ---- reading ----------------------------------------------------------------------------------------------------
def _be(buf, dtype, count, offset):
    if offset + count * np.dtype(dtype).itemsize > len(buf):
        raise BmffError("table runs past its box")
    return np.frombuffer(buf, dtype, count, offset)


MAX_TRACK_SAMPLES = 1 << 25      # per-sample tables beyond this are refused (a corrupt count must not allocate gigabytes)
UNIFORM_MIN_SAMPLES = 1 << 21    # uniform tracks (one size, one duration, every sample sync) larger than this are grouped
UNIFORM_GROUP = 1024             # ... into runs of at most this many samples inside a chunk (libavformat's PCM packets)


def _expand_runs(values: np.ndarray, counts: np.ndarray, n: int) -> np.ndarray:
    """np.repeat(values, counts) cut off at n elements WITHOUT materialising more: counts come from the file and a
    corrupt run length (0xCE000000 ...) would otherwise allocate and fill gigabytes before the result is trimmed."""
    c = np.clip(np.asarray(counts, np.int64), 0, max(n, 0))
    before = np.cumsum(c) - c
    c = np.clip(n - before, 0, c)
    return np.repeat(values, c)


def _parse_track(moov: bytes, s: int, e: int) -> Track | None:
    tkhd = find_box(moov, s, e, b"tkhd")
    mdia = find_box(moov, s, e, b"mdia")
    if tkhd is None or mdia is None:
        return None
    tk = moov[tkhd[0]:tkhd[1]]
    v = tk[0]
    track_id = struct.unpack_from(">I", tk, 20 if v == 1 else 12)[0]
    edits = []
    edts = find_box(moov, s, e, b"edts")
    if edts:
        elst = find_box(moov, edts[0], edts[1], b"elst")
        if elst:
            ev = moov[elst[0]]
            n = struct.unpack_from(">I", moov, elst[0] + 4)[0]
            p = elst[0] + 8
            for _ in range(n):
                if ev == 1:
                    if p + 20 > elst[1]:
                        break
                    sd, mt, rate = struct.unpack_from(">Qqi", moov, p)
                    p += 20
                else:
                    if p + 12 > elst[1]:
                        break
                    sd, mt, rate = struct.unpack_from(">Iii", moov, p)
                    p += 12
                edits.append((int(sd), int(mt), int(rate)))
    mdhd = find_box(moov, mdia[0], mdia[1], b"mdhd")
    hdlr = find_box(moov, mdia[0], mdia[1], b"hdlr")
    minf = find_box(moov, mdia[0], mdia[1], b"minf")
    if mdhd is None or hdlr is None or minf is None:
        return None
    md = moov[mdhd[0]:mdhd[1]]
    if md[0] == 1:
        timescale, mdur = struct.unpack_from(">IQ", md, 20)
    else:
        timescale, mdur = struct.unpack_from(">II", md, 12)
    handler = moov[hdlr[0] + 8:hdlr[0] + 12]
    stbl = find_box(moov, minf[0], minf[1], b"stbl")
    if stbl is None:
        return None
    minf_other = b""
    for k, bs, be in iter_boxes(moov, minf[0], minf[1]):
        if k != b"stbl":
            hdr = 16 if struct.unpack_from(">I", moov, bs - 8)[0] != be - (bs - 8) else 8
            minf_other += moov[bs - hdr:be]
    tabs = {k: (bs, be) for k, bs, be in iter_boxes(moov, stbl[0], stbl[1])}
    if b"stsd" not in tabs:
        return None
    sd_s, sd_e = tabs[b"stsd"]
    stsd_box = box(b"stsd", moov[sd_s:sd_e])
    codec, width, height = b"", 0, 0
    entry_payload = (0, 0)
    if struct.unpack_from(">I", moov, sd_s + 4)[0] >= 1 and sd_s + 16 <= sd_e:
        esize, codec = struct.unpack_from(">I4s", moov, sd_s + 8)
        entry_payload = (24, min(16 + esize, len(stsd_box)))
        if handler == b"vide" and sd_s + 8 + 36 <= sd_e:
            width, height = struct.unpack_from(">HH", moov, sd_s + 8 + 8 + 24)
    # sample sizes
    uniform = None
    if b"stsz" in tabs:
        zs = tabs[b"stsz"][0]
        fixed, n = struct.unpack_from(">II", moov, zs + 4)
        if fixed and n > UNIFORM_MIN_SAMPLES:
            uniform = _uniform_track(moov, tabs, int(fixed), int(n))
        if uniform is None:
            if n > MAX_TRACK_SAMPLES:
                raise BmffError("track with %d samples (limit %d)" % (n, MAX_TRACK_SAMPLES))
            sizes = np.full(n, fixed, np.uint64) if fixed else _be(moov, ">u4", n, zs + 12).astype(np.uint64)
    elif b"stz2" in tabs:
        zs = tabs[b"stz2"][0]
        fsize = moov[zs + 7]
        n = struct.unpack_from(">I", moov, zs + 8)[0]
        if n > MAX_TRACK_SAMPLES:
            raise BmffError("track with %d samples (limit %d)" % (n, MAX_TRACK_SAMPLES))
        if fsize == 16:
            sizes = _be(moov, ">u2", n, zs + 12).astype(np.uint64)
        elif fsize == 8:
            sizes = _be(moov, np.uint8, n, zs + 12).astype(np.uint64)
        elif fsize == 4:
            raw = _be(moov, np.uint8, (n + 1) // 2, zs + 12)
            sizes = np.stack([raw >> 4, raw & 15], 1).reshape(-1)[:n].astype(np.uint64)
        else:
            raise BmffError("stz2 field size %d" % fsize)
    else:
        sizes = np.zeros(0, np.uint64)
    if uniform is not None:
        sizes, offsets, deltas, unit = uniform
        n = int(sizes.size)
        dts = np.concatenate(([0], np.cumsum(deltas)[:-1])).astype(np.int64) if n else np.zeros(0, np.int64)
        return Track(track_id, handler, codec, int(timescale), int(mdur), int(width), int(height), tk, md,
                     moov[hdlr[0]:hdlr[1]], minf_other, stsd_box, sizes, offsets, dts, deltas, None, np.ones(n, bool),
                     False, edits, entry_payload, unit=unit)
    n = int(sizes.size)
    # chunk offsets
    if b"co64" in tabs:
        cs = tabs[b"co64"][0]
        n_co = struct.unpack_from(">I", moov, cs + 4)[0]
        chunk_off = _be(moov, ">u8", n_co, cs + 8).astype(np.uint64)
    elif b"stco" in tabs:
        cs = tabs[b"stco"][0]
        n_co = struct.unpack_from(">I", moov, cs + 4)[0]
        chunk_off = _be(moov, ">u4", n_co, cs + 8).astype(np.uint64)
    else:
        chunk_off = np.zeros(0, np.uint64)
    n_co = int(chunk_off.size)
    offsets = np.zeros(n, np.uint64)
    desc = None
    if n and n_co and b"stsc" in tabs:
        ss = tabs[b"stsc"][0]
        n_sc = struct.unpack_from(">I", moov, ss + 4)[0]
        sc = _be(moov, ">u4", 3 * n_sc, ss + 8).reshape(-1, 3).astype(np.int64)
        if n_sc == 0:
            raise BmffError("empty stsc")
        first = np.clip(sc[:, 0] - 1, 0, n_co)
        nxt = np.append(first[1:], n_co)
        per_chunk = _expand_runs(sc[:, 1], np.maximum(nxt - first, 0), n_co)
        if per_chunk.size < n_co:
            per_chunk = np.append(per_chunk, np.zeros(n_co - per_chunk.size, np.int64))
        per_chunk = per_chunk[:n_co]
        first_sample = np.concatenate(([0], np.cumsum(per_chunk)[:-1]))
        chunk_of = _expand_runs(np.arange(n_co), per_chunk, n)
        if chunk_of.size < n:
            raise BmffError("stsc/stco describe %d samples, stsz has %d" % (chunk_of.size, n))
        chunk_of = chunk_of[:n]
        if (sc[:, 2] != 1).any():                 # several sample entries (stsd) in use: remember which sample uses which
            d_chunk = _expand_runs(sc[:, 2], np.maximum(nxt - first, 0), n_co)
            if d_chunk.size < n_co:
                d_chunk = np.append(d_chunk, np.ones(n_co - d_chunk.size, np.int64))
            desc = d_chunk[chunk_of].astype(np.int64)
        excl = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.uint64)
        offsets = chunk_off[chunk_of] + excl - excl[first_sample[chunk_of]]
    elif n:
        raise BmffError("samples without chunk tables")
    # decode times
    deltas = np.zeros(n, np.int64)
    if b"stts" in tabs and n:
        ts_ = tabs[b"stts"][0]
        n_tt = struct.unpack_from(">I", moov, ts_ + 4)[0]
        tt = _be(moov, ">u4", 2 * n_tt, ts_ + 8).reshape(-1, 2).astype(np.int64)
        d = _expand_runs(tt[:, 1], tt[:, 0], n)
        if d.size < n:
            d = np.append(d, np.full(n - d.size, d[-1] if d.size else 0, np.int64))
        deltas = d[:n].copy()
    dts = np.concatenate(([0], np.cumsum(deltas)[:-1])).astype(np.int64) if n else np.zeros(0, np.int64)
    cts_off = None
    if b"ctts" in tabs and n:
        cs = tabs[b"ctts"][0]
        cv = moov[cs]
        n_ct = struct.unpack_from(">I", moov, cs + 4)[0]
        raw = _be(moov, ">u4", 2 * n_ct, cs + 8).reshape(-1, 2)
        cnt = raw[:, 0].astype(np.int64)
        # version 1 is signed by definition; writers also put negative values into version 0 (QuickTime), and
        # libavformat reads the field as signed in both cases
        off = raw[:, 1].astype(np.uint32).view(np.int32).astype(np.int64)
        _ = cv
        c = _expand_runs(off, cnt, n)
        if c.size < n:
            c = np.append(c, np.zeros(n - c.size, np.int64))
        cts_off = c[:n].copy()
    sync = np.ones(n, bool)
    has_stss = b"stss" in tabs
    if has_stss and n:
        s0 = tabs[b"stss"][0]
        n_ss = struct.unpack_from(">I", moov, s0 + 4)[0]
        k = _be(moov, ">u4", n_ss, s0 + 8).astype(np.int64) - 1
        sync[:] = False
        sync[k[(k >= 0) & (k < n)]] = True
    return Track(track_id, handler, codec, int(timescale), int(mdur), int(width), int(height), tk, md,
                 moov[hdlr[0]:hdlr[1]], minf_other, stsd_box, sizes, offsets, dts, deltas, cts_off, sync, has_stss,
                 edits, entry_payload, desc=desc)


def _chunk_table(moov: bytes, tabs: dict):
    """(chunk offsets u64[n_co], samples per chunk i64[n_co]) from stco|co64 + stsc, or None."""
    if b"co64" in tabs:
        cs = tabs[b"co64"][0]
        chunk_off = _be(moov, ">u8", struct.unpack_from(">I", moov, cs + 4)[0], cs + 8).astype(np.uint64)
    elif b"stco" in tabs:
        cs = tabs[b"stco"][0]
        chunk_off = _be(moov, ">u4", struct.unpack_from(">I", moov, cs + 4)[0], cs + 8).astype(np.uint64)
    else:
        return None
    n_co = int(chunk_off.size)
    if not n_co or b"stsc" not in tabs:
        return None
    ss = tabs[b"stsc"][0]
    n_sc = struct.unpack_from(">I", moov, ss + 4)[0]
    if n_sc == 0:
        return None
    sc = _be(moov, ">u4", 3 * n_sc, ss + 8).reshape(-1, 3).astype(np.int64)
    first = np.clip(sc[:, 0] - 1, 0, n_co)
    nxt = np.append(first[1:], n_co)
    per_chunk = _expand_runs(sc[:, 1], np.maximum(nxt - first, 0), n_co)
    if per_chunk.size < n_co:
        per_chunk = np.append(per_chunk, np.zeros(n_co - per_chunk.size, np.int64))
    return chunk_off, per_chunk[:n_co]


def _uniform_track(moov: bytes, tabs: dict, fixed: int, n: int):
    """Grouped description of a track whose n samples all have `fixed` bytes, one duration, no composition offsets and
    no sync table -- uncompressed audio, where a sample is ONE audio frame: two hours at 48 kHz are 345 million samples,
    and per-sample tables of that length cost gigabytes.  Samples are grouped into runs of at most UNIFORM_GROUP inside a
    chunk (the packets libavformat's demuxer forms for PCM, so a stream copy cuts where ffmpeg's would); the arrays then
    describe groups, `unit` = (bytes, duration) of one sample lets the writer emit true per-sample tables again
    (fixed-size stsz, one stts run).  Returns (sizes, offsets, deltas, unit) or None when the track is not uniform."""
    if b"ctts" in tabs or b"stss" in tabs or b"stts" not in tabs:
        return None
    if b"stsc" in tabs:
        ss = tabs[b"stsc"][0]
        n_sc = struct.unpack_from(">I", moov, ss + 4)[0]
        if (_be(moov, ">u4", 3 * n_sc, ss + 8).reshape(-1, 3)[:, 2] != 1).any():
            return None                           # several sample entries in use: the per-sample path keeps track of them
    ts_ = tabs[b"stts"][0]
    n_tt = struct.unpack_from(">I", moov, ts_ + 4)[0]
    tt = _be(moov, ">u4", 2 * n_tt, ts_ + 8).reshape(-1, 2).astype(np.int64)
    tt = tt[tt[:, 0] > 0]
    if tt.shape[0] == 0 or (tt[:, 1] != tt[0, 1]).any():
        return None
    delta = int(tt[0, 1])
    table = _chunk_table(moov, tabs)
    if table is None:
        return None
    chunk_off, per_chunk = table
    per_chunk = np.clip(per_chunk, 0, n)
    room = n - (np.cumsum(per_chunk) - per_chunk)
    per_chunk = np.clip(room, 0, per_chunk)                    # the chunk table may not describe more than n samples
    if int(per_chunk.sum()) < n:
        raise BmffError("stsc/stco describe %d samples, stsz has %d" % (int(per_chunk.sum()), n))
    groups = -(-per_chunk // UNIFORM_GROUP)                     # groups per chunk
    n_g = int(groups.sum())
    if n_g > MAX_TRACK_SAMPLES:
        raise BmffError("track with %d sample groups (limit %d)" % (n_g, MAX_TRACK_SAMPLES))
    chunk_of = np.repeat(np.arange(per_chunk.size), groups)
    k_in_chunk = np.arange(n_g) - np.repeat(np.cumsum(groups) - groups, groups)
    count = np.minimum(UNIFORM_GROUP, per_chunk[chunk_of] - k_in_chunk * UNIFORM_GROUP).astype(np.int64)
    offsets = chunk_off[chunk_of] + (k_in_chunk * UNIFORM_GROUP * fixed).astype(np.uint64)
    return (count * fixed).astype(np.uint64), offsets.astype(np.uint64), count * delta, (fixed, delta)


def _fragments_duration(path: Path, tops, moov: bytes, movie: Movie) -> float:
    """Longest track of a fragmented file: mehd when present, else the sum of the fragments' sample durations."""
    mvex = find_box(moov, 0, len(moov), b"mvex")
    defaults = {}
    if mvex:
        mehd = find_box(moov, mvex[0], mvex[1], b"mehd")
        if mehd and movie.timescale:
            v = moov[mehd[0]]
            dur = struct.unpack_from(">Q" if v == 1 else ">I", moov, mehd[0] + 4)[0]
            if dur:
                return _us(dur, movie.timescale)
        for k, s, e in iter_boxes(moov, mvex[0], mvex[1]):
            if k == b"trex":
                tid, _desc, ddur = struct.unpack_from(">III", moov, s + 4)
                defaults[tid] = ddur
    scales = {t.track_id: t.timescale for t in movie.tracks}
    total = {}
    with open(path, "rb") as f:
        for kind, b0, p0, b1 in tops:
            if kind != b"moof":
                continue
            f.seek(p0)
            moof = f.read(b1 - p0)
            for k, s, e in iter_boxes(moof, 0, len(moof)):
                if k != b"traf":
                    continue
                tfhd = find_box(moof, s, e, b"tfhd")
                if not tfhd:
                    continue
                fl = struct.unpack_from(">I", moof, tfhd[0])[0] & 0xFFFFFF
                tid = struct.unpack_from(">I", moof, tfhd[0] + 4)[0]
                p = tfhd[0] + 8 + (8 if fl & 1 else 0) + (4 if fl & 2 else 0)
                ddur = defaults.get(tid, 0)
                if fl & 8:
                    ddur = struct.unpack_from(">I", moof, p)[0]
                for k2, s2, e2 in iter_boxes(moof, s, e):
                    if k2 != b"trun":
                        continue
                    tf = struct.unpack_from(">I", moof, s2)[0] & 0xFFFFFF
                    cnt = struct.unpack_from(">I", moof, s2 + 4)[0]
                    q = s2 + 8 + (4 if tf & 1 else 0) + (4 if tf & 4 else 0)
                    stride = 4 * (bool(tf & 0x100) + bool(tf & 0x200) + bool(tf & 0x400) + bool(tf & 0x800))
                    if tf & 0x100:
                        d = int(_be(moof, ">u4", cnt * stride // 4, q).reshape(cnt, -1)[:, 0].astype(np.int64).sum())
                    else:
                        d = cnt * ddur
                    total[tid] = total.get(tid, 0) + d
    best = 0.0
    for tid, d in total.items():
        if scales.get(tid):
            best = max(best, _us(d, scales[tid]))
    return best


def _append_fragments(path: Path, tops, moov: bytes, movie: Movie) -> None:
    """Index the samples of a fragmented file (moof/traf/tfhd/tfdt/trun) into the tracks' tables, so that a
    fragmented source is cut like any other (the output is an ordinary, non-fragmented faststart MP4)."""
    trex = {}
    mvex = find_box(moov, 0, len(moov), b"mvex")
    if mvex:
        for k, s, e in iter_boxes(moov, mvex[0], mvex[1]):
            if k == b"trex" and e - s >= 24:
                tid, _desc, ddur, dsize, dflags = struct.unpack_from(">IIIII", moov, s + 4)
                trex[tid] = (ddur, dsize, dflags)
    by_id = {t.track_id: t for t in movie.tracks}
    acc = {tid: {"sizes": [], "offsets": [], "deltas": [], "cts": [], "sync": [], "dts": [], "next_dts": None, "has_cts": False}
           for tid in by_id}
    with open(path, "rb") as f:
        for kind, b0, p0, b1 in tops:
            if kind != b"moof":
                continue
            f.seek(p0)
            moof = f.read(b1 - p0)
            prev_end = None                                   # end of the previous traf's data (legacy base rule)
            for k, s, e in iter_boxes(moof, 0, len(moof)):
                if k != b"traf":
                    continue
                tfhd = find_box(moof, s, e, b"tfhd")
                if not tfhd:
                    continue
                fl = struct.unpack_from(">I", moof, tfhd[0])[0] & 0xFFFFFF
                tid = struct.unpack_from(">I", moof, tfhd[0] + 4)[0]
                a = acc.get(tid)
                if a is None:
                    continue
                p = tfhd[0] + 8
                base = None
                if fl & 0x1:
                    base = struct.unpack_from(">Q", moof, p)[0]
                    p += 8
                if fl & 0x2:
                    p += 4
                ddur, dsize, dflags = trex.get(tid, (0, 0, 0))
                if fl & 0x8:
                    ddur = struct.unpack_from(">I", moof, p)[0]
                    p += 4
                if fl & 0x10:
                    dsize = struct.unpack_from(">I", moof, p)[0]
                    p += 4
                if fl & 0x20:
                    dflags = struct.unpack_from(">I", moof, p)[0]
                    p += 4
                if base is None:
                    base = b0 if (fl & 0x020000 or prev_end is None) else prev_end
                tfdt = find_box(moof, s, e, b"tfdt")
                if tfdt:
                    v = moof[tfdt[0]]
                    a["next_dts"] = struct.unpack_from(">Q" if v == 1 else ">I", moof, tfdt[0] + 4)[0]
                elif a["next_dts"] is None:
                    a["next_dts"] = 0
                run_pos = base
                for k2, s2, e2 in iter_boxes(moof, s, e):
                    if k2 != b"trun":
                        continue
                    tv = moof[s2]
                    tf = struct.unpack_from(">I", moof, s2)[0] & 0xFFFFFF
                    cnt = struct.unpack_from(">I", moof, s2 + 4)[0]
                    q = s2 + 8
                    if tf & 0x1:
                        run_pos = base + struct.unpack_from(">i", moof, q)[0]
                        q += 4
                    first_flags = None
                    if tf & 0x4:
                        first_flags = struct.unpack_from(">I", moof, q)[0]
                        q += 4
                    ncol = bool(tf & 0x100) + bool(tf & 0x200) + bool(tf & 0x400) + bool(tf & 0x800)
                    tab = _be(moof, ">u4", cnt * ncol, q).reshape(cnt, ncol).astype(np.int64) if ncol else None
                    col = 0
                    if tf & 0x100:
                        dur = tab[:, col]; col += 1
                    else:
                        dur = np.full(cnt, ddur, np.int64)
                    if tf & 0x200:
                        size = tab[:, col]; col += 1
                    else:
                        size = np.full(cnt, dsize, np.int64)
                    if tf & 0x400:
                        flags = tab[:, col]; col += 1
                    else:
                        flags = np.full(cnt, dflags, np.int64)
                        if first_flags is not None and cnt:
                            flags[0] = first_flags
                    if tf & 0x800:
                        cts = tab[:, col].astype(np.uint32).view(np.int32).astype(np.int64) if tv else tab[:, col]
                        a["has_cts"] = True
                    else:
                        cts = np.zeros(cnt, np.int64)
                    offs = run_pos + np.concatenate(([0], np.cumsum(size)[:-1])) if cnt else np.zeros(0, np.int64)
                    dts = a["next_dts"] + np.concatenate(([0], np.cumsum(dur)[:-1])) if cnt else np.zeros(0, np.int64)
                    a["sizes"].append(size); a["offsets"].append(offs); a["deltas"].append(dur); a["cts"].append(cts)
                    a["sync"].append((flags & 0x00010000) == 0); a["dts"].append(dts)
                    run_pos += int(size.sum())
                    a["next_dts"] += int(dur.sum())
                prev_end = run_pos
    for tid, a in acc.items():
        if not a["sizes"]:
            continue
        t = by_id[tid]
        cat = lambda key, dt: np.concatenate(a[key]).astype(dt)          # noqa: E731
        t.sizes = np.concatenate((t.sizes, cat("sizes", np.uint64)))
        t.offsets = np.concatenate((t.offsets, cat("offsets", np.uint64)))
        base_dts = int(t.dts[-1] + t.deltas[-1]) if t.dts.size else 0
        frag_dts = cat("dts", np.int64)
        t.dts = np.concatenate((t.dts, frag_dts + (base_dts if frag_dts.size and frag_dts[0] == 0 and t.dts.size else 0)))
        t.deltas = np.concatenate((t.deltas, cat("deltas", np.int64)))
        new_cts = cat("cts", np.int64)
        if a["has_cts"] or t.cts_off is not None:
            old = t.cts_off if t.cts_off is not None else np.zeros(t.sync.size, np.int64)
            t.cts_off = np.concatenate((old, new_cts))
        t.sync = np.concatenate((t.sync, cat("sync", bool)))
        t.has_stss = t.has_stss or not bool(t.sync.all())


def read_movie(path: str | Path) -> Movie:
    """Parse the movie box of an ISO-BMFF file.  Raises BmffError when there is none."""
    path = Path(path)
    tops = top_level(path)
    moov_at = next(((p0, b1) for kind, b0, p0, b1 in tops if kind == b"moov"), None)
    if moov_at is None:
        raise BmffError("no moov box")
    ftyp = b""
    with open(path, "rb") as f:
        for kind, b0, p0, b1 in tops:
            if kind == b"ftyp" and b1 - b0 <= 4096:
                f.seek(b0)
                ftyp = f.read(b1 - b0)
                break
        f.seek(moov_at[0])
        moov = f.read(moov_at[1] - moov_at[0])
    mvhd = find_box(moov, 0, len(moov), b"mvhd")
    if mvhd is None:
        raise BmffError("no mvhd box")
    if moov[mvhd[0]] == 1:
        ts, dur = struct.unpack_from(">IQ", moov, mvhd[0] + 20)
    else:
        ts, dur = struct.unpack_from(">II", moov, mvhd[0] + 12)
        if dur == 0xFFFFFFFF:
            dur = 0
    movie = Movie(path, int(ts), int(dur), ftyp=ftyp)
    for kind, s, e in iter_boxes(moov, 0, len(moov)):
        if kind == b"trak":
            t = _parse_track(moov, s, e)
            if t is not None:
                movie.tracks.append(t)
    movie.fragmented = any(kind == b"moof" for kind, *_ in tops)
    if movie.fragmented:
        if not movie.duration:
            movie.fragment_duration_s = _fragments_duration(path, tops, moov, movie)
        _append_fragments(path, tops, moov, movie)
    return movie


# ---- cutting ----------------------------------------------------------------------------------------------------
@dataclass
class CutResult:
    first: int                      # reference-track sample range [first, last) that was copied (decode order)
    last: int
    first_accurate: int             # first sample whose presentation time is >= the requested start
    tracks: int                     # tracks written
    bytes_copied: int
    presentation_start: float       # seconds, on the source timeline, where the output's presentation begins


def select_reference_range(pres: np.ndarray, sync: np.ndarray, start: float, end: float, stream_copy: bool):
    """(first, last, first_accurate) over decode-order samples, or None when the window holds no sample."""
    s = float("%.3f" % start)
    d = float("%.3f" % (end - start))
    if d <= 0 or pres.size == 0:
        return None
    keep = np.nonzero((pres >= s) & (pres < s + d))[0]
    if keep.size == 0:
        return None
    first_acc, last = int(keep[0]), int(keep[-1]) + 1
    first = first_acc
    if stream_copy:
        k = np.nonzero(sync[:first_acc + 1])[0]
        if k.size:
            first = int(k[-1])
    return first, last, first_acc


def _runs(values: np.ndarray):
    """Run-length encode an int64 array -> (counts, values)."""
    if values.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    brk = np.nonzero(np.diff(values))[0] + 1
    starts = np.concatenate(([0], brk))
    counts = np.diff(np.concatenate((starts, [values.size])))
    return counts.astype(np.int64), values[starts].astype(np.int64)


def _patch_duration(payload: bytes, at_v0: int, at_v1: int, value: int) -> bytes:
    """Return a header payload (tkhd/mdhd/mvhd, incl. version+flags) with its duration field replaced."""
    b = bytearray(payload)
    if b[0] == 1:
        struct.pack_into(">Q", b, at_v1, value)
    else:
        struct.pack_into(">I", b, at_v0, min(value, 0xFFFFFFFF))
    return bytes(b)


def cut_movie(movie: Movie, start: float, end: float, dst: str | Path, *, stream_copy: bool = True,
              accurate_presentation: bool = False, selection: tuple | None = None,
              first_sample: bytes | None = None, chunk_seconds: float = 0.5,
              chunk_bytes: int = 4 << 20, mapped: bool = False) -> CutResult | None:
    """Write the samples of [start, end) of every track of `movie` into a new faststart MP4 at `dst`.

    stream_copy            : extend the reference track back to its last sync sample (what `-c copy` does).
    accurate_presentation  : keep the keyframe lead-in in the file but start the PRESENTATION at the first sample
                             at or after `start` through the edit list (frame-accurate for any player that honours
                             edit lists, without re-encoding).
    selection              : (first, last, first_accurate) of the reference track when the caller already chose the
                             range (select_reference_range on its own time table).
    first_sample           : replacement bytes for the first copied sample of the reference track, which is then
                             marked as a sync sample (the PCM-intra frame-accurate path: a picture that repeats its
                             IDR is re-expressed by that IDR's sample).
    Returns None when the window selects nothing; raises BmffError/OSError on malformed input.
    """
    ref = movie.video_track() or next((t for t in movie.tracks if t.n), None)
    if ref is None:
        return None
    mts = movie.timescale or 1000
    pres = ref.pres_times(mts)
    if selection is not None:
        first, last, first_acc = (int(v) for v in selection)
        if not (0 <= first <= first_acc < last <= ref.n):
            raise BmffError("selection outside the reference track")
    else:
        sel = select_reference_range(pres, ref.sync, start, end, stream_copy)
        if sel is None:
            return None
        first, last, first_acc = sel
    span = pres[first:last]
    t_lo = float(span.min())
    dur_ref = ref.deltas[first:last].astype(np.float64) / ref.timescale
    t_hi = float((span + dur_ref).max())
    t_present = float(pres[first_acc]) if accurate_presentation else t_lo

    plans = []
    for t in movie.tracks:
        if not t.n:
            continue
        if t is ref:
            a, b = first, last
        else:
            p = t.pres_times(mts)
            e = p + t.deltas.astype(np.float64) / t.timescale
            idx = np.nonzero((e > t_lo) & (p < t_hi))[0]
            if idx.size == 0:
                continue
            a, b = int(idx[0]), int(idx[-1]) + 1
            if t.has_stss:                      # non-reference tracks with sync tables also start on a sync sample
                k = np.nonzero(t.sync[:a + 1])[0]
                if k.size:
                    a = int(k[-1])
            else:
                # transform-coded audio needs the packets just before the first audible one to decode it correctly
                # (AAC: one frame of overlap; Opus: 80 ms): keep them in the file, the edit list below skips them
                a = max(0, a - AUDIO_PREROLL_SAMPLES.get(t.codec, 0))
        sizes = t.sizes[a:b].copy()
        override = None
        if t is ref and first_sample is not None:
            override = first_sample
            sizes[0] = len(override)
        deltas = t.deltas[a:b]
        cts_off = None if t.cts_off is None else t.cts_off[a:b]
        sync = t.sync[a:b].copy()
        if override is not None:
            sync[0] = True
        # edit list of the cut: presentation begins at t_present on the source timeline
        _empty_s, mt_src = t.edit_shift(mts)
        tp = t.pres_times(mts)[a:b]
        cts_rel = (t.dts[a:b] - t.dts[a]) + (cts_off if cts_off is not None else 0)   # media time, new origin
        first_pres = float(tp.min())
        min_cts = int(np.min(cts_rel))
        if first_pres >= t_present:
            empty_ticks = int(round((first_pres - t_present) * mts))
            media_time = min_cts
        else:                                    # this track starts before the presentation start: skip into it
            empty_ticks = 0
            media_time = min_cts + int(round((t_present - first_pres) * t.timescale))
        media_dur = int(deltas.sum())
        pres_ticks = max(0, (media_dur - (media_time - min_cts)) * mts + t.timescale - 1) // t.timescale
        # chunking: consecutive samples up to chunk_seconds / chunk_bytes
        rel = (t.dts[a:b] - t.dts[a]).astype(np.float64) / t.timescale
        desc = None if t.desc is None else t.desc[a:b]
        chunk_first, chunk_count, chunk_bytes_ = _chunk_plan(rel, sizes, chunk_seconds, chunk_bytes, desc)
        chunk_time = tp[chunk_first] if cts_off is None else (t.dts[a:b][chunk_first] - mt_src) / t.timescale
        plans.append({"t": t, "a": a, "b": b, "sizes": sizes, "deltas": deltas, "cts_off": cts_off, "sync": sync,
                      "chunk_desc": None if desc is None else desc[chunk_first],
                      "override": override, "empty_ticks": empty_ticks, "media_time": media_time,
                      "media_dur": media_dur, "pres_ticks": pres_ticks + empty_ticks, "chunk_first": chunk_first,
                      "chunk_count": chunk_count, "chunk_bytes": chunk_bytes_,
                      "chunk_time": np.asarray(chunk_time, np.float64)})
    if not plans:
        return None
    total_bytes = write_plans(dst, plans, mts, movie.ftyp, movie.path, mapped=mapped)
    return CutResult(first, last, first_acc, len(plans), total_bytes, t_present)


def _chunk_plan(rel_seconds: np.ndarray, sizes: np.ndarray, chunk_seconds: float, chunk_bytes: int, desc=None):
    """Group consecutive samples into chunks of at most chunk_seconds / chunk_bytes -> (first, count, bytes) per chunk.
    A chunk has ONE sample description (stsc), so a change of `desc` also starts a new chunk."""
    n = sizes.size
    bucket_t = np.floor(rel_seconds / chunk_seconds).astype(np.int64)
    csum = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    bucket_b = csum // chunk_bytes
    key = bucket_t * (1 << 20) + (bucket_b - bucket_b[np.searchsorted(bucket_t, bucket_t, side="left")])
    change = np.diff(key) != 0
    if desc is not None:
        change |= np.diff(desc) != 0
    brk = np.nonzero(change)[0] + 1
    chunk_first = np.concatenate(([0], brk)).astype(np.int64)
    chunk_count = np.diff(np.concatenate((chunk_first, [n]))).astype(np.int64)
    chunk_bytes_ = np.add.reduceat(sizes.astype(np.int64), chunk_first)
    return chunk_first, chunk_count, chunk_bytes_


def plan_whole_track(t: Track, movie_timescale: int, chunk_seconds: float = 0.5, chunk_bytes: int = 4 << 20) -> dict:
    """Plan that carries every sample of a source track unchanged (its own edit list semantics re-expressed)."""
    empty_s, mt = t.edit_shift(movie_timescale)
    media_dur = int(t.deltas.sum())
    rel = t.dts.astype(np.float64) / t.timescale
    cf, cc, cb = _chunk_plan(rel, t.sizes, chunk_seconds, chunk_bytes, t.desc)
    empty_ticks = int(round(empty_s * movie_timescale))
    pres_ticks = max(0, (media_dur - mt) * movie_timescale + t.timescale - 1) // t.timescale
    return {"t": t, "a": 0, "b": t.n, "sizes": t.sizes.copy(), "deltas": t.deltas, "cts_off": t.cts_off,
            "sync": t.sync.copy(), "override": None, "empty_ticks": empty_ticks, "media_time": mt,
            "media_dur": media_dur, "pres_ticks": pres_ticks + empty_ticks, "chunk_first": cf, "chunk_count": cc,
            "chunk_bytes": cb, "chunk_time": rel[cf] + empty_s,
            "chunk_desc": None if t.desc is None else t.desc[cf]}


def make_video_track(track_id: int, codec: bytes, width: int, height: int, timescale: int,
                     compressor: bytes = b"", extra_boxes: bytes = b"") -> Track:
    """Header boxes of a new video track (no samples yet): tkhd, mdhd, hdlr, vmhd + dinf, stsd with one
    VisualSampleEntry `codec`."""
    matrix = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)
    tkhd = struct.pack(">I", 3) + struct.pack(">IIIII", 0, 0, track_id, 0, 0) + bytes(8) + \
        struct.pack(">hhhH", 0, 0, 0, 0) + matrix + struct.pack(">II", width << 16, height << 16)
    mdhd = struct.pack(">I", 0) + struct.pack(">IIIIHH", 0, 0, timescale, 0, 0x55C4, 0)
    hdlr = struct.pack(">I", 0) + struct.pack(">I4s12x", 0, b"vide") + b"VideoHandler\x00"
    dinf = box(b"dinf", full_box(b"dref", 0, 0, struct.pack(">I", 1) + full_box(b"url ", 0, 1, b"")))
    name = compressor[:31]
    entry = struct.pack(">6xH", 1) + bytes(16) + struct.pack(">HH", width, height) + \
        struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + bytes([len(name)]) + name + bytes(31 - len(name)) + \
        struct.pack(">Hh", 0x18, -1) + extra_boxes
    stsd = full_box(b"stsd", 0, 0, struct.pack(">I", 1) + box(codec, entry))
    z64, z = np.zeros(0, np.uint64), np.zeros(0, np.int64)
    return Track(track_id, b"vide", codec, timescale, 0, width, height, tkhd, mdhd, hdlr,
                 full_box(b"vmhd", 0, 1, bytes(8)) + dinf, stsd, z64, z64, z, z, None, np.zeros(0, bool), False, [])


def plan_memory_track(t: Track, samples: list, deltas, sync=None, chunk_seconds: float = 0.5,
                      chunk_bytes: int = 4 << 20) -> dict:
    """Plan for a track whose samples are byte strings in memory (e.g. the reducer's Motion-JPEG pictures)."""
    n = len(samples)
    sizes = np.asarray([len(x) for x in samples], np.uint64)
    deltas = np.asarray(deltas, np.int64)
    if deltas.size != n:
        raise BmffError("one duration per sample expected")
    dts = np.concatenate(([0], np.cumsum(deltas)[:-1])).astype(np.int64) if n else np.zeros(0, np.int64)
    sync = np.ones(n, bool) if sync is None else np.asarray(sync, bool)
    rel = dts.astype(np.float64) / t.timescale
    cf, cc, cb = _chunk_plan(rel, sizes, chunk_seconds, chunk_bytes)
    return {"t": t, "a": 0, "b": n, "sizes": sizes, "deltas": deltas, "cts_off": None, "sync": sync, "override": None,
            "empty_ticks": 0, "media_time": 0, "media_dur": int(deltas.sum()), "pres_ticks": None, "chunk_first": cf,
            "chunk_count": cc, "chunk_bytes": cb, "chunk_time": rel[cf], "mem": samples, "all_sync": bool(sync.all())}


def write_plans(dst: str | Path, plans: list, mts: int, ftyp: bytes = b"", src_path: Path | None = None,
                mapped: bool = False) -> int:
    """Write the planned tracks into a faststart MP4: ftyp, moov, mdat with the tracks' chunks interleaved by time.
    Sample bytes come from src_path (plans made from a parsed file) or from memory (plan_memory_track).  Returns the
    media bytes written.  mapped=True writes through landing.acquire_mapped (a recycled mapping of the output, RAM-
    backed file systems only) and falls back to the in-kernel copy when that is not available."""
    for p in plans:
        if p["pres_ticks"] is None:
            p["pres_ticks"] = (p["media_dur"] * mts + p["t"].timescale - 1) // p["t"].timescale
    # interleave: chunks of all tracks ordered by time (stable by track order)
    order = []
    for ti, p in enumerate(plans):
        for ci in range(p["chunk_first"].size):
            order.append((float(p["chunk_time"][ci]), ti, ci))
    order.sort()
    total_bytes = int(sum(int(p["chunk_bytes"].sum()) for p in plans))

    def build_moov(base: int, wide: bool) -> bytes:
        pos = base
        offs = [np.zeros(p["chunk_first"].size, np.uint64) for p in plans]
        for _t, ti, ci in order:
            offs[ti][ci] = pos
            pos += int(plans[ti]["chunk_bytes"][ci])
        traks = b""
        movie_dur = 0
        for ti, p in enumerate(plans):
            t = p["t"]
            n = p["b"] - p["a"]
            unit = getattr(t, "unit", None)
            if unit is not None:                 # groups of equal samples: one stts run over the true sample count
                n = int(p["sizes"].sum()) // unit[0]
                c, v = np.asarray([n], np.int64), np.asarray([unit[1]], np.int64)
            else:
                c, v = _runs(p["deltas"])
            stts = full_box(b"stts", 0, 0, struct.pack(">I", c.size) +
                            np.stack([c, v], 1).astype(">u4").tobytes())
            ctts = b""
            if p["cts_off"] is not None:
                c, v = _runs(p["cts_off"])
                ver = 1 if (v < 0).any() else 0
                ctts = full_box(b"ctts", ver, 0, struct.pack(">I", c.size) +
                                np.stack([c, v & 0xFFFFFFFF], 1).astype(">u4").tobytes())
            stss = b""
            if (t.has_stss and not p.get("all_sync")) or not p["sync"].all():
                k = np.nonzero(p["sync"])[0] + 1
                stss = full_box(b"stss", 0, 0, struct.pack(">I", k.size) + k.astype(">u4").tobytes())
            per = p["chunk_count"] if unit is None else p["chunk_bytes"] // unit[0]
            cd = p.get("chunk_desc")
            if cd is None:
                c, v = _runs(per)
                dv = np.ones_like(v)
            else:                                 # runs of (samples per chunk, sample description index)
                c, pair = _runs(per.astype(np.int64) * (1 << 20) + cd.astype(np.int64))
                v, dv = pair >> 20, pair & ((1 << 20) - 1)
            firsts = np.concatenate(([0], np.cumsum(c)[:-1])) + 1
            stsc = full_box(b"stsc", 0, 0, struct.pack(">I", c.size) +
                            np.stack([firsts, v, dv], 1).astype(">u4").tobytes())
            sz = p["sizes"]
            if unit is not None:
                stsz = full_box(b"stsz", 0, 0, struct.pack(">II", unit[0], n))
            elif n and (sz == sz[0]).all():
                stsz = full_box(b"stsz", 0, 0, struct.pack(">II", int(sz[0]), n))
            else:
                stsz = full_box(b"stsz", 0, 0, struct.pack(">II", 0, n) + sz.astype(">u4").tobytes())
            if wide:
                stco = full_box(b"co64", 0, 0, struct.pack(">I", offs[ti].size) + offs[ti].astype(">u8").tobytes())
            else:
                stco = full_box(b"stco", 0, 0, struct.pack(">I", offs[ti].size) + offs[ti].astype(">u4").tobytes())
            stbl = box(b"stbl", t.stsd + stts + ctts + stss + stsc + stsz + stco)
            minf = box(b"minf", t.minf_other + stbl)
            mdhd = box(b"mdhd", _patch_duration(t.mdhd, 16, 24, p["media_dur"]))
            mdia = box(b"mdia", mdhd + box(b"hdlr", t.hdlr) + minf)
            entries = b""
            n_e = 0
            shown = p["pres_ticks"] - p["empty_ticks"]
            if p["empty_ticks"] > 0:
                entries += struct.pack(">Qqi", p["empty_ticks"], -1, 0x10000)
                n_e += 1
            entries += struct.pack(">Qqi", shown, p["media_time"], 0x10000)
            n_e += 1
            edts = box(b"edts", full_box(b"elst", 1, 0, struct.pack(">I", n_e) + entries))
            tkhd = box(b"tkhd", _patch_duration(t.tkhd, 20, 28, p["pres_ticks"]))
            traks += box(b"trak", tkhd + edts + mdia)
            movie_dur = max(movie_dur, p["pres_ticks"])
        next_id = max(p["t"].track_id for p in plans) + 1
        mvhd = full_box(b"mvhd", 0, 0, struct.pack(">IIIIIH", 0, 0, mts, min(movie_dur, 0xFFFFFFFF), 0x10000, 0x0100) +
                        bytes(10) + struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000) + bytes(24) +
                        struct.pack(">I", next_id))
        return box(b"moov", mvhd + traks)

    ftyp = ftyp or box(b"ftyp", b"isom" + struct.pack(">I", 0x200) + b"isomiso2mp41")
    wide = False
    moov = build_moov(0, wide)
    base = len(ftyp) + len(moov) + 16
    if base + total_bytes >= (1 << 32) - 1:
        wide = True
        moov = build_moov(0, wide)
        base = len(ftyp) + len(moov) + 16
    moov = build_moov(base, wide)

    # source byte ranges in output order, adjacent ranges merged
    src_lo, src_len, over_at = [], [], {}
    for _t, ti, ci in order:
        p = plans[ti]
        t = p["t"]
        f0 = p["a"] + int(p["chunk_first"][ci])
        cnt = int(p["chunk_count"][ci])
        if "mem" in p:
            blob = b"".join(bytes(x) for x in p["mem"][f0:f0 + cnt])
            over_at[len(src_lo)] = blob
            src_lo.append(-1)
            src_len.append(len(blob))
            continue
        o = t.offsets[f0:f0 + cnt].astype(np.int64)
        z = t.sizes[f0:f0 + cnt].astype(np.int64)
        if p["override"] is not None and ci == 0:
            over_at[len(src_lo)] = p["override"]
            src_lo.append(-1)
            src_len.append(len(p["override"]))
            o, z = o[1:], z[1:]
            if o.size == 0:
                continue
        brk = np.nonzero(o[1:] != o[:-1] + z[:-1])[0] + 1
        st = np.concatenate(([0], brk))
        en = np.concatenate((brk, [o.size]))
        zc = np.concatenate(([0], np.cumsum(z)))
        for s_, e_ in zip(st, en):
            lo, ln = int(o[s_]), int(zc[e_] - zc[s_])
            if src_lo and src_lo[-1] >= 0 and src_lo[-1] + src_len[-1] == lo:
                src_len[-1] += ln
            else:
                src_lo.append(lo)
                src_len.append(ln)
    dst = Path(dst)
    need_src = any(lo >= 0 for lo in src_lo)
    if need_src and src_path is None:
        raise BmffError("plans refer to a source file but none was given")
    src_size = os.path.getsize(src_path) if need_src else 0
    head = ftyp + moov + struct.pack(">I4sQ", 1, b"mdat", 16 + total_bytes)

    def emit(out_fd: int) -> int:
        pos = _write_all(out_fd, head, 0)
        for i, (lo, ln) in enumerate(zip(src_lo, src_len)):
            if lo < 0:
                pos = _write_all(out_fd, over_at[i], pos)
                continue
            if lo + ln > src_size:
                raise BmffError("sample data past the end of the file (truncated source)")
            _copy_range(in_fd, out_fd, lo, pos, ln)
            pos += ln
        return pos

    out = None
    if mapped:
        for lo, ln in zip(src_lo, src_len):
            if lo >= 0 and lo + ln > src_size:
                raise BmffError("sample data past the end of the file (truncated source)")
        from . import landing
        out = landing.acquire_mapped(dst, len(head) + int(sum(src_len)))
    if out is not None and out.array is not None:         # recycled pages: plain stores
        try:
            view = _source_view(src_path) if need_src else None
            a = out.array
            a[:len(head)] = np.frombuffer(head, np.uint8)
            pos = len(head)
            jobs = []
            for i, (lo, ln) in enumerate(zip(src_lo, src_len)):
                if lo < 0:
                    a[pos:pos + ln] = np.frombuffer(over_at[i], np.uint8)
                elif ln < _PARALLEL_COPY_MIN:
                    a[pos:pos + ln] = view[lo:lo + ln]
                else:                             # hour-long segments are gigabytes: one thread moves 9-15 GB/s
                    step = -(-ln // _PARALLEL_COPY_THREADS) + 4095 & ~4095
                    for k in range(0, ln, step):
                        n = min(step, ln - k)
                        jobs.append(_copy_pool().submit(np.copyto, a[pos + k:pos + k + n], view[lo + k:lo + k + n]))
                pos += ln
            for j in jobs:
                j.result()
        except BaseException:
            out.abort()
            raise
        return total_bytes
    in_fd = os.open(src_path, os.O_RDONLY) if need_src else -1
    try:
        if out is not None:                                # a new arena file: in-kernel copy, then a warm mapping
            try:
                emit(out.fd)
            except BaseException:
                out.abort()
                raise
            out.populate()
            return total_bytes
        # An existing output is overwritten IN PLACE and trimmed at the end: rewriting pages a file already owns is
        # faster than allocating fresh ones (tmpfs on the GPU box: 5.1 vs 3.8 GB/s, tools/copy_probe.py), which
        # matters when a segment is re-cut (retries, `ffmpeg -y` semantics).
        out_fd = os.open(dst, os.O_RDWR | os.O_CREAT, 0o644)
        try:
            os.ftruncate(out_fd, emit(out_fd))
        finally:
            os.close(out_fd)
    finally:
        if in_fd >= 0:
            os.close(in_fd)
    return total_bytes


_SOURCE_VIEWS: dict = {}
_PARALLEL_COPY_MIN = 256 << 20
_PARALLEL_COPY_THREADS = 4
_COPY_POOL: list = []


def _copy_pool():
    if not _COPY_POOL:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL.append(ThreadPoolExecutor(max_workers=_PARALLEL_COPY_THREADS, thread_name_prefix="vt-mp4-copy"))
    return _COPY_POOL[0]


def _source_view(path: Path) -> np.ndarray:
    """Read-only mapping of a source file, kept for the segments that follow (two files at most)."""
    st = os.stat(path)
    key = (str(path), st.st_size, st.st_mtime_ns, st.st_ino)
    v = _SOURCE_VIEWS.get(key)
    if v is None:
        while len(_SOURCE_VIEWS) >= 2:
            _SOURCE_VIEWS.pop(next(iter(_SOURCE_VIEWS)))
        v = _SOURCE_VIEWS[key] = np.memmap(path, dtype=np.uint8, mode="r")
    return v


def _write_all(fd: int, data: bytes, offset: int) -> int:
    mv = memoryview(data)
    done = 0
    while done < len(mv):
        done += os.pwrite(fd, mv[done:], offset + done)
    return offset + done


_COPY_MODE = ["copy_file_range"]


def _copy_range(in_fd: int, out_fd: int, src_off: int, dst_off: int, length: int) -> None:
    """Copy bytes between files inside the kernel (copy_file_range, else sendfile), falling back to pread/pwrite."""
    done = 0
    while done < length:
        mode = _COPY_MODE[0]
        want = min(length - done, 1 << 30)
        try:
            if mode == "copy_file_range":
                k = os.copy_file_range(in_fd, out_fd, want, src_off + done, dst_off + done)
            elif mode == "sendfile":
                os.lseek(out_fd, dst_off + done, os.SEEK_SET)
                k = os.sendfile(out_fd, in_fd, src_off + done, want)
            else:
                buf = os.pread(in_fd, min(want, 8 << 20), src_off + done)
                k = os.pwrite(out_fd, buf, dst_off + done) if buf else 0
        except (OSError, AttributeError):
            if mode == "rw":
                raise
            _COPY_MODE[0] = "sendfile" if mode == "copy_file_range" else "rw"
            continue
        if k == 0:
            raise BmffError("unexpected end of source while copying samples")
        done += k

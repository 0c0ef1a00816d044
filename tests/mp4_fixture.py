"""Test-side MP4 writer, independent of the product's container code: builds files the product did not write.

Video: H.264 samples given as lists of NAL units (so a sample can carry AUD + SEI + slice), optional `ctts` offsets and
an edit list; audio: a second trak of 16-bit little-endian PCM (`sowt`), one PCM frame per sample, chunked.  Chunks of
the two tracks are interleaved in `mdat`, `moov` goes first or last, chunk offsets are `stco` or `co64`.
"""
import struct

import numpy as np


def box(kind, payload):
    return struct.pack(">I4s", 8 + len(payload), kind) + payload


def full(kind, version, flags, payload):
    return box(kind, struct.pack(">I", (version << 24) | flags) + payload)


MATRIX = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)


def _runs(values):
    out = []
    for v in values:
        if out and out[-1][1] == v:
            out[-1][0] += 1
        else:
            out.append([1, v])
    return out


def write_av_mp4(path, *, sps, pps, video_samples, keyframes, width, height, timescale=15360, delta=512,
                 ctts=None, video_media_time=0, audio_rate=48000, audio_channels=2, audio_pcm=None,
                 audio_chunk=1024, video_chunk=5, moov_first=False, co64=False, movie_timescale=1000,
                 audio_empty_edit=0, version1=False, stz2=False, audio_codec=b"sowt", second_entry_chunk=None):
    """video_samples: list of lists of NAL byte strings (no start codes / lengths).  audio_pcm: int16 array
    [n, channels] or None.  version1: 64-bit mvhd / tkhd / mdhd / elst; stz2: the video sizes as a compact (16-bit)
    `stz2` box.  second_entry_chunk: video chunks from this index on use a SECOND sample entry in `stsd` (same parameter
    sets, another compressor name), the way files assembled from several encodes do.  Returns a dict describing what
    was written (sample byte strings per track)."""
    def hdr_times(ts, dur):            # creation, modification, timescale, duration of mvhd / mdhd
        return struct.pack(">QQIQ", 0, 0, ts, dur) if version1 else struct.pack(">IIII", 0, 0, ts, dur)

    def tk_times(track_id, dur):       # creation, modification, track id, reserved, duration of tkhd
        return struct.pack(">QQIIQ", 0, 0, track_id, 0, dur) if version1 else struct.pack(">IIIII", 0, 0, track_id, 0, dur)

    def elst(entries):                 # [(segment duration, media time)]
        if version1:
            body = b"".join(struct.pack(">QqI", d, mt, 0x10000) for d, mt in entries)
        else:
            body = b"".join(struct.pack(">IiI", d, mt, 0x10000) for d, mt in entries)
        return full(b"elst", 1 if version1 else 0, 0, struct.pack(">I", len(entries)) + body)

    ver = 1 if version1 else 0
    vs = [b"".join(struct.pack(">I", len(n)) + n for n in nals) for nals in video_samples]
    n_v = len(vs)
    a_bytes = b""
    n_a = 0
    bpf = 2 * audio_channels
    if audio_pcm is not None:
        a_bytes = np.ascontiguousarray(audio_pcm, "<i2").tobytes()
        n_a = len(a_bytes) // bpf
    # chunk plan: video chunks of `video_chunk` samples, audio chunks of `audio_chunk` PCM frames, interleaved by time
    v_chunks = [(i, min(i + video_chunk, n_v)) for i in range(0, n_v, video_chunk)]
    a_chunks = [(i, min(i + audio_chunk, n_a)) for i in range(0, n_a, audio_chunk)]
    order = [(a * delta / timescale, 0, ci) for ci, (a, b) in enumerate(v_chunks)] + \
            [(a / audio_rate, 1, ci) for ci, (a, b) in enumerate(a_chunks)]
    order.sort()

    def moov(base):
        pos = base
        v_off, a_off = [0] * len(v_chunks), [0] * len(a_chunks)
        for _t, trk, ci in order:
            if trk == 0:
                v_off[ci] = pos
                pos += sum(len(s) for s in vs[v_chunks[ci][0]:v_chunks[ci][1]])
            else:
                a_off[ci] = pos
                pos += (a_chunks[ci][1] - a_chunks[ci][0]) * bpf
        v_media = n_v * delta
        v_movie = (v_media - video_media_time) * movie_timescale // timescale
        avcc = struct.pack(">BBBBBB", 1, sps[1], sps[2], sps[3], 0xFF, 0xE1) + struct.pack(">H", len(sps)) + sps + \
            struct.pack(">BH", 1, len(pps)) + pps
        avc1 = struct.pack(">6xH", 1) + bytes(16) + struct.pack(">HH", width, height) + \
            struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + bytes(32) + struct.pack(">Hh", 0x18, -1) + \
            box(b"avcC", avcc) + box(b"pasp", struct.pack(">II", 1, 1))
        if second_entry_chunk is None:
            stbl = full(b"stsd", 0, 0, struct.pack(">I", 1) + box(b"avc1", avc1))
        else:
            name = b"\x0csecond entry" + bytes(19)
            avc1_b = avc1[:42] + name + avc1[74:]
            stbl = full(b"stsd", 0, 0, struct.pack(">I", 2) + box(b"avc1", avc1) + box(b"avc1", avc1_b))
        stbl += full(b"stts", 0, 0, struct.pack(">III", 1, n_v, delta))
        if ctts is not None:
            r = _runs(list(ctts))
            stbl += full(b"ctts", 0, 0, struct.pack(">I", len(r)) + b"".join(struct.pack(">II", c, v) for c, v in r))
        sync = [i + 1 for i, k in enumerate(keyframes) if k]
        stbl += full(b"stss", 0, 0, struct.pack(">I", len(sync)) + struct.pack(">%dI" % len(sync), *sync))
        r = _runs([(b - a, 1 if second_entry_chunk is None or ci < second_entry_chunk else 2)
                   for ci, (a, b) in enumerate(v_chunks)])
        ent, first = b"", 1
        for c, (v, d) in r:
            ent += struct.pack(">III", first, v, d)
            first += c
        stbl += full(b"stsc", 0, 0, struct.pack(">I", len(r)) + ent)
        if stz2 and max(len(s) for s in vs) < 65536:
            stbl += full(b"stz2", 0, 0, struct.pack(">3xBI", 16, n_v) + b"".join(struct.pack(">H", len(s)) for s in vs))
        else:
            stbl += full(b"stsz", 0, 0, struct.pack(">II", 0, n_v) + b"".join(struct.pack(">I", len(s)) for s in vs))
        if co64:
            stbl += full(b"co64", 0, 0, struct.pack(">I", len(v_off)) + struct.pack(">%dQ" % len(v_off), *v_off))
        else:
            stbl += full(b"stco", 0, 0, struct.pack(">I", len(v_off)) + struct.pack(">%dI" % len(v_off), *v_off))
        dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
        minf = box(b"minf", full(b"vmhd", 0, 1, bytes(8)) + dinf + box(b"stbl", stbl))
        mdia = box(b"mdia", full(b"mdhd", ver, 0, hdr_times(timescale, v_media) + struct.pack(">HH", 0x55C4, 0)) +
                   full(b"hdlr", 0, 0, struct.pack(">I4s12x", 0, b"vide") + b"FixtureVideo\x00") + minf)
        edts = box(b"edts", elst([(v_movie, video_media_time)]))
        tkhd = full(b"tkhd", ver, 3, tk_times(1, v_movie) + bytes(8) +
                    struct.pack(">hhhH", 0, 0, 0, 0) + MATRIX + struct.pack(">II", width << 16, height << 16))
        traks = box(b"trak", tkhd + edts + mdia)
        movie_dur = v_movie
        if n_a:
            a_movie = n_a * movie_timescale // audio_rate
            sowt = struct.pack(">6xH", 1) + struct.pack(">HHIHHHHI", 0, 0, 0, audio_channels, 16, 0, 0,
                                                        audio_rate << 16)
            stbl = full(b"stsd", 0, 0, struct.pack(">I", 1) + box(audio_codec, sowt))
            stbl += full(b"stts", 0, 0, struct.pack(">III", 1, n_a, 1))
            r = _runs([b - a for a, b in a_chunks])
            ent, first = b"", 1
            for c, v in r:
                ent += struct.pack(">III", first, v, 1)
                first += c
            stbl += full(b"stsc", 0, 0, struct.pack(">I", len(r)) + ent)
            stbl += full(b"stsz", 0, 0, struct.pack(">II", bpf, n_a))
            if co64:
                stbl += full(b"co64", 0, 0, struct.pack(">I", len(a_off)) + struct.pack(">%dQ" % len(a_off), *a_off))
            else:
                stbl += full(b"stco", 0, 0, struct.pack(">I", len(a_off)) + struct.pack(">%dI" % len(a_off), *a_off))
            minf = box(b"minf", full(b"smhd", 0, 0, bytes(4)) + dinf + box(b"stbl", stbl))
            mdia = box(b"mdia", full(b"mdhd", ver, 0, hdr_times(audio_rate, n_a) + struct.pack(">HH", 0x55C4, 0)) +
                       full(b"hdlr", 0, 0, struct.pack(">I4s12x", 0, b"soun") + b"FixtureAudio\x00") + minf)
            edts = box(b"edts", elst(([(audio_empty_edit, -1)] if audio_empty_edit else []) + [(a_movie, 0)]))
            tkhd = full(b"tkhd", ver, 3, tk_times(2, a_movie + audio_empty_edit) + bytes(8) +
                        struct.pack(">hhhH", 0, 1, 0x0100, 0) + MATRIX + struct.pack(">II", 0, 0))
            traks += box(b"trak", tkhd + edts + mdia)
            movie_dur = max(movie_dur, a_movie + audio_empty_edit)
        mvhd = full(b"mvhd", ver, 0, hdr_times(movie_timescale, movie_dur) + struct.pack(">IH", 0x10000, 0x0100) +
                    bytes(10) + MATRIX + bytes(24) + struct.pack(">I", 3))
        return box(b"moov", mvhd + traks + box(b"udta", box(b"name", b"fixture")))

    ftyp = box(b"ftyp", b"isom" + struct.pack(">I", 0x200) + b"isomiso2avc1mp41")
    free = box(b"free", b"")
    mdat_payload = b""
    for _t, trk, ci in order:
        if trk == 0:
            mdat_payload += b"".join(vs[v_chunks[ci][0]:v_chunks[ci][1]])
        else:
            mdat_payload += a_bytes[a_chunks[ci][0] * bpf:a_chunks[ci][1] * bpf]
    if moov_first:
        m = moov(0)
        base = len(ftyp) + len(m) + len(free) + 8
        data = ftyp + moov(base) + free + box(b"mdat", mdat_payload)
    else:
        base = len(ftyp) + len(free) + 8
        data = ftyp + free + box(b"mdat", mdat_payload) + moov(base)
    with open(path, "wb") as f:
        f.write(data)
    return {"video_samples": vs, "audio_bytes": a_bytes, "audio_frames": n_a, "bytes_per_audio_frame": bpf}


def write_fragmented_mp4(path, *, n_frag=4, per_frag=25, timescale=12800, delta=512, with_mehd=False):
    """A minimal fragmented (moof) file with dummy 16-byte samples: only its timing is meaningful."""
    total = n_frag * per_frag * delta
    mp4v = struct.pack(">6xH", 1) + bytes(16) + struct.pack(">HH", 64, 48) + \
        struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + bytes(32) + struct.pack(">Hh", 0x18, -1)
    stbl = full(b"stsd", 0, 0, struct.pack(">I", 1) + box(b"mp4v", mp4v)) + full(b"stts", 0, 0, struct.pack(">I", 0)) + \
        full(b"stsc", 0, 0, struct.pack(">I", 0)) + full(b"stsz", 0, 0, struct.pack(">II", 0, 0)) + \
        full(b"stco", 0, 0, struct.pack(">I", 0))
    dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
    minf = box(b"minf", full(b"vmhd", 0, 1, bytes(8)) + dinf + box(b"stbl", stbl))
    mdia = box(b"mdia", full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, 0, 0x55C4, 0)) +
               full(b"hdlr", 0, 0, struct.pack(">I4s12x", 0, b"vide") + b"V\x00") + minf)
    tkhd = full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, 0) + bytes(8) + struct.pack(">hhhH", 0, 0, 0, 0) +
                MATRIX + struct.pack(">II", 64 << 16, 48 << 16))
    mvex = full(b"trex", 0, 0, struct.pack(">IIIII", 1, 1, delta, 16, 0))
    if with_mehd:
        mvex = full(b"mehd", 0, 0, struct.pack(">I", total * 1000 // timescale)) + mvex
    mvhd = full(b"mvhd", 0, 0, struct.pack(">IIIIIH", 0, 0, 1000, 0, 0x10000, 0x0100) + bytes(10) + MATRIX +
                bytes(24) + struct.pack(">I", 2))
    data = box(b"ftyp", b"iso5" + struct.pack(">I", 0x200) + b"iso5iso6mp41") + \
        box(b"moov", mvhd + box(b"trak", tkhd + mdia) + box(b"mvex", mvex))
    for k in range(n_frag):
        tfhd = full(b"tfhd", 0, 0x020000, struct.pack(">I", 1))
        tfdt = full(b"tfdt", 0, 0, struct.pack(">I", k * per_frag * delta))
        trun = full(b"trun", 0, 0x000001, struct.pack(">Ii", per_frag, 0))          # durations from trex
        if k == n_frag - 1:                                                          # explicit durations in the last
            trun = full(b"trun", 0, 0x000101, struct.pack(">Ii", per_frag, 0) + struct.pack(">I", delta) * per_frag)
        moof = box(b"moof", full(b"mfhd", 0, 0, struct.pack(">I", k + 1)) + box(b"traf", tfhd + tfdt + trun))
        data += moof + box(b"mdat", bytes(16 * per_frag))
    with open(path, "wb") as f:
        f.write(data)
    return total / timescale


def write_fragmented_av(path, *, sps, pps, video_samples, keyframes, width, height, timescale=30000, delta=1000,
                        per_fragment=12, base_is_moof=True, with_tfdt=True):
    """A fragmented MP4 (`moov` with empty tables + `mvex`, then `moof` + `mdat` per fragment) that carries real H.264
    samples: per-sample sizes and flags in `trun`, `tfdt`, default-base-is-moof or an explicit base_data_offset."""
    vs = [b"".join(struct.pack(">I", len(n)) + n for n in nals) for nals in video_samples]
    avcc = struct.pack(">BBBBBB", 1, sps[1], sps[2], sps[3], 0xFF, 0xE1) + struct.pack(">H", len(sps)) + sps + \
        struct.pack(">BH", 1, len(pps)) + pps
    avc1 = struct.pack(">6xH", 1) + bytes(16) + struct.pack(">HH", width, height) + \
        struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + bytes(32) + struct.pack(">Hh", 0x18, -1) + box(b"avcC", avcc)
    stbl = full(b"stsd", 0, 0, struct.pack(">I", 1) + box(b"avc1", avc1)) + full(b"stts", 0, 0, struct.pack(">I", 0)) + \
        full(b"stsc", 0, 0, struct.pack(">I", 0)) + full(b"stsz", 0, 0, struct.pack(">II", 0, 0)) + \
        full(b"stco", 0, 0, struct.pack(">I", 0))
    dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
    minf = box(b"minf", full(b"vmhd", 0, 1, bytes(8)) + dinf + box(b"stbl", stbl))
    mdia = box(b"mdia", full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, 0, 0x55C4, 0)) +
               full(b"hdlr", 0, 0, struct.pack(">I4s12x", 0, b"vide") + b"V\x00") + minf)
    tkhd = full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, 0) + bytes(8) + struct.pack(">hhhH", 0, 0, 0, 0) +
                MATRIX + struct.pack(">II", width << 16, height << 16))
    mvex = full(b"trex", 0, 0, struct.pack(">IIIII", 1, 1, delta, 0, 0x00010000))
    mvhd = full(b"mvhd", 0, 0, struct.pack(">IIIIIH", 0, 0, 1000, 0, 0x10000, 0x0100) + bytes(10) + MATRIX +
                bytes(24) + struct.pack(">I", 2))
    data = box(b"ftyp", b"iso5" + struct.pack(">I", 0x200) + b"iso5iso6mp41") + \
        box(b"moov", mvhd + box(b"trak", tkhd + mdia) + box(b"mvex", mvex))
    n = len(vs)
    seq = 0
    for f0 in range(0, n, per_fragment):
        f1 = min(n, f0 + per_fragment)
        seq += 1
        rows = b"".join(struct.pack(">II", len(vs[i]), 0x02000000 if keyframes[i] else 0x01010000) for i in range(f0, f1))

        def moof_bytes(data_offset, base):
            flags = 0x020000 if base_is_moof else 0x000001
            tfhd = full(b"tfhd", 0, flags, struct.pack(">I", 1) + (b"" if base_is_moof else struct.pack(">Q", base)))
            tfdt = full(b"tfdt", 1, 0, struct.pack(">Q", f0 * delta)) if with_tfdt else b""
            trun = full(b"trun", 0, 0x000601, struct.pack(">Ii", f1 - f0, data_offset) + rows)   # sizes + flags per sample
            return box(b"moof", full(b"mfhd", 0, 0, struct.pack(">I", seq)) + box(b"traf", tfhd + tfdt + trun))

        moof_len = len(moof_bytes(0, 0))
        here = len(data)
        moof = moof_bytes(moof_len + 8, here) if base_is_moof else moof_bytes(0, here + moof_len + 8)
        data += moof + box(b"mdat", b"".join(vs[f0:f1]))
    with open(path, "wb") as f:
        f.write(data)
    return {"video_samples": vs}


# ---- a minimal Matroska writer (EBML) for tests: one AVC video track and one Opus audio track ---------------------------
def _ebml(eid, payload):
    n = len(payload)
    for width in range(1, 9):
        if n < (1 << (7 * width)) - 1:
            size = ((1 << (7 * width)) | n).to_bytes(width, "big")
            break
    return eid + size + payload


def _u(eid, value, width=None):
    width = width or max(1, (value.bit_length() + 7) // 8)
    return _ebml(eid, value.to_bytes(width, "big"))


def write_mkv(path, *, sps, pps, video_samples, keyframes, width, height, fps=30, opus_packets=None, opus_ms=20,
              frames_per_cluster=10, audio_codec=None):
    """video_samples: lists of NAL byte strings (stored length-prefixed, CodecPrivate = avcC, as V_MPEG4/ISO/AVC);
    opus_packets: list of byte strings, one per opus_ms milliseconds (A_OPUS with an OpusHead CodecPrivate).  SimpleBlocks,
    1 ms timestamps, Duration filled in.  audio_codec = (codec id, CodecPrivate, sampling rate, channels) labels the
    audio packets as another codec (their bytes are opaque to a stream copy).  Returns the sample byte strings per track."""
    vs = [b"".join(struct.pack(">I", len(n)) + n for n in nals) for nals in video_samples]
    avcc = struct.pack(">BBBBBB", 1, sps[1], sps[2], sps[3], 0xFF, 0xE1) + struct.pack(">H", len(sps)) + sps + \
        struct.pack(">BH", 1, len(pps)) + pps
    head = _ebml(b"\x1a\x45\xdf\xa3", _u(b"\x42\x86", 1) + _u(b"\x42\xf7", 1) + _u(b"\x42\xf2", 4) + _u(b"\x42\xf3", 8) +
                 _ebml(b"\x42\x82", b"matroska") + _u(b"\x42\x87", 4) + _u(b"\x42\x85", 2))
    n = len(vs)
    dur_ms = n * 1000.0 / fps
    info = _ebml(b"\x15\x49\xa9\x66", _u(b"\x2a\xd7\xb1", 1000000) + _ebml(b"\x44\x89", struct.pack(">d", dur_ms)) +
                 _ebml(b"\x4d\x80", b"fixture") + _ebml(b"\x57\x41", b"fixture"))
    vtrack = _ebml(b"\xae", _u(b"\xd7", 1) + _u(b"\x73\xc5", 1) + _u(b"\x83", 1) + _ebml(b"\x86", b"V_MPEG4/ISO/AVC") +
                   _ebml(b"\x63\xa2", avcc) + _u(b"\x23\xe3\x83", int(1e9 / fps)) +
                   _ebml(b"\xe0", _u(b"\xb0", width) + _u(b"\xba", height)))
    tracks = vtrack
    opus_head = b""
    if opus_packets:
        opus_head = b"OpusHead" + bytes([1, 2]) + struct.pack("<HIh", 312, 48000, 0) + bytes([0])
        a_id, a_priv, a_rate, a_ch = audio_codec or (b"A_OPUS", opus_head, 48000.0, 2)
        tracks += _ebml(b"\xae", _u(b"\xd7", 2) + _u(b"\x73\xc5", 2) + _u(b"\x83", 2) + _ebml(b"\x86", a_id) +
                        (_ebml(b"\x63\xa2", a_priv) if a_priv else b"") +
                        _ebml(b"\xe1", _ebml(b"\xb5", struct.pack(">f", float(a_rate))) + _u(b"\x9f", a_ch)))
    tracks = _ebml(b"\x16\x54\xae\x6b", tracks)
    events = [(int(round(i * 1000.0 / fps)), 1, i) for i in range(n)]
    events += [(i * opus_ms, 2, i) for i in range(len(opus_packets or []))]
    events.sort()
    clusters = b""
    cur, cur_ts, in_cluster = b"", None, 0
    for ts, trk, i in events:
        if cur_ts is None or (trk == 1 and keyframes[i] and in_cluster >= frames_per_cluster) or ts - cur_ts > 30000:
            if cur_ts is not None:
                clusters += _ebml(b"\x1f\x43\xb6\x75", _u(b"\xe7", cur_ts) + cur)
            cur, cur_ts, in_cluster = b"", ts, 0
        payload = vs[i] if trk == 1 else opus_packets[i]
        flags = 0x80 if (trk == 2 or keyframes[i]) else 0
        cur += _ebml(b"\xa3", bytes([0x80 | trk]) + struct.pack(">h", ts - cur_ts) + bytes([flags]) + payload)
        in_cluster += trk == 1
    clusters += _ebml(b"\x1f\x43\xb6\x75", _u(b"\xe7", cur_ts) + cur)
    with open(path, "wb") as f:
        f.write(head + _ebml(b"\x18\x53\x80\x67", info + tracks + clusters))
    return {"video_samples": vs, "opus_packets": list(opus_packets or []), "opus_head": opus_head}


def write_flv(path, *, sps, pps, video_samples, keyframes, fps=30, aac_frames=None, aac_rate_index=3, cto_ms=None,
              with_duration=True):
    """FLV with AVC video (sequence header + one NALU tag per sample, composition offsets cto_ms) and AAC audio
    (AudioSpecificConfig tag + raw frames of 1024 samples at the rate of aac_rate_index, 3 = 48 kHz), tags interleaved by
    time, onMetaData first.  Returns the sample byte strings per stream and the AudioSpecificConfig."""
    vs = [b"".join(struct.pack(">I", len(n)) + n for n in nals) for nals in video_samples]
    avcc = struct.pack(">BBBBBB", 1, sps[1], sps[2], sps[3], 0xFF, 0xE1) + struct.pack(">H", len(sps)) + sps + \
        struct.pack(">BH", 1, len(pps)) + pps
    asc = struct.pack(">H", (2 << 11) | (aac_rate_index << 7) | (2 << 3))        # AAC-LC, rate index, stereo
    rate = (96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050)[aac_rate_index]

    def tag(kind, ts, data):
        body = bytes([kind]) + len(data).to_bytes(3, "big") + (ts & 0xFFFFFF).to_bytes(3, "big") + \
            bytes([(ts >> 24) & 0xFF]) + bytes(3) + data
        return body + struct.pack(">I", len(body))

    n = len(vs)
    dur = n / fps
    meta = b"\x02" + struct.pack(">H", 10) + b"onMetaData" + b"\x08" + struct.pack(">I", 1) + \
        (struct.pack(">H", 8) + b"duration" + b"\x00" + struct.pack(">d", dur) if with_duration
         else struct.pack(">H", 5) + b"width" + b"\x00" + struct.pack(">d", 64.0)) + b"\x00\x00\x09"
    out = b"FLV\x01\x05" + struct.pack(">I", 9) + struct.pack(">I", 0) + tag(18, 0, meta)
    out += tag(9, 0, bytes([0x17, 0]) + bytes(3) + avcc)
    events = [(int(round(i * 1000.0 / fps)), 0, i) for i in range(n)]
    if aac_frames:
        out += tag(8, 0, bytes([0xAF, 0]) + asc)
        events += [(int(round(i * 1024 * 1000.0 / rate)), 1, i) for i in range(len(aac_frames))]
    events.sort()
    for ts, kind, i in events:
        if kind == 0:
            cto = (cto_ms[i] if cto_ms else 0)
            out += tag(9, ts, bytes([0x17 if keyframes[i] else 0x27, 1]) + cto.to_bytes(3, "big", signed=True) + vs[i])
        else:
            out += tag(8, ts, bytes([0xAF, 1]) + aac_frames[i])
    with open(path, "wb") as f:
        f.write(out)
    return {"video_samples": vs, "aac_frames": list(aac_frames or []), "asc": asc, "avcc": avcc}

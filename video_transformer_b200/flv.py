"""FLV demux for the stream copy (`extract_segment`, /root/reference/src/utils/video_segmenter.py:118-137): the
downloader's `best[height<=N]` fallback (/root/reference/src/downloader/video_downloader.py:56) can deliver FLV, and
`ffmpeg -c copy OUT.mp4` re-wraps its H.264 / AAC / MP3 streams without touching a sample.  This reader indexes the tags
into an isobmff.Movie (the same structure the MP4 and Matroska readers produce), so isobmff.cut_movie treats the file
like any other source.

Layout (Adobe FLV v10.1, third party, not under /root/reference): 9-byte header, then tags
`type(1) size(3) timestamp(3) timestamp_ext(1) stream_id(3) | data | previous_tag_size(4)`.
Video data: `frame_type<<4 | codec_id`, codec 7 = AVC: `packet_type(1) composition_time(s24 ms)` then an
AVCDecoderConfigurationRecord (type 0) or length-prefixed NAL units (type 1).  Audio data: `format<<4 | rate | size |
channels`, format 10 = AAC: `packet_type(1)` then an AudioSpecificConfig (0) or a raw frame (1); format 2 = MP3.
"""
from __future__ import annotations

import struct
from pathlib import Path

MAX_TAGS = 1 << 25


def read_movie(path: str | Path):
    """Index an FLV file into an isobmff.Movie.  Raises isobmff.BmffError when the file is not FLV or carries no
    stream with an MP4 mapping (AVC video; AAC or MP3 audio)."""
    from . import isobmff, matroska
    path = Path(path)
    size = path.stat().st_size
    video, audio = [], []                 # (file offset, size, presentation ms, keyframe)
    avcc = asc = None
    a_fmt = a_rate = a_ch = None
    with open(path, "rb") as f:
        head = f.read(9)
        if len(head) < 9 or head[:3] != b"FLV":
            raise isobmff.BmffError("not an FLV file")
        pos = struct.unpack(">I", head[5:9])[0] + 4
        n_tags = 0
        while pos + 11 <= size:
            f.seek(pos)
            th = f.read(16)
            if len(th) < 11:
                break
            kind = th[0] & 0x1F
            dsize = int.from_bytes(th[1:4], "big")
            ts = (th[7] << 24) | int.from_bytes(th[4:7], "big")
            body = pos + 11
            if body + dsize > size:
                break                                     # truncated download: keep what is complete
            n_tags += 1
            if n_tags > MAX_TAGS:
                raise isobmff.BmffError("too many FLV tags")
            if kind == 9 and dsize >= 5 and len(th) >= 16 and (th[11] & 0x0F) == 7:
                ptype = th[12]
                cto = int.from_bytes(th[13:16], "big", signed=True)
                if ptype == 0 and avcc is None:
                    f.seek(body + 5)
                    avcc = f.read(dsize - 5)
                elif ptype == 1 and dsize > 5:
                    video.append((body + 5, dsize - 5, ts + cto, (th[11] >> 4) == 1))
            elif kind == 8 and dsize >= 2 and len(th) >= 13:
                fmt = th[11] >> 4
                if fmt == 10:
                    if th[12] == 0 and asc is None:
                        f.seek(body + 2)
                        asc = f.read(dsize - 2)
                    elif th[12] == 1 and dsize > 2:
                        audio.append((body + 2, dsize - 2, ts, True))
                    a_fmt = 10
                elif fmt == 2 and dsize > 1:
                    audio.append((body + 1, dsize - 1, ts, True))
                    a_fmt = 2
                if a_rate is None:
                    a_rate = (5512, 11025, 22050, 44100)[(th[11] >> 2) & 3]
                    a_ch = 2 if th[11] & 1 else 1
            pos = body + dsize + 4
    movie = isobmff.Movie(path, 1000, 0)
    next_id = 1
    if video and avcc and len(avcc) >= 7:
        w = h = 0
        try:
            from . import container
            sps_len = struct.unpack_from(">H", avcc, 6)[0]
            w, h = container._sps_size(avcc[8:8 + sps_len])
        except Exception:  # noqa: BLE001 - dimensions are informative (tkhd / sample entry); the copy does not need them
            w = h = 0
        entry = matroska._video_entry(b"avc1", w, h, isobmff.box(b"avcC", avcc))
        movie.tracks.append(matroska.build_track(next_id, True, b"avc1", entry, 1000, w, h, video))
        next_id += 1
    if audio and (a_fmt == 2 or (a_fmt == 10 and asc)):
        if a_fmt == 10:
            rate, ch = _asc_rate_channels(asc, a_rate or 44100, a_ch or 2)
            entry = matroska._audio_entry(b"mp4a", ch, rate, matroska._esds(0x40, 5, asc))
        else:
            entry = matroska._audio_entry(b"mp4a", a_ch or 2, a_rate or 44100, matroska._esds(0x6B, 5, b""))
        movie.tracks.append(matroska.build_track(next_id, False, b"mp4a", entry, 1000, 0, 0, audio))
        next_id += 1
    if not movie.tracks:
        raise isobmff.BmffError("no FLV stream with an MP4 mapping")
    movie.duration = max(int(t.dts[-1] + t.deltas[-1]) + int(t.edits[0][0] if t.edits and t.edits[0][1] < 0 else 0)
                         for t in movie.tracks)
    return movie


def _asc_rate_channels(asc: bytes, rate: int, channels: int):
    """Sampling rate and channel count of an AudioSpecificConfig (FLV's own header always says 44.1 kHz stereo for AAC)."""
    if len(asc) < 2:
        return rate, channels
    bits = int.from_bytes(asc[:5].ljust(5, b"\0"), "big")
    sf = (bits >> 31) & 0xF                       # 5 bits object type, 4 bits frequency index
    table = (96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350)
    if sf == 15:
        rate = (bits >> 7) & 0xFFFFFF
        ch = (bits >> 3) & 0xF
    else:
        rate = table[sf] if sf < len(table) else rate
        ch = (bits >> 27) & 0xF
    return (rate or 44100), (ch if 0 < ch <= 8 else channels)

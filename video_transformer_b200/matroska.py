"""Matroska / WebM duration without ffprobe (EBML walk; host-side byte I/O).

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

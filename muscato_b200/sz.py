"""Snappy *framing format* reader/writer for the reference's ".sz" files
(golang/snappy NewReader / NewBufferedWriter).

Stream = "\\xff\\x06\\x00\\x00sNaPpY" then chunks: type(1) len(3, LE) [masked CRC32C(4) payload].
type 0x00 = Snappy-compressed block, 0x01 = stored block, 0x80-0xfd skippable, 0xfe padding.
The writer emits stored chunks (valid for every conforming reader, incl. golang/snappy and
sztool); the reader handles both kinds.  Container format only: content is plain text.
"""
from __future__ import annotations

import struct
from typing import Iterator

import numpy as np

_MAGIC = b"\xff\x06\x00\x00sNaPpY"
_MAX_BLOCK = 65536


def _make_crc_table() -> np.ndarray:
    poly = 0x82F63B78  # CRC-32C (Castagnoli), reflected
    tab = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab[i] = c
    return tab


_CRC_TABLE = _make_crc_table()
_CRC_LIST = [int(x) for x in _CRC_TABLE]


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    tab = _CRC_LIST
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def snappy_block_decode(buf: bytes) -> bytes:
    """Raw Snappy block format (varint length, then literal / copy elements)."""
    n = len(buf)
    i = 0
    ulen = 0
    shift = 0
    while True:
        b = buf[i]
        i += 1
        ulen |= (b & 0x7F) << shift
        if b < 0x80:
            break
        shift += 7
    out = bytearray()
    while i < n:
        tag = buf[i]
        i += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[i:i + nb], "little")
                i += nb
            ln += 1
            out += buf[i:i + ln]
            i += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[i]
            i += 1
        elif kind == 2:
            ln = 1 + (tag >> 2)
            off = buf[i] | (buf[i + 1] << 8)
            i += 2
        else:
            ln = 1 + (tag >> 2)
            off = int.from_bytes(buf[i:i + 4], "little")
            i += 4
        if off == 0 or off > len(out):
            raise ValueError("snappy: bad copy offset")
        start = len(out) - off
        if off >= ln:
            out += out[start:start + ln]
        else:
            for k in range(ln):
                out.append(out[start + k])
    if len(out) != ulen:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


def iter_chunks(raw: bytes, verify: bool = True) -> Iterator[bytes]:
    i = 0
    n = len(raw)
    seen_magic = False
    while i < n:
        if i + 4 > n:
            raise ValueError("sz: truncated chunk header")
        ctype = raw[i]
        clen = raw[i + 1] | (raw[i + 2] << 8) | (raw[i + 3] << 16)
        body = raw[i + 4:i + 4 + clen]
        if len(body) != clen:
            raise ValueError("sz: truncated chunk")
        i += 4 + clen
        if ctype == 0xFF:
            if body != _MAGIC[4:]:
                raise ValueError("sz: bad stream identifier")
            seen_magic = True
            continue
        if not seen_magic:
            raise ValueError("sz: missing stream identifier")
        if ctype in (0x00, 0x01):
            (crc,) = struct.unpack("<I", body[:4])
            data = body[4:]
            if ctype == 0x00:
                data = snappy_block_decode(data)
            if verify and masked_crc32c(data) != crc:
                raise ValueError("sz: CRC mismatch")
            yield data
        elif 0x02 <= ctype <= 0x7F:
            raise ValueError("sz: reserved unskippable chunk")
        # 0x80..0xfe: skippable / padding


def decompress(raw: bytes, verify: bool = True) -> bytes:
    return b"".join(iter_chunks(raw, verify))


def compress(data: bytes) -> bytes:
    out = [_MAGIC]
    for i in range(0, len(data), _MAX_BLOCK):
        blk = data[i:i + _MAX_BLOCK]
        out.append(b"\x01" + struct.pack("<I", len(blk) + 4)[:3] + struct.pack("<I", masked_crc32c(blk)) + blk)
    return b"".join(out)


def read_file(path: str) -> bytes:
    """Read a text file, transparently un-framing ".sz"."""
    with open(path, "rb") as f:
        raw = f.read()
    if path.endswith(".sz") or raw[:10] == _MAGIC:
        return decompress(raw)
    return raw


def write_file(path: str, data: bytes) -> None:
    with open(path, "wb") as f:
        f.write(compress(data) if path.endswith(".sz") else data)

"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the LZ4 frame format (decoder + a small encoder).

The reference reads ``fdt_matrix2.dot.lz4`` with ``lz4.frame.open(matfile, "rb")``
(NFACT/base/matrix_handling.py:108-110) and writes it with ``lz4.frame.open(..., "wb")``
(NFACT/config/nfact_config_functions.py:601-604).  ``lz4`` (python-lz4, unpinned at pyproject.toml:27) is a
third-party dependency that is NOT installed in this image, so neither the reference's reader nor a file written
by it can be produced here: **parity unpinned** for this function.  What anchors it instead:
  * the published LZ4 Frame Format (v1.6.x) / LZ4 Block Format / XXH32 specifications restated below;
  * byte-level known answers of the format that every LZ4 implementation emits: the ``lz4`` command line's
    frame headers ``04 22 4D 18 64 40 A7`` (64 KB blocks) and ``04 22 4D 18 64 70 B9`` (4 MB blocks), whose last
    byte is the XXH32-derived header checksum, the empty frame ``04 22 4D 18 64 40 A7 00 00 00 00 05 5D CC 02``,
    and XXH32's published test values (``tests/test_lz4_cpu.py``);
  * round trips of this module's encoder through the native decoder and through this module's decoder.
Only ``tests/`` may import this module.
"""
from __future__ import annotations

import struct

MAGIC = 0x184D2204
_P1, _P2, _P3, _P4, _P5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393
_M = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M


def xxh32(data: bytes, seed: int = 0) -> int:
    """XXH32 as specified by the xxHash project (used for every checksum of an LZ4 frame)."""
    n = len(data)
    i = 0
    if n >= 16:
        v = [(seed + _P1 + _P2) & _M, (seed + _P2) & _M, seed & _M, (seed - _P1) & _M]
        while i + 16 <= n:
            lanes = struct.unpack_from("<4I", data, i)
            for a in range(4):
                v[a] = (_rotl((v[a] + lanes[a] * _P2) & _M, 13) * _P1) & _M
            i += 16
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & _M
    else:
        h = (seed + _P5) & _M
    h = (h + n) & _M
    while i + 4 <= n:
        h = (_rotl((h + struct.unpack_from("<I", data, i)[0] * _P3) & _M, 17) * _P4) & _M
        i += 4
    while i < n:
        h = (_rotl((h + data[i] * _P5) & _M, 11) * _P1) & _M
        i += 1
    h ^= h >> 15
    h = (h * _P2) & _M
    h ^= h >> 13
    h = (h * _P3) & _M
    h ^= h >> 16
    return h


def decode_block(src: bytes, history: bytes = b"") -> bytes:
    """One LZ4 block (sequence stream) -> bytes; ``history`` is what linked blocks may reference."""
    out = bytearray(history)
    base = len(history)
    i, n = 0, len(src)
    if n == 0:
        raise ValueError("empty block")
    while True:
        token = src[i]
        i += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        if i + lit > n:
            raise ValueError("literals run past the block")
        out += src[i:i + lit]
        i += lit
        if i == n:
            break
        off = src[i] | (src[i + 1] << 8)
        i += 2
        ml = token & 15
        if ml == 15:
            while True:
                b = src[i]
                i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        if off == 0 or off > len(out):
            raise ValueError("match offset outside the window")
        start = len(out) - off
        for j in range(ml):          # byte by byte: matches may overlap their own output
            out.append(out[start + j])
    return bytes(out[base:])


def decompress(data: bytes) -> bytes:
    """All concatenated frames of ``data`` -> content (what ``lz4.frame.open(f, 'rb').read()`` returns)."""
    out = bytearray()
    p, n = 0, len(data)
    while p < n:
        magic, = struct.unpack_from("<I", data, p)
        if 0x184D2A50 <= magic <= 0x184D2A5F:
            size, = struct.unpack_from("<I", data, p + 4)
            p += 8 + size
            continue
        if magic != MAGIC:
            raise ValueError("not an LZ4 frame")
        p += 4
        d0 = p
        flg, bd = data[p], data[p + 1]
        p += 2
        if flg >> 6 != 1:
            raise ValueError("frame version")
        independent, block_sum = bool(flg & 0x20), bool(flg & 0x10)
        has_size, content_sum, dict_id = bool(flg & 0x08), bool(flg & 0x04), bool(flg & 0x01)
        block_max = 1 << (8 + 2 * ((bd >> 4) & 7))
        content_size = None
        if has_size:
            content_size, = struct.unpack_from("<Q", data, p)
            p += 8
        if dict_id:
            p += 4
        if data[p] != (xxh32(data[d0:p]) >> 8) & 0xFF:
            raise ValueError("header checksum")
        p += 1
        frame = bytearray()
        while True:
            word, = struct.unpack_from("<I", data, p)
            p += 4
            if word == 0:
                break
            size = word & 0x7FFFFFFF
            if size > block_max:
                raise ValueError("block larger than the frame maximum")
            blk = data[p:p + size]
            if len(blk) != size:
                raise ValueError("truncated block")
            p += size
            if block_sum:
                if struct.unpack_from("<I", data, p)[0] != xxh32(blk):
                    raise ValueError("block checksum")
                p += 4
            if word >> 31:
                frame += blk
            else:
                frame += decode_block(blk, b"" if independent else bytes(frame[-65536:]))
        if content_sum:
            if struct.unpack_from("<I", data, p)[0] != xxh32(bytes(frame)):
                raise ValueError("content checksum")
            p += 4
        if content_size is not None and content_size != len(frame):
            raise ValueError("content size")
        out += frame
    return bytes(out)


# ---- encoder (greedy hash-chain-free matcher; only to manufacture test inputs) ---------------------------

def _emit(out: bytearray, literals: bytes, match_len: int, offset: int) -> None:
    lit = len(literals)
    ml = match_len - 4 if match_len else 0
    out.append((min(lit, 15) << 4) | (min(ml, 15) if match_len else 0))
    if lit >= 15:
        rest = lit - 15
        while rest >= 255:
            out.append(255)
            rest -= 255
        out.append(rest)
    out += literals
    if match_len:
        out += struct.pack("<H", offset)
        if ml >= 15:
            rest = ml - 15
            while rest >= 255:
                out.append(255)
                rest -= 255
            out.append(rest)


def encode_block(chunk: bytes, history: bytes = b"") -> bytes:
    """Greedy LZ4 block encoder obeying the format's end-of-block rules (last 5 bytes literals, last match
    starts >= 12 bytes before the end)."""
    buf = history + chunk
    base = len(history)
    n = len(buf)
    out = bytearray()
    table = {}
    for j in range(max(0, base - 65535), max(base - 3, 0)):
        table[buf[j:j + 4]] = j
    i = anchor = base
    limit = n - 12
    while i < limit:
        key = buf[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            ml = 4
            while i + ml < n - 5 and buf[cand + ml] == buf[i + ml]:
                ml += 1
            _emit(out, buf[anchor:i], ml, i - cand)
            for j in range(i + 1, min(i + ml, limit)):
                table[buf[j:j + 4]] = j
            i += ml
            anchor = i
        else:
            i += 1
    _emit(out, buf[anchor:n], 0, 0)
    return bytes(out)


def compress(data: bytes, block_size_id: int = 4, independent: bool = True, block_checksum: bool = False,
             content_checksum: bool = True, content_size: bool = False, store_incompressible: bool = True) -> bytes:
    """``data`` -> one LZ4 frame.  python-lz4's ``lz4.frame.open(..., 'wb')`` defaults correspond to
    ``independent=False, content_checksum=False, block_size_id=4``; the ``lz4`` command line to
    ``independent=True, content_checksum=True, block_size_id=7``."""
    block_max = 1 << (8 + 2 * block_size_id)
    flg = (1 << 6) | (0x20 if independent else 0) | (0x10 if block_checksum else 0) | \
        (0x08 if content_size else 0) | (0x04 if content_checksum else 0)
    desc = bytes([flg, block_size_id << 4]) + (struct.pack("<Q", len(data)) if content_size else b"")
    out = bytearray(struct.pack("<I", MAGIC) + desc + bytes([(xxh32(desc) >> 8) & 0xFF]))
    for start in range(0, len(data), block_max):
        chunk = data[start:start + block_max]
        hist = b"" if independent else data[max(0, start - 65536):start]
        enc = encode_block(chunk, hist)
        if len(enc) > block_max or (store_incompressible and len(enc) >= len(chunk)):
            out += struct.pack("<I", len(chunk) | 0x80000000) + chunk
            enc = chunk
        else:
            out += struct.pack("<I", len(enc)) + enc
        if block_checksum:
            out += struct.pack("<I", xxh32(enc))
    out += struct.pack("<I", 0)
    if content_checksum:
        out += struct.pack("<I", xxh32(data))
    return bytes(out)

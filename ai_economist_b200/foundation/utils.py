"""Episode-log files in the reference's wire format (ai_economist/foundation/utils.py:19-43): the JSON text of
`env.previous_episode_dense_log` inside an LZ4 *frame* (what `lz4.frame.open` reads and writes).

The `lz4` package is not part of this image, so the frame container is implemented here: the writer emits standard
LZ4 frames whose blocks are stored uncompressed (a valid encoding every LZ4 reader accepts; `compression_level` is
accepted for signature compatibility), the reader handles stored and LZ4-compressed blocks, independent or LINKED
(lz4.frame's default, which the reference's writer uses: matches reach into the previous blocks), i.e. it also loads
logs written by the reference.  When `lz4` is importable it is used instead.
"""
import json
import struct

_MAGIC = 0x184D2204
_P1, _P2, _P3, _P4, _P5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393
_M = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M


def xxh32(data, seed=0):
    """xxHash32 (the checksum of the LZ4 frame descriptor and content)."""
    n, i = len(data), 0
    if n >= 16:
        v = [(seed + _P1 + _P2) & _M, (seed + _P2) & _M, seed & _M, (seed - _P1) & _M]
        while i <= n - 16:
            for j in range(4):
                w = struct.unpack_from("<I", data, i + 4 * j)[0]
                v[j] = (_rotl((v[j] + w * _P2) & _M, 13) * _P1) & _M
            i += 16
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & _M
    else:
        h = (seed + _P5) & _M
    h = (h + n) & _M
    while i <= n - 4:
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


def lz4_block_decompress(src, history=b""):
    """One LZ4 block (sequences of literals + back-references) -> bytes.  `history` is the output that precedes the
    block in a frame with LINKED blocks (the lz4.frame default, which the reference's writer uses): matches may reach
    back up to 64 KB into it."""
    out, i, n = bytearray(history), 0, len(src)
    base = len(out)
    while i < n:
        tok = src[i]; i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = src[i]; i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]; i += lit
        if i >= n:
            break
        off = src[i] | (src[i + 1] << 8); i += 2
        if off == 0 or off > len(out):
            raise ValueError("corrupt LZ4 block: bad offset")
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[i]; i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        if off >= ml:
            out += out[start:start + ml]
        else:
            for k in range(ml):  # overlaps its own output (run-length style match)
                out.append(out[start + k])
    return bytes(out[base:])


def lz4_block_compress(data, history=b""):
    """Greedy LZ4 block compressor (4-byte hash chain of length 1).  With `history` the matches may point into the
    previous blocks' output (linked blocks).  Small and slow: used for tests and on request, not by default."""
    buf = bytes(history) + bytes(data)
    base, n = len(history), len(history) + len(data)
    out, table = bytearray(), {}
    for k in range(max(0, base - 65535), base - 3):
        table[buf[k:k + 4]] = k
    anchor = i = base

    def emit(lit_end, match_len, offset):
        lit = lit_end - anchor
        tok_l, tok_m = min(lit, 15), (min(match_len - 4, 15) if match_len else 0)
        out.append((tok_l << 4) | tok_m)
        if lit >= 15:
            r = lit - 15
            while r >= 255:
                out.append(255); r -= 255
            out.append(r)
        out.extend(buf[anchor:lit_end])
        if match_len:
            out.extend((offset & 0xFF, offset >> 8))
            if match_len - 4 >= 15:
                r = match_len - 4 - 15
                while r >= 255:
                    out.append(255); r -= 255
                out.append(r)

    limit = n - 12   # the last 5 bytes are literals and a match may not start within the last 12 bytes
    while i < limit:
        key = buf[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            ml = 4
            while i + ml < n - 5 and buf[cand + ml] == buf[i + ml]:
                ml += 1
            emit(i, ml, i - cand)
            i += ml
            anchor = i
        else:
            i += 1
    emit(n, 0, 0)
    return bytes(out)


def lz4_frame_compress(data, block_size=4 << 20, compress=False, linked=False):
    """LZ4 frame, content checksum on.  Default: stored (uncompressed) independent blocks - valid for every reader and
    fast in pure Python.  compress=True runs the built-in block compressor; linked=True writes block-dependent frames
    like lz4.frame's defaults (matches reach into the previous blocks)."""
    flg = (1 << 6) | (0 if linked else (1 << 5)) | (1 << 2)   # version 01, block independence flag, content checksum
    bd_code = 7
    for code, size in ((4, 64 << 10), (5, 256 << 10), (6, 1 << 20), (7, 4 << 20)):
        if block_size <= size:
            bd_code = code
            break
    block_size = min(block_size, 4 << 20)
    desc = bytes([flg, bd_code << 4])
    out = bytearray(struct.pack("<I", _MAGIC) + desc + bytes([(xxh32(desc) >> 8) & 0xFF]))
    for i in range(0, len(data), block_size):
        blk = data[i:i + block_size]
        comp = lz4_block_compress(blk, data[max(0, i - 65536):i] if linked else b"") if compress else None
        if comp is not None and len(comp) < len(blk):
            out += struct.pack("<I", len(comp)) + comp
        else:
            out += struct.pack("<I", len(blk) | 0x80000000) + blk
    out += struct.pack("<I", 0) + struct.pack("<I", xxh32(data))
    return bytes(out)


def lz4_frame_decompress(buf):
    if struct.unpack_from("<I", buf, 0)[0] != _MAGIC:
        raise ValueError("not an LZ4 frame")
    flg = buf[4]
    linked = not (flg & (1 << 5))
    i = 6
    if flg & (1 << 3):
        i += 8          # content size
    if flg & 1:
        i += 4          # dictionary id
    if ((xxh32(buf[4:i]) >> 8) & 0xFF) != buf[i]:
        raise ValueError("LZ4 frame header checksum mismatch")
    i += 1
    out = bytearray()
    while True:
        size = struct.unpack_from("<I", buf, i)[0]; i += 4
        if size == 0:
            break
        raw, size = bool(size & 0x80000000), size & 0x7FFFFFFF
        blk = buf[i:i + size]; i += size
        if flg & (1 << 4):
            i += 4      # block checksum
        # linked blocks (lz4.frame's default): back-references may reach up to 64 KB into the earlier blocks' output
        out += blk if raw else lz4_block_decompress(blk, bytes(out[-65536:]) if linked else b"")
    if flg & (1 << 2) and struct.unpack_from("<I", buf, i)[0] != xxh32(bytes(out)):
        raise ValueError("LZ4 frame content checksum mismatch")
    return bytes(out)


def save_episode_log(game_object, filepath, compression_level=16):
    """Save the dense log of the last logged episode as lz4-framed JSON (reference utils.py:19-36)."""
    log_bytes = json.dumps(game_object.previous_episode_dense_log, ensure_ascii=False).encode("utf-8")
    try:
        import lz4.frame
        payload = lz4.frame.compress(log_bytes, compression_level=max(0, min(16, int(compression_level))))
    except (ImportError, AttributeError):   # no lz4 package (or an empty stand-in module): the built-in frame writer
        payload = lz4_frame_compress(log_bytes)
    with open(filepath, "wb") as fh:
        fh.write(payload)


def load_episode_log(filepath):
    """Load a dense log saved by this module or by the reference (utils.py:39-43)."""
    with open(filepath, "rb") as fh:
        buf = fh.read()
    try:
        import lz4.frame
        return json.loads(lz4.frame.decompress(buf))
    except (ImportError, AttributeError):
        return json.loads(lz4_frame_decompress(buf))

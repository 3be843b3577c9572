"""Episode-log files in the reference's wire format (ai_economist/foundation/utils.py:19-43): the JSON text of
`env.previous_episode_dense_log` inside an LZ4 *frame* (what `lz4.frame.open` reads and writes).

The `lz4` package is not part of this image, so the frame container is implemented here: the writer emits standard
LZ4 frames whose blocks are stored uncompressed (a valid encoding every LZ4 reader accepts; `compression_level` is
accepted for signature compatibility), the reader handles stored and LZ4-compressed blocks, i.e. it also loads logs
written by the reference.  When `lz4` is importable it is used instead.
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


def lz4_block_decompress(src):
    """One LZ4 block (sequences of literals + back-references) -> bytes."""
    out, i, n = bytearray(), 0, len(src)
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
        for k in range(ml):  # may overlap its own output
            out.append(out[start + k])
    return bytes(out)


def lz4_frame_compress(data, block_size=4 << 20):
    """LZ4 frame with stored (uncompressed) blocks, content checksum on."""
    flg = (1 << 6) | (1 << 5) | (1 << 2)   # version 01, independent blocks, content checksum
    bd = 7 << 4                            # 4 MB max block size
    desc = bytes([flg, bd])
    out = bytearray(struct.pack("<I", _MAGIC) + desc + bytes([(xxh32(desc) >> 8) & 0xFF]))
    for i in range(0, len(data), block_size):
        blk = data[i:i + block_size]
        out += struct.pack("<I", len(blk) | 0x80000000) + blk
    out += struct.pack("<I", 0) + struct.pack("<I", xxh32(data))
    return bytes(out)


def lz4_frame_decompress(buf):
    if struct.unpack_from("<I", buf, 0)[0] != _MAGIC:
        raise ValueError("not an LZ4 frame")
    flg = buf[4]
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
        out += blk if raw else lz4_block_decompress(blk)
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

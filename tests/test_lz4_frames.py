"""The built-in LZ4 frame codec (ai_economist_b200/foundation/utils.py) against the frame variants a reader of the
reference's episode logs meets: stored blocks, compressed independent blocks, and LINKED multi-block frames - the
default of lz4.frame.open(..., compression_level=16), which the reference's writer uses (foundation/utils.py:19-36) -
where matches of block N reach back into the output of block N-1."""
import json
import random
import struct

import pytest

from ai_economist_b200.foundation import utils as u


def _payload(n_words=40000, seed=3):
    rng = random.Random(seed)
    words = [bytes(rng.choices(b'abcdefghij{}":, 0123456789', k=rng.randint(3, 12))) for _ in range(150)]
    return b"".join(rng.choice(words) for _ in range(n_words))


@pytest.mark.parametrize("linked", [False, True])
@pytest.mark.parametrize("block_size", [64 << 10, 256 << 10])
def test_compressed_frames_round_trip(linked, block_size):
    data = _payload()
    assert len(data) > 3 * (64 << 10)   # several blocks
    frame = u.lz4_frame_compress(data, block_size=block_size, compress=True, linked=linked)
    assert len(frame) < len(data)       # blocks really are LZ4-compressed
    assert bool(frame[4] & (1 << 5)) == (not linked)
    assert u.lz4_frame_decompress(frame) == data


def test_linked_frame_really_depends_on_earlier_blocks():
    """Flip the block-independence flag of a linked frame: decoding every block on its own must fail (or give different
    bytes), i.e. the linked test frame exercises back-references across block boundaries."""
    data = _payload()
    frame = bytearray(u.lz4_frame_compress(data, block_size=64 << 10, compress=True, linked=True))
    frame[4] |= 1 << 5
    frame[6] = (u.xxh32(bytes(frame[4:6])) >> 8) & 0xFF
    with pytest.raises(ValueError):
        u.lz4_frame_decompress(bytes(frame))


def test_stored_frames_and_block_codec_edges():
    for data in (b"", b"x", b"abc" * 5, bytes(range(256)) * 3, b"a" * 70000):
        assert u.lz4_frame_decompress(u.lz4_frame_compress(data)) == data
        assert u.lz4_block_decompress(u.lz4_block_compress(data)) == data
        assert u.lz4_frame_decompress(u.lz4_frame_compress(data, block_size=64 << 10, compress=True, linked=True)) == data
    # known-answer block from the LZ4 block format description: literals "abcd", then a match of 8 at offset 4
    assert u.lz4_block_decompress(bytes([0x44]) + b"abcd" + struct.pack("<H", 4) + bytes([0x50]) + b"tail!") == b"abcd" * 3 + b"tail!"


def test_episode_log_files(tmp_path):
    class Env:
        previous_episode_dense_log = {"world": [{"Stone": [[0, 1], [1, 0]]}] * 50, "states": [{"0": {"loc": [1, 2]}}] * 50}

    p = str(tmp_path / "log.lz4")
    u.save_episode_log(Env(), p)
    assert u.load_episode_log(p) == Env.previous_episode_dense_log
    # a log written the reference's way (compressed, linked blocks) loads too
    raw = json.dumps(Env.previous_episode_dense_log).encode()
    with open(p, "wb") as fh:
        fh.write(u.lz4_frame_compress(raw * 40, block_size=64 << 10, compress=True, linked=True))
    assert u.lz4_frame_decompress(open(p, "rb").read()) == raw * 40

"""L0: the oracle's numpy-legacy MT19937 stream (raw words, rand(), permutation()) against numpy itself."""
import numpy as np

from oracle.oracle import OracleBatch
from tests import golden_utils as gu


def _batch_with_state(rs):
    z, meta, init = gu.load_fixture(gu.golden_files()[0])
    b = OracleBatch(meta["spec"], 1)
    st = dict(init)
    key = rs.get_state()
    st["mt_key"], st["mt_pos"] = np.array(key[1], np.uint32), int(key[2])
    b.load_env(0, st)
    return b


def test_raw_words_match_numpy():
    rs = np.random.RandomState(12345)
    b = _batch_with_state(rs)
    got = b.rng_words(0, 5000)
    # legacy randint over the full uint32 range consumes exactly one tempered word per draw
    ref = rs.randint(0, 2 ** 32, size=5000, dtype=np.uint32)
    assert np.array_equal(got, ref)


def test_rand_matches_numpy():
    rs = np.random.RandomState(99)
    b = _batch_with_state(rs)
    got = np.array([b.rng_rand(0) for _ in range(3000)])
    assert np.array_equal(got, rs.rand(3000))


def test_permutation_matches_numpy_and_interleaves():
    rs = np.random.RandomState(2024)
    b = _batch_with_state(rs)
    for n in [1, 2, 3, 4, 5, 10, 31, 32, 33, 64, 100]:
        for _ in range(20):
            assert np.array_equal(b.rng_permutation(0, n), rs.permutation(n))
            assert b.rng_rand(0) == rs.rand()

"""Dense logs: `env.previous_episode_dense_log` of replica 0 against the unmodified reference's, for
every episode of the multi-episode golden traces (tests/golden_dense/, recorded with dense_log_frequency=1).
The component logs come from the device-side per-step event buffer; world / states from the state record."""
import glob
import json
import math
import os

import numpy as np
import pytest

from ai_economist_b200 import foundation
from tests import golden_utils as gu

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden_dense", "*.json")))


def same(ref, got, path="log"):
    """Deep compare: same structure and keys, numbers within 1e-6 relative (ints exactly), None == NaN."""
    if isinstance(ref, dict):
        assert isinstance(got, dict) and set(ref) == set(got), "%s: keys %s vs %s" % (path, sorted(ref)[:8], sorted(got)[:8])
        for k in ref:
            same(ref[k], got[k], path + "/" + str(k))
    elif isinstance(ref, list):
        assert isinstance(got, (list, tuple)) and len(ref) == len(got), "%s: length %d vs %d" % (path, len(ref), len(got))
        for i, (r, g) in enumerate(zip(ref, got)):
            same(r, g, "%s[%d]" % (path, i))
    elif isinstance(ref, str):
        assert ref == got, "%s: %r vs %r" % (path, ref, got)
    elif ref is None:
        assert got is None or (isinstance(got, float) and math.isnan(got)), path
    else:
        assert abs(float(got) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref))), "%s: %r vs reference %r" % (path, got, ref)


def _replay(path, factory):
    ref = json.load(open(path))
    if "fixture" in ref:   # actions / config of a multi-episode golden trace
        z, meta, init = gu.load_fixture(os.path.join(HERE, "golden_reset", ref["fixture"]))
        act_a, act_p = z["act_a"], z["act_p"]
    else:                  # self-contained fixture (build-policy trace)
        meta = {"reference_kwargs": ref["reference_kwargs"], "seed": ref["seed"], "n_steps": len(ref["act_a"])}
        act_a, act_p = np.asarray(ref["act_a"], np.int32), np.zeros((len(ref["act_a"]), 0), np.int32)
    kw = dict(meta["reference_kwargs"])
    name = kw.pop("scenario_name")
    kw["components"] = [tuple(c) for c in kw["components"]]
    extra = dict(stepper_factory=factory) if factory else dict(device="cuda:0")
    env = foundation.make_env_instance(name, n_envs=2, auto_reset=True, dense_log_frequency=1,
                                       world_dense_log_frequency=ref["world_dense_log_frequency"], **kw, **extra)
    env.seed([meta["seed"], meta["seed"] + 1])   # replica 0 is the logged one
    env.reset()
    want = dict(zip(ref["steps"], ref["logs"]))
    A, seen = env.n_agents, 0
    for t in range(1, int(meta["n_steps"]) + 1):
        acts = {str(i): np.repeat(act_a[t - 1][i][None], 2, axis=0) for i in range(A)}
        if act_p.shape[1]:
            acts["p"] = np.repeat(act_p[t - 1][None], 2, axis=0)
        env.step(acts)
        if t in want:
            same(want[t], env.previous_episode_dense_log, "episode ending at t=%d" % t)
            seen += 1
    assert seen >= 2


def test_fixtures_cover_every_event_kind():
    cnt = {k: 0 for k in ("Build", "Gather", "Trade", "PeriodicTax")}
    for p in FILES:
        for log in json.load(open(p))["logs"]:
            for k in cnt:
                cnt[k] += sum(1 for x in log.get(k, []) if x)
    assert all(v >= 5 for v in cnt.values()), cnt


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p))
def test_emulated_dense_log_matches_reference(path):
    from tests.emu.emu_stepper import emu_factory
    _replay(path, emu_factory)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p))
def test_cuda_dense_log_matches_reference(path):
    _replay(path, None)


def test_episode_log_wire_format_round_trip(tmp_path):
    """lz4-frame JSON files (reference utils.py:19-43) without the lz4 package: known-answer checksum, a hand-built
    compressed block, and a save/load round trip."""
    from ai_economist_b200.foundation import utils
    assert utils.xxh32(b"") == 0x02CC5D05 and utils.xxh32(b"abc") == 0x32D153FF   # published xxHash32 test vectors
    assert utils.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F
    assert utils.lz4_block_decompress(bytes([0x35]) + b"abc" + bytes([3, 0])) == b"abc" * 4
    # a frame with one compressed block, as the reference's lz4.frame writer would produce
    body = bytes([0x35]) + b"abc" + bytes([3, 0])
    desc = bytes([0x60, 0x40])
    frame = (b"\x04\x22\x4d\x18" + desc + bytes([(utils.xxh32(desc) >> 8) & 0xFF]) +
             len(body).to_bytes(4, "little") + body + (0).to_bytes(4, "little"))
    assert utils.lz4_frame_decompress(frame) == b"abcabcabcabc"

    class Obj:
        previous_episode_dense_log = json.load(open(FILES[0]))["logs"][0]
    path = str(tmp_path / "episode.lz4")
    utils.save_episode_log(Obj, path)
    assert open(path, "rb").read(4) == b"\x04\x22\x4d\x18"
    assert utils.load_episode_log(path) == Obj.previous_episode_dense_log

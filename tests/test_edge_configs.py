"""Edge-of-range configurations (oracle/configs.py: EDGE_CONFIGS): two agents, 1x1 and wider-than-world windows, one
order slot, 32 price levels, no auction / no build component, tax every step, disabled taxes, linear / log brackets,
eta near 1, unscaled observations, a non-square world with 33 agents.

  * build container (reference present): the C oracle and the product's host-side reset against the live reference;
  * everywhere: the device source (1-lane emulation) against the oracle through the public API;
  * `-m gpu`: the CUDA build against the oracle, same harness (ordered last, see tests/conftest.py).
"""
import numpy as np
import pytest

from oracle import configs
from oracle import ref_harness as rh

EDGE = sorted(configs.EDGE_CONFIGS)


def _all_configs():
    configs.CONFIGS.update(configs.EDGE_CONFIGS)


def make_product_env(cfg, n_envs, stepper_factory, seed, **extra):
    from ai_economist_b200 import foundation
    kw = dict(configs.EDGE_CONFIGS[cfg])
    name = kw.pop("scenario_name")
    return foundation.make_env_instance(name, n_envs=n_envs, stepper_factory=stepper_factory, auto_reset=False,
                                        seed=seed, **kw, **extra)


def run_against_oracle(env, steps, check_every):
    from oracle.oracle import OracleBatch
    from tests import batch_utils as bu
    host = env.host_reset_arrays()
    env.load_host_state(host)
    orc = OracleBatch(env.spec, env.n_envs)
    for e in range(env.n_envs):
        orc.load_env(e, {k: v[e] for k, v in host.items()})
    for e in range(env.n_envs):
        bu.compare_env(orc, env.stepper, e, "reset", spatial=bool(env.spec["planner_gets_spatial_info"]))
    bu.run_pair(env, orc, steps, np.random.RandomState(17), check_every=check_every)


@pytest.mark.parametrize("cfg", EDGE)
def test_emulated_device_code_matches_oracle_on_edge_config(cfg):
    from tests.emu.emu_stepper import emu_factory
    run_against_oracle(make_product_env(cfg, 3, emu_factory, seed=900), steps=40, check_every=10)


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("cfg", EDGE)
def test_oracle_and_host_reset_track_live_reference_on_edge_config(cfg):
    from oracle.validate_vs_reference import run
    _all_configs()
    assert run(cfg, 501, 40, verbose=False)
    # the product's host-side reset: same spec, same post-reset state, same stream position
    f = rh.load_reference_foundation()
    ref = f.make_env_instance(**configs.EDGE_CONFIGS[cfg])
    ref.seed(77)
    ref.reset()
    want_spec, want = rh.spec_from_reference_env(ref), rh.state_from_reference_env(ref)
    env = make_product_env(cfg, 1, lambda *a, **k: None, seed=None)
    env.seed(77)
    got = env.host_reset_arrays()
    for k, v in want_spec.items():
        if k != "components":
            assert env.spec[k] == v, k
    for k in ["stone", "wood", "stone_src", "wood_src", "water", "loc", "mt_key", "coin", "build_payment",
              "build_skill", "bonus_gather_prob"]:
        assert np.array_equal(np.asarray(got[k][0]), np.asarray(want[k])), k
    assert int(got["mt_pos"][0]) == int(want["mt_pos"])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", EDGE)
def test_cuda_matches_oracle_on_edge_config(cfg):
    run_against_oracle(make_product_env(cfg, 24, None, seed=900, device="cuda:0"), steps=40, check_every=20)

"""The device random policy (aie_sample_random_actions, and the same policy fused into the step's observation pass by
aie_set_fused_policy): every drawn action is unmasked, every unmasked action of a segment can be drawn, and the draws
change from step to step."""
import numpy as np
import pytest

from ai_economist_b200 import foundation
from oracle.configs import CONFIGS
from tests import batch_utils as bu


def _check(env, rounds=6, fused=False, min_seen=10):
    st = env.stepper
    spec = env.spec
    seg_a, seg_p = bu.segments(spec, "a"), bu.segments(spec, "p")
    seen = set()
    prev = None
    if fused:
        st.set_fused_policy(77)   # samples once now; afterwards every step refreshes the action buffers itself
    for it in range(rounds):
        if not fused:
            st.sample_random_actions(seed=100 + it)
        aa = st.to_numpy(st.buf["actions_agent"]).copy()
        ma = st.to_numpy(st.buf["mask_agent"])
        off = 0
        for si, n in enumerate(seg_a):        # every subspace of every agent of every env: the pick is open
            pick = aa[..., si if spec["multi_action_agents"] else 0]
            assert np.all((pick >= 0) & (pick < n))
            assert np.all(np.take_along_axis(ma[..., off:off + n], pick[..., None], axis=-1) == 1.0)
            off += n
            if not spec["multi_action_agents"]:
                break
        if seg_p:
            ap = st.to_numpy(st.buf["actions_planner"])
            mp = st.to_numpy(st.buf["mask_planner"])
            off = 0
            for b, n in enumerate(seg_p):
                assert np.all(np.take_along_axis(mp[:, off:off + n], ap[:, b][:, None], axis=-1) == 1.0)
                off += n
        seen.update(np.unique(aa).tolist())
        assert prev is None or not np.array_equal(prev, aa)
        prev = aa
        env.step(env.action_buffers)
    assert len(seen) > min_seen


@pytest.mark.parametrize("cfg", ["c1_tutorial", "tax_us_federal", "c3_reset", "tax_single_planner"])
def test_emulated_sampler_draws_only_unmasked_actions(cfg):
    from tests.emu.emu_stepper import emu_factory
    kw = dict(CONFIGS[cfg])
    name = kw.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=5, stepper_factory=emu_factory, **kw)
    env.seed(list(range(5)))
    env.reset()
    _check(env)
    _check(env, rounds=12, fused=True, min_seen=3)   # same env, a few steps later: less coin, fewer open price levels


@pytest.mark.parametrize("nt", [1, 32, 128])
@pytest.mark.parametrize("cfg", ["c1_tutorial", "tax_us_federal", "c3_reset", "tax_single_planner"])
def test_emulated_fused_policy_draws_only_unmasked_actions(cfg, nt, monkeypatch):
    from tests.emu.emu_stepper import emu_factory
    monkeypatch.setenv("AIE_EMU_NT", str(nt))
    kw = dict(CONFIGS[cfg])
    name = kw.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=5, stepper_factory=emu_factory, **kw)
    env.seed(list(range(5)))
    env.reset()
    _check(env, rounds=12, fused=True)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["c1_tutorial", "tax_us_federal", "c3_reset", "tax_single_planner"])
def test_cuda_sampler_draws_only_unmasked_actions(cfg):
    kw = dict(CONFIGS[cfg])
    name = kw.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=257, device="cuda:0", seeds=list(range(257)), **kw)
    env.reset()
    _check(env)
    _check(env, rounds=12, fused=True, min_seen=3)   # same env, a few steps later: less coin, fewer open price levels

"""-m gpu: the CUDA build against the 1-lane host emulation of the same device source (tests/emu), many distinct seeds.

The algorithmic content of the reset paths is pinned on the CPU against the reference (golden traces, fuzz); what only the
GPU can show is that the warp-level code - lane-parallel draws from the MT19937 key (np.random.rand / randn fills with
ballot-numbered accepted trials), the 128-word twist chunks, warp reductions - lands on the same state whatever the
stream alignment.  The golden traces exercise one seed each; here 48 replicas with their own streams run through three
auto-resets (dynamic layouts: a new clumped layout per reset) and every replica is compared with its emulated twin.
Bit-exact: maps, locations, inventories, masks, the MT19937 key and position; 1e-9 relative: float64 state.
"""
import numpy as np
import pytest

from ai_economist_b200 import foundation
from oracle import configs
from tests import batch_utils as bu

pytestmark = pytest.mark.gpu


def _pair(name, kw, E, seeds):
    from tests.emu.emu_stepper import emu_factory
    cuda = foundation.make_env_instance(name, n_envs=E, device="cuda:0", auto_reset=True, seeds=seeds, **kw)
    emu = foundation.make_env_instance(name, n_envs=E, stepper_factory=emu_factory, auto_reset=True, seeds=seeds, **kw)
    for env in (cuda, emu):
        env.seed([s + 7 for s in seeds])
        env.reset()
    return cuda, emu


def _bits(a):
    """Same bit pattern whatever the signedness of the view (torch has no uint32: the CUDA stepper's key view is int32)."""
    a = np.ascontiguousarray(a)
    return a.view(np.dtype("u%d" % a.dtype.itemsize)) if a.dtype.kind in "iu" else a


def _compare(cuda, emu, E, label):
    sc, se = cuda.stepper, emu.stepper
    for k in ("cell", "owner", "loc", "inv", "esc", "mt_key", "mt_pos", "t", "completions"):
        a, b = _bits(sc.to_numpy(sc.state_view(k))), _bits(se.to_numpy(se.state_view(k)))
        assert np.array_equal(a, b), "%s: state %s (first bad env %s)" % (
            label, k, np.argwhere(a.reshape(E, -1) != b.reshape(E, -1))[:1].tolist())
    for k in ("coin", "labor", "build_payment", "build_skill", "bonus_gather_prob"):
        a, b = sc.to_numpy(sc.state_view(k)), se.to_numpy(se.state_view(k))
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12), "%s: state %s" % (label, k)
    for k in ("mask_agent", "obs_agent_idx", "obs_agent_map"):
        if int(np.prod(sc.buf[k].shape)):
            assert np.array_equal(sc.to_numpy(sc.buf[k]), se.to_numpy(se.buf[k])), "%s: %s" % (label, k)
    for k in ("obs_agent_flat", "obs_planner_flat", "reward"):
        assert np.allclose(sc.to_numpy(sc.buf[k]), se.to_numpy(se.buf[k]), rtol=1e-6, atol=1e-7), "%s: %s" % (label, k)


@pytest.mark.parametrize("cfg", ["uniform_reset", "quadrant_reset", "multi_zone_reset", "lognormal_reset", "c5_full"])
def test_cuda_device_reset_equals_emulation_for_many_streams(cfg):
    kw = dict(configs.CONFIGS[cfg])
    name = kw.pop("scenario_name")
    kw["components"] = [tuple(c) for c in kw["components"]]
    kw["episode_length"] = 9
    E = 6 if cfg == "c5_full" else 48   # c5: 64 agents on 64x64 (one CTA of four warps per env, 26 key regenerations per fill)
    cuda, emu = _pair(name, kw, E, [5000 + 13 * e for e in range(E)])
    assert cuda.spec["reset_mode"] == 1
    _compare(cuda, emu, E, "reset")
    seg_a, seg_p = bu.segments(cuda.spec, "a"), bu.segments(cuda.spec, "p")
    rng = np.random.RandomState(4)
    se = emu.stepper
    for t in range(1, 3 * 9 + 2):
        aa = bu.sample_from_masks(se.to_numpy(se.buf["mask_agent"]), seg_a, rng)
        ap = bu.sample_from_masks(se.to_numpy(se.buf["mask_planner"]), seg_p, rng) if seg_p else None
        cuda.step((aa, ap)); emu.step((aa, ap))
        if t % 9 in (0, 1) or t == 5:   # the step that resets (new layout, placement, skills), the one after, one in between
            _compare(cuda, emu, E, "t=%d" % t)
    assert int(np.min(cuda.stepper.to_numpy(cuda.stepper.state_view("completions")))) == 3


def test_cuda_one_step_economy_equals_emulation_for_many_streams():
    kw = dict(components=[("SimpleLabor", dict(mask_first_step=True, payment_max_skill_multiplier=3)),
                          ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=2, tax_model="model_wrapper", rate_disc=0.05))],
              n_agents=40, world_size=[1, 1], episode_length=2, planner_reward_type="coin_eq_times_productivity",
              agent_reward_type="isoelastic_coin_minus_labor", labor_cost=0.02)
    E = 64
    cuda, emu = _pair("one-step-economy", kw, E, [900 + e for e in range(E)])
    seg_a, seg_p = bu.segments(cuda.spec, "a"), bu.segments(cuda.spec, "p")
    rng = np.random.RandomState(8)
    se, sc = emu.stepper, cuda.stepper
    for t in range(1, 9):
        aa = bu.sample_from_masks(se.to_numpy(se.buf["mask_agent"]), seg_a, rng)
        ap = bu.sample_from_masks(se.to_numpy(se.buf["mask_planner"]), seg_p, rng)
        cuda.step((aa, ap)); emu.step((aa, ap))
        for k in ("mt_key", "mt_pos", "t"):
            assert np.array_equal(_bits(sc.to_numpy(sc.state_view(k))), _bits(se.to_numpy(se.state_view(k)))), (t, k)
        for k in ("coin", "labor", "build_payment"):
            assert np.allclose(sc.to_numpy(sc.state_view(k)), se.to_numpy(se.state_view(k)), rtol=1e-9), (t, k)
        assert np.array_equal(sc.to_numpy(sc.buf["mask_agent"]), se.to_numpy(se.buf["mask_agent"]))
        for k in ("obs_agent_flat", "obs_planner_flat", "obs_planner_agents", "reward"):
            assert np.allclose(sc.to_numpy(sc.buf[k]), se.to_numpy(se.buf[k]), rtol=1e-6, atol=1e-7), (t, k)

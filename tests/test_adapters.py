"""RL-framework adapters: WarpDrive-style names/spaces and the RLlib-style per-replica dict env,
driven on the CPU through the emulated device code."""
import numpy as np

from ai_economist_b200 import adapters, foundation
from oracle.configs import CONFIGS
from tests.emu.emu_stepper import emu_factory


def _env(n_envs=3):
    kw = dict(CONFIGS["c3_reset"])
    name = kw.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=n_envs, auto_reset=True, stepper_factory=emu_factory, **kw)
    env.seed([11 + i for i in range(n_envs)])
    return env


def test_warpdrive_style_wrapper_names_spaces_and_zero_copy():
    env = _env()
    w = adapters.WarpDriveStyleEnvWrapper(env)
    assert w.n_agents == env.n_agents + 1 and w.episode_length == env.episode_length
    w.reset_all_envs()
    A = env.n_agents
    assert set(env.observation_space.keys()) == set(env.action_space.keys()) == {str(i) for i in range(A)} | {"p"}
    assert env.action_space["0"].n == env.get_agent(0).action_spaces            # single-action agents
    assert list(env.action_space["p"].nvec) == list(env.get_agent("p").action_spaces)
    for k, sp in env.observation_space["0"].items():
        assert tuple(sp.shape) == tuple(env.obs["0"][k].shape[1:]), k
    # reserved names are views of the very buffers the kernels write
    assert np.shares_memory(w.tensor("actions_a"), env.stepper.buf["actions_agent"])
    assert np.shares_memory(w.tensor("rewards_a"), env.stepper.buf["reward"])
    assert w.tensor("observations_a_flat").shape[:2] == (env.n_envs, A)
    # a "policy" writes actions in place, the wrapper steps from the buffers
    mask = w.tensor("observations_a_action_mask")
    w.tensor("actions_a")[..., 0] = np.argmax(mask * (np.arange(mask.shape[-1]) > 0), axis=-1)
    obs, rew, done, info = w.step_all_envs()
    assert int(w.tensor("_timestep_")[0]) == 1 and w.tensor("rewards_p").shape == (env.n_envs,)
    assert rew["0"].shape == (env.n_envs,) and not bool(done["__all__"][0])


def test_rllib_style_dict_env_matches_direct_stepping():
    env_a, env_b = _env(2), _env(2)
    env_b.reset(); env_b.reset()    # the wrapper resets once at construction (like the reference's) and once below
    d = adapters.MultiAgentDictEnv(env_a, e=1)
    obs = d.reset()
    assert d.observation_space["flat"].shape == obs["0"]["flat"].shape
    assert d.observation_space_pl["flat"].shape == obs["p"]["flat"].shape
    rng = np.random.RandomState(3)
    for t in range(35):       # crosses an episode boundary (episode_length 30): summary becomes available
        acts = {}
        for i in range(env_a.n_agents):
            m = obs[str(i)]["action_mask"]
            acts[str(i)] = int(rng.choice(len(m), p=m / m.sum()))
        full = {k: np.array([0, v], np.int32) for k, v in acts.items()}   # replica 0 idle, replica 1 acts
        env_b.step(full)
        obs, rew, done, info = d.step(acts)
        ob, rb, db = env_b.reference_view(1)
        assert np.array_equal(obs["0"]["flat"], ob["0"]["flat"]) and rew == rb and done == db
    s = d.summary
    assert s["completions"] == 1 and "social/productivity" in s


import pytest


@pytest.mark.gpu
def test_warpdrive_style_wrapper_on_cuda_is_zero_copy():
    import torch
    kw = dict(CONFIGS["c3_reset"])
    name = kw.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=64, device="cuda:0", seeds=list(range(64)), **kw)
    w = adapters.WarpDriveStyleEnvWrapper(env)
    w.reset_all_envs()
    assert w.tensor("actions_a").data_ptr() == env.stepper.buf["actions_agent"].data_ptr()
    assert w.tensor("observations_a_world-map").is_cuda and w.tensor("rewards_a").shape == (64, env.n_agents)
    mask = w.tensor("observations_a_action_mask")
    w.tensor("actions_a")[..., 0] = torch.argmax(mask * (torch.arange(mask.shape[-1], device=mask.device) > 0), dim=-1)
    w.step_all_envs()
    assert int(w.tensor("_timestep_")[5]) == 1 and bool(torch.isfinite(w.tensor("rewards_a")).all())
    d = adapters.MultiAgentDictEnv(env, e=7)
    obs, rew, done, _ = d.step({"0": 0})
    assert obs["p"]["flat"].dtype == np.float32 and set(rew) == {str(i) for i in range(env.n_agents)} | {"p"}


def test_warpdrive_style_wrapper_over_the_covid_env():
    """The env the reference's FoundationEnvWrapper was written for (env_wrapper.py:84-418 + covid19_env.py's reserved
    array names): spaces per US state + planner, zero-copy named tensors, reset_all_envs / step_all_envs."""
    import json
    import os
    from ai_economist_b200 import foundation
    from ai_economist_b200.adapters import WarpDriveStyleEnvWrapper
    from oracle.gen_golden_covid import COVID_KWARGS, reference_config
    from tests.emu.emu_stepper import EmuCovidStepper
    cfg = reference_config(COVID_KWARGS)
    name = cfg.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=3, auto_reset=True,
                                       stepper_factory=lambda p, n, ar: EmuCovidStepper(p, n, ar), **cfg)
    w = WarpDriveStyleEnvWrapper(env)
    assert w.n_agents == 52 and w.n_envs == 3 and w.episode_length == env.episode_length
    assert set(env.action_space.keys()) == {str(i) for i in range(51)} | {"p"}
    assert env.action_space["0"].n == 11 and env.action_space["p"].n == 21
    assert env.observation_space["7"]["world-agent_state"].shape == (6,) and env.observation_space["p"]["world-agent_state"].shape == (6, 51)
    w.reset_all_envs()
    assert np.shares_memory(w.tensor("actions_a"), env.stepper.buf["actions_agent"])
    w.tensor("actions_a")[...] = 0
    w.tensor("actions_a")[1, 5] = 3                      # state 5 of replica 1 goes to stringency level 3
    w.tensor("actions_p")[...] = 2
    w.step_all_envs()
    assert int(w.tensor("_timestep_")[0]) == 1 and not w.tensor("_done_").any()
    pol = w.tensor("observations_a_ControlUSStateOpenCloseStatus-agent_policy_indicators")
    assert w.tensor("rewards_a").shape == (3, 51) and w.tensor("rewards_p").shape == (3,)
    st = env.stepper.read_state(1)
    assert int(st["stringency"][5]) == 3 and np.asarray(pol).shape == (3, 51)

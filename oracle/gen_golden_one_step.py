"""TEST INFRASTRUCTURE: golden traces of the one-step-economy scenario (SimpleLabor + PeriodicBracketTax) recorded from the
UNMODIFIED reference (build container only) -> tests/golden_one_step/*.npz.  Several episodes per trace with
env.reset() between them (the global numpy stream continues), random unmasked actions from a separate RandomState.
Usage: python oracle/gen_golden_one_step.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import ref_harness as rh  # noqa: E402
import fuzz_one_step_vs_reference as fz  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden_one_step")
_TAX = ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=2, tax_model="model_wrapper", rate_disc=0.05, usd_scaling=1000.0))
CONFIGS = {
    # the intended use (simple_labor.py:29-31, one_step_economy.py:27-33): 2-step episodes, taxes first, then labor
    "paper_like": dict(components=[("SimpleLabor", dict(mask_first_step=True, payment_max_skill_multiplier=3, pareto_param=4.0)), _TAX],
                       n_agents=10, episode_length=2, agent_reward_type="coin_minus_labor_cost", labor_exponent=2.0, labor_cost=1.0,
                       planner_reward_type="inv_income_weighted_utility"),
    # 31 agents (sorted Gini form), isoelastic utilities, equality-weighted planner, longer episodes, tax component first
    "isoelastic_31": dict(components=[_TAX, ("SimpleLabor", dict(mask_first_step=False, payment_max_skill_multiplier=5))],
                          n_agents=31, episode_length=4, agent_reward_type="isoelastic_coin_minus_labor", isoelastic_eta=0.23,
                          labor_cost=0.05, planner_reward_type="coin_eq_times_productivity", mixing_weight_gini_vs_coin=0.4),
    # no tax component, multi-action agents, non-integer labor exponent
    "labor_only": dict(components=[("SimpleLabor", dict(mask_first_step=True, payment_max_skill_multiplier=1.5))],
                       n_agents=4, episode_length=3, agent_reward_type="coin_minus_labor_cost", labor_exponent=1.5, labor_cost=0.2,
                       multi_action_mode_agents=True, planner_reward_type="inv_income_weighted_utility"),
}
BASE = dict(scenario_name="one-step-economy", world_size=[1, 1], multi_action_mode_agents=False, multi_action_mode_planner=True,
            flatten_observations=True, flatten_masks=True)


def generate(name, seed=1001, episodes=5):
    cfg = dict(BASE, **CONFIGS[name])
    f = rh.load_reference_foundation()
    np.random.seed(seed)
    ref = f.make_env_instance(**cfg)
    ref.seed(seed + 1)
    obs = ref.reset()
    rec = {k: [v] for k, v in fz.reference_arrays(ref, obs).items()}
    arng = np.random.RandomState(seed + 2)
    A, T = ref.n_agents, cfg["episode_length"]
    acts_a, acts_p, rews, dones, metrics = [], [], [], [], {}
    for t in range(1, episodes * T + 1):
        actions, a_act, p_act = rh.sample_actions(ref, obs, arng)
        obs, rew, done, _ = ref.step(actions)
        acts_a.append(a_act); acts_p.append(p_act)
        rews.append(np.array([rew[str(i)] for i in range(A)] + [rew["p"]]))
        dones.append(int(done["__all__"]))
        if done["__all__"]:
            with np.errstate(all="ignore"):
                metrics[t] = {k: float(v) for k, v in ref.metrics.items()}
            obs = ref.reset()
        for k, v in fz.reference_arrays(ref, obs).items():
            rec[k].append(v)
    out = {"meta_json": np.array(json.dumps(dict(reference_kwargs=cfg, seed=seed, n_steps=episodes * T, metrics=metrics)))}
    out["act_a"] = np.stack(acts_a).astype(np.int16)
    out["act_p"] = np.stack(acts_p).astype(np.int16) if acts_p[0].size else np.zeros((len(acts_p), 0), np.int16)
    out["rew"], out["done"] = np.stack(rews), np.array(dones, np.int32)
    for k, v in rec.items():
        out[k] = np.stack([np.asarray(x) for x in v])
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "%s_seed%d.npz" % (name, seed))
    np.savez_compressed(path, **out)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    for name in CONFIGS:
        generate(name)

"""TEST INFRASTRUCTURE: pin `env.metrics` against the UNMODIFIED reference.

For every single-episode fixture in tests/golden/ this re-runs the reference with the fixture's config and seed,
replays the recorded action trace (so the trajectory is the one the fixture pins step by step) and records
`env.metrics` (base_env.py:421-432) at the fixture's full-compare steps.  Output: tests/golden_metrics/<name>.json.
Runs in the build container only (needs /root/reference).  Usage: python oracle/gen_golden_metrics.py
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden_metrics")


def actions_dict(env, a_row, p_row):
    acts = {}
    for i in range(env.n_agents):
        ag = env.get_agent(i)
        acts[str(i)] = [int(v) for v in a_row[i]] if ag.multi_action_mode else int(a_row[i][0])
    if len(p_row):
        acts["p"] = [int(v) for v in p_row] if env.get_agent("p").multi_action_mode else int(p_row[0])
    return acts


def clean(m):
    return {k: (None if isinstance(v, float) and np.isnan(v) else (float(v) if isinstance(v, (float, np.floating)) else int(v)))
            for k, v in ((k, (v.item() if hasattr(v, "item") else v)) for k, v in m.items())}


def generate(path):
    z = np.load(path)
    meta = json.loads(str(z["spec_json"]))
    f = rh.load_reference_foundation()
    env = f.make_env_instance(**meta["reference_kwargs"])
    env.seed(meta["seed"])
    env.reset()
    steps = sorted(int(t) for t in z["full_steps"] if t > 0)
    out = {"fixture": os.path.basename(path), "steps": [], "metrics": []}
    with np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        for t in range(1, int(meta["n_steps"]) + 1):
            env.step(actions_dict(env, z["act_a"][t - 1], z["act_p"][t - 1]))
            if t in steps:
                out["steps"].append(t)
                out["metrics"].append(clean(env.metrics))
    os.makedirs(OUT, exist_ok=True)
    dst = os.path.join(OUT, os.path.basename(path).replace(".npz", ".json"))
    with open(dst, "w") as fh:
        json.dump(out, fh)
    print("%s: %d snapshots, %d keys" % (dst, len(out["steps"]), len(out["metrics"][-1])))


def generate_reset(path):
    """Multi-episode traces: `env.previous_episode_metrics` (base_env.py:414-418, stored by reset()) after every reset."""
    z = np.load(path)
    meta = json.loads(str(z["spec_json"]))
    f = rh.load_reference_foundation()
    env = f.make_env_instance(**meta["reference_kwargs"])
    env.seed(meta["seed"])
    env.reset()
    out = {"fixture": os.path.basename(path), "steps": [], "metrics": []}
    import warnings
    warnings.simplefilter("ignore")
    n = int(meta["n_steps"])
    for t in range(1, n + 1):
        _, _, done, _ = env.step(actions_dict(env, z["act_a"][t - 1], z["act_p"][t - 1]))
        if done["__all__"] and t < n:
            env.reset()
            out["steps"].append(t)
            out["metrics"].append(clean(env.previous_episode_metrics))
    dst = os.path.join(OUT + "_reset", os.path.basename(path).replace(".npz", ".json"))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as fh:
        json.dump(out, fh)
    print("%s: %d episodes" % (dst, len(out["steps"])))


def generate_dense(path):
    """Multi-episode traces with dense_log_frequency=1: `env.previous_episode_dense_log` (base_env.py:763-814) of every
    finished episode -> tests/golden_dense/<name>.json."""
    z = np.load(path)
    meta = json.loads(str(z["spec_json"]))
    f = rh.load_reference_foundation()
    kw = dict(meta["reference_kwargs"]); kw["dense_log_frequency"] = 1; kw["world_dense_log_frequency"] = 10
    env = f.make_env_instance(**kw)
    env.seed(meta["seed"])
    env.reset()
    out = {"fixture": os.path.basename(path), "world_dense_log_frequency": 10, "steps": [], "logs": []}
    n = int(meta["n_steps"])
    for t in range(1, n + 1):
        _, _, done, _ = env.step(actions_dict(env, z["act_a"][t - 1], z["act_p"][t - 1]))
        if done["__all__"] and t < n:
            out["steps"].append(t)
            out["logs"].append(env.previous_episode_dense_log)
            env.reset()
    dst = os.path.join(ROOT, "tests", "golden_dense", os.path.basename(path).replace(".npz", ".json"))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as fh:
        json.dump(out, fh)
    print("%s: %d episode logs, %.0f KB" % (dst, len(out["logs"]), os.path.getsize(dst) / 1024))


def generate_dense_build(seed=7, episode_length=120, episodes=2):
    """A trace that builds houses: agents take Build whenever it is unmasked, otherwise a random unmasked action.
    Self-contained fixture (kwargs, seed, actions, dense logs) -> tests/golden_dense/build_policy_seed<seed>.json."""
    from oracle.configs import CONFIGS
    f = rh.load_reference_foundation()
    kw = dict(CONFIGS["c1_reset"]); kw["episode_length"] = episode_length
    kw["dense_log_frequency"] = 1; kw["world_dense_log_frequency"] = 40
    env = f.make_env_instance(**kw)
    env.seed(seed)
    obs = env.reset()
    rng = np.random.RandomState(seed + 1)
    out = {"reference_kwargs": {k: v for k, v in kw.items() if k not in ("dense_log_frequency", "world_dense_log_frequency")},
           "seed": seed, "world_dense_log_frequency": 40, "act_a": [], "steps": [], "logs": []}
    n = episode_length * episodes
    for t in range(1, n + 1):
        acts, a_act, _ = rh.sample_actions(env, obs, rng)
        for i in range(env.n_agents):
            if np.asarray(obs[str(i)]["action_mask"])[1] > 0:   # single-action agents: index 1 = Build (first component)
                acts[str(i)] = 1
                a_act[i] = [1]
        out["act_a"].append([[int(v) for v in row] for row in a_act])
        obs, _, done, _ = env.step(acts)
        if done["__all__"]:
            out["steps"].append(t)
            out["logs"].append(env.previous_episode_dense_log)
            if t < n:
                obs = env.reset()
    dst = os.path.join(ROOT, "tests", "golden_dense", "build_policy_seed%d.json" % seed)
    with open(dst, "w") as fh:
        json.dump(out, fh)
    nb = sum(len(x) for log in out["logs"] for x in log["Build"])
    print("%s: %d episode logs, %d builds, %.0f KB" % (dst, len(out["logs"]), nb, os.path.getsize(dst) / 1024))


if __name__ == "__main__":
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        generate(p)
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden_reset", "*.npz"))):
        generate_reset(p)
        if "saez" not in p:   # the 790-step Saez trace would add 4 MB of dense logs without new event kinds
            generate_dense(p)
    generate_dense_build()

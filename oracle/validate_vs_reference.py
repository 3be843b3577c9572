"""TEST INFRASTRUCTURE: step the imported reference and the C oracle side by side.

Usage: python oracle/validate_vs_reference.py [config ...] [--steps N] [--seeds a,b]
"""
import argparse
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402
from oracle.oracle import OracleBatch  # noqa: E402
from oracle.configs import CONFIGS  # noqa: E402


def compare(name, a, b, exact=True, rtol=1e-6, atol=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return "%s: shape %s vs %s" % (name, a.shape, b.shape)
    if exact:
        if not np.array_equal(a, b):
            idx = np.argwhere(a != b)[:5]
            return "%s: mismatch at %s ref=%s orc=%s" % (name, idx.tolist(), a[tuple(idx[0])], b[tuple(idx[0])])
    else:
        if not np.allclose(a, b, rtol=rtol, atol=atol):
            idx = np.argwhere(~np.isclose(a, b, rtol=rtol, atol=atol))[:5]
            return "%s: mismatch at %s ref=%s orc=%s" % (name, idx.tolist(), a[tuple(idx[0])], b[tuple(idx[0])])
    return None


EXACT_STATE = ["cell", "owner", "loc", "inv", "esc", "n_orders", "bid_hist", "ask_hist", "tax_pos", "rate_idx",
               "mt_key", "mt_pos", "t"]
CLOSE_STATE = ["coin", "esc_coin", "labor", "price_hist", "last_coin", "last_income", "last_marg"]
EXACT_OBS = ["a_map", "a_idx", "a_mask", "p_map", "p_idx", "p_mask", "done"]
CLOSE_OBS = ["a_flat", "p_flat", "p_agents", "time", "rew"]


def run(cfg_name, seed, steps, verbose=True):
    f = rh.load_reference_foundation()
    cfg = dict(CONFIGS[cfg_name])
    env = f.make_env_instance(**cfg)
    env.seed(seed)
    obs = env.reset()
    spec = rh.spec_from_reference_env(env)
    orc = OracleBatch(spec, 1)
    orc.load_env(0, rh.state_from_reference_env(env))
    arng = np.random.RandomState(seed + 7919)
    errs = []

    def check(t, obs, rew=None, done=None):
        ro = rh.obs_arrays_from_reference(env, obs, rew, done)
        oo = orc.obs(0)
        rs = rh.state_arrays_from_reference(env)
        os_ = orc.state(0)
        for k in EXACT_OBS:
            if k in ro:
                e = compare(k, ro[k], oo[k]); errs.append(e) if e else None
        for k in CLOSE_OBS:
            if k in ro:
                e = compare(k, ro[k], oo[k], exact=False, rtol=1e-6, atol=1e-7); errs.append(e) if e else None
        for k in EXACT_STATE:
            if k in rs:
                e = compare(k, rs[k], os_[k]); errs.append(e) if e else None
        for k in CLOSE_STATE:
            if k in rs:
                e = compare(k, rs[k], os_[k], exact=False, rtol=1e-9, atol=1e-9); errs.append(e) if e else None
        if "book" in rs:
            for (c, side), rows in rs["book"].items():
                e = compare("book%d%d" % (c, side), rows, orc.book(0, c, side)); errs.append(e) if e else None
        if errs:
            print("[%s seed %d] step %d: %d mismatches" % (cfg_name, seed, t, len(errs)))
            for e in errs[:10]:
                print("   ", e)
            return False
        return True

    if not check(0, obs):
        return False
    n_trades = n_builds = 0
    for t in range(1, steps + 1):
        actions, a_act, p_act = rh.sample_actions(env, obs, arng)
        obs, rew, done, _ = env.step(actions)
        orc.step(a_act[None], p_act[None] if p_act.size else None)
        if not check(t, obs, rew, done):
            return False
        if done["__all__"]:
            break
    if verbose:
        m = env.metrics
        print("[%s seed %d] %d steps OK  (trades=%s, builds=%s)" % (
            cfg_name, seed, t, m.get("Trade/n_trades", m.get("ContinuousDoubleAuction/n_trades")),
            m.get("Build/total_builds")))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=list(CONFIGS))
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--seeds", default="1001,1002")
    a = ap.parse_args()
    ok = True
    for c in a.configs:
        for s in [int(x) for x in a.seeds.split(",")]:
            ok &= run(c, s, a.steps)
    sys.exit(0 if ok else 1)

"""Helpers for batched product-vs-oracle runs (used by the `-m gpu` tests, smoke-style checks and bench.py)."""
import os

import numpy as np

from ai_economist_b200 import workloads as _wl
from oracle.configs import CONFIGS


def product_kwargs(cfg_name):
    return _wl.product_kwargs(cfg_name, CONFIGS)


segments = _wl.mask_segments
sample_from_masks = _wl.sample_from_masks


EXACT_STATE = ["cell", "owner", "loc", "inv", "esc", "n_orders", "bid_hist", "ask_hist", "tax_pos", "rate_idx",
               "mt_key", "mt_pos", "t"]
FLOAT_STATE = ["coin", "esc_coin", "labor", "price_hist", "last_coin", "last_income", "last_marg"]
EXACT_OBS = ["a_map", "a_idx", "a_mask", "p_map", "p_idx", "p_mask", "done"]
FLOAT_OBS = ["a_flat", "p_flat", "p_agents", "time", "rew"]


def compare_env(orc, stepper, e, label, spatial=True, rtol=1e-6, skip=()):
    """Bit-exact on integer/grid/index/mask work, <= rtol relative on coin/labor/utility floats."""
    oo, os_ = orc.obs(e), orc.state(e)
    po, ps = stepper.read_obs(e), stepper.read_state(e)
    has_tax = "PeriodicBracketTax" in stepper.spec["components"]
    tax_keys = ("tax_pos", "rate_idx", "last_coin", "last_income", "last_marg")
    has_cda = "ContinuousDoubleAuction" in stepper.spec["components"]
    if not has_cda:   # no auction: the book / histogram arrays are empty on both sides (and sized differently)
        tax_keys = tax_keys + ("n_orders", "bid_hist", "ask_hist", "price_hist")
    for k in EXACT_STATE:
        if (k in tax_keys[:5] and not has_tax) or (k in tax_keys[5:] and not has_cda):
            continue
        assert np.array_equal(os_[k], np.asarray(ps[k]).reshape(os_[k].shape)), "%s env %d: state %s" % (label, e, k)
    for k in FLOAT_STATE:
        if (k in tax_keys[:5] and not has_tax) or (k in tax_keys[5:] and not has_cda):
            continue
        assert np.allclose(os_[k], np.asarray(ps[k]).reshape(os_[k].shape), rtol=rtol, atol=1e-9), \
            "%s env %d: state %s" % (label, e, k)
    # episode statistics behind env.metrics (running sums of the reference's event logs) and the reward trackers
    assert np.allclose(os_["stats"], ps["stats"], rtol=rtol, atol=1e-9), "%s env %d: episode statistics" % (label, e)
    if "util_prev" not in skip:
        assert np.allclose(os_["util_prev"], ps["util_prev"], rtol=rtol, atol=1e-9), "%s env %d: util_prev" % (label, e)
    for c in ((0, 1) if has_cda else ()):
        for s in (0, 1):
            assert np.array_equal(orc.book(e, c, s), ps["books"][(c, s)]), "%s env %d: book %d/%d" % (label, e, c, s)
    for k in EXACT_OBS:
        if k in skip:
            continue
        if k in po and (spatial or k not in ("p_map", "p_idx")):
            assert np.array_equal(oo[k], np.asarray(po[k]).reshape(oo[k].shape)), "%s env %d: obs %s" % (label, e, k)
    for k in FLOAT_OBS:
        if k in skip:
            continue
        assert np.allclose(oo[k], np.asarray(po[k]).reshape(oo[k].shape), rtol=rtol, atol=1e-7), \
            "%s env %d: obs %s" % (label, e, k)


def run_pair(env, orc, steps, rng, check_every=0, check_envs=None, on_check=None):
    """Step the product env and the oracle with identical mask-aware random actions."""
    spec, st = env.spec, env.stepper
    seg_a, seg_p = segments(spec, "a"), segments(spec, "p")
    E = env.n_envs
    check_envs = list(range(E)) if check_envs is None else check_envs
    for t in range(1, steps + 1):
        ma = st.to_numpy(st.buf["mask_agent"])
        aa = sample_from_masks(ma, seg_a, rng)
        ap = None
        if seg_p:
            ap = sample_from_masks(st.to_numpy(st.buf["mask_planner"]), seg_p, rng)
        env.step((aa, ap))
        orc.step(aa, ap, n_threads=min(os.cpu_count() or 1, max(1, E // 8)))
        if check_every and (t % check_every == 0 or t == steps):
            for e in check_envs:
                compare_env(orc, st, e, "t=%d" % t, spatial=bool(spec["planner_gets_spatial_info"]))
            if on_check:
                on_check(t)

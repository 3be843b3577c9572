"""Replay a golden fixture (tests/golden/*.npz, generated from the unmodified reference by
oracle/gen_golden.py) against a stepper and compare.

A "stepper" exposes:  load(init_state_dict), step(act_a [A,n], act_p [B]), obs() -> dict, state() -> dict
in the array layout of oracle/oracle.py (which the CUDA product's debug readback mirrors).

Tolerances (BASELINE.json north_star): bit-exact for grid / inventory / integer / index / mask work;
<= 1e-6 relative for coin / labor / utility / reward / float observations.
"""
import glob
import json
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-6          # the north-star float tolerance
ATOL_F64 = 1e-9      # absolute slack for values that are exactly 0 in the reference
ATOL_F32 = 1e-7      # float32 observations

EXACT_OBS = ["a_map", "a_idx", "a_mask", "p_map", "p_idx", "p_mask"]
FLOAT_OBS = ["a_flat", "p_flat", "p_agents", "time"]
EXACT_STEP_STATE = ["loc", "inv", "esc", "n_orders", "tax_pos", "rate_idx", "mt_pos"]
FLOAT_STEP_STATE = ["coin", "esc_coin", "labor", "last_coin", "last_income", "last_marg"]


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def load_fixture(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["spec_json"]))
    init = {k[5:]: z[k] for k in z.files if k.startswith("init_")}
    init["mt_pos"] = int(init["mt_pos"])
    init["completions"] = int(init["completions"])
    return z, meta, init


def book_crc(books):
    """books: dict {(c, side): rows[n,3] in the reference's stored order}."""
    return crc(np.concatenate([np.asarray(books[(c, s)], np.int32).ravel() for c in (0, 1) for s in (0, 1)]
                              + [np.zeros(1, np.int32)]))


def check_step(z, t, obs, state, full_idx, label, books=None, spatial_planner=True):
    def fail(msg):
        raise AssertionError("%s step %d: %s" % (label, t, msg))

    for k in EXACT_STEP_STATE:
        key = "step_" + k
        if key in z.files:
            got = np.asarray(state[k]).reshape(z[key][t].shape)
            if not np.array_equal(z[key][t], got):
                fail("%s ref=%s got=%s" % (k, z[key][t].tolist(), got.tolist()))
    for k in FLOAT_STEP_STATE:
        key = "step_" + k
        if key in z.files:
            got = np.asarray(state[k]).reshape(z[key][t].shape)
            if not np.allclose(z[key][t], got, rtol=RTOL, atol=ATOL_F64):
                fail("%s ref=%s got=%s" % (k, z[key][t].tolist(), got.tolist()))
    if crc(np.asarray(state["cell"], np.uint8)) != int(z["step_cell_crc"][t]):
        fail("cell map differs")
    if crc(np.asarray(state["owner"], np.int8)) != int(z["step_owner_crc"][t]):
        fail("house owner map differs")
    if crc(np.asarray(state["mt_key"], np.uint32)) != int(z["step_mt_key_crc"][t]):
        fail("MT19937 key differs")
    if "step_hist_crc" in z.files:
        h = crc(np.concatenate([np.asarray(state["bid_hist"], np.int32).ravel(),
                                np.asarray(state["ask_hist"], np.int32).ravel()]))
        if h != int(z["step_hist_crc"][t]):
            fail("bid/ask histograms differ")
        if not np.isclose(float(np.asarray(state["price_hist"]).sum()), float(z["step_price_hist_sum"][t]),
                          rtol=RTOL, atol=ATOL_F64):
            fail("price_history sum differs")
        if books is not None and book_crc(books) != int(z["step_book_crc"][t]):
            fail("order book differs")
    if t > 0:
        if not np.allclose(z["step_rew"][t], obs["rew"], rtol=RTOL, atol=ATOL_F64):
            fail("reward ref=%s got=%s" % (z["step_rew"][t].tolist(), np.asarray(obs["rew"]).tolist()))
        if int(z["step_done"][t]) != int(np.asarray(obs["done"]).ravel()[0]):
            fail("done flag")
    for k in EXACT_OBS:
        key = "step_%s_crc" % k
        if key in z.files and (spatial_planner or k not in ("p_map", "p_idx")):
            if crc(obs[k]) != int(z[key][t]):
                fail("observation %s differs (crc)" % k)
    if full_idx is not None:
        for k in FLOAT_OBS:
            ref = z["full_" + k][full_idx]
            got = np.asarray(obs[k]).reshape(ref.shape)
            if not np.allclose(ref, got, rtol=RTOL, atol=ATOL_F32):
                bad = np.argwhere(~np.isclose(ref, got, rtol=RTOL, atol=ATOL_F32))[:4]
                fail("%s differs at %s ref=%s got=%s" % (k, bad.tolist(), ref[tuple(bad[0])], got[tuple(bad[0])]))
        for k in EXACT_OBS:
            if "full_" + k in z.files and (spatial_planner or k not in ("p_map", "p_idx")):
                ref = z["full_" + k][full_idx]
                got = np.asarray(obs[k]).reshape(ref.shape)
                if not np.array_equal(ref.astype(got.dtype), got):
                    fail("%s differs (full compare)" % k)


def replay(path, make_stepper, max_steps=None):
    """make_stepper(spec, init) -> stepper.  Replays the recorded action trace."""
    z, meta, init = load_fixture(path)
    spec = meta["spec"]
    stepper = make_stepper(spec, init)
    label = os.path.basename(path)
    full_steps = {int(t): i for i, t in enumerate(z["full_steps"])}
    n = int(meta["n_steps"]) if max_steps is None else min(int(meta["n_steps"]), max_steps)
    sp = bool(spec["planner_gets_spatial_info"])
    check_step(z, 0, stepper.obs(), stepper.state(), full_steps.get(0), label,
               books=stepper.books() if hasattr(stepper, "books") else None, spatial_planner=sp)
    for t in range(1, n + 1):
        ap = z["act_p"][t - 1].astype(np.int32)
        stepper.step(z["act_a"][t - 1].astype(np.int32), ap if ap.size else None)
        check_step(z, t, stepper.obs(), stepper.state(), full_steps.get(t), label,
                   books=stepper.books() if hasattr(stepper, "books") else None, spatial_planner=sp)
    return n

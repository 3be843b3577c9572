"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/libfoundation_oracle.so.

The oracle is the CPU checker for the CUDA product.  Only tests/, bench.py's
cpu_baseline / `--impl reference` leg and __graft_entry__.smoke() may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfoundation_oracle.so")

MAX_COMP, MAX_BRACKETS, MAX_RATES = 8, 16, 64
COMP = {"Build": 0, "ContinuousDoubleAuction": 1, "Gather": 2, "PeriodicBracketTax": 3, "WealthRedistribution": 4}


class OrcConfig(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("episode_length", C.c_int32), ("multi_action_agents", C.c_int32),
        ("n_comp", C.c_int32), ("comp", C.c_int32 * MAX_COMP),
        ("has_water", C.c_int32), ("obs_range", C.c_int32),
        ("planner_gets_spatial_info", C.c_int32), ("allow_observation_scaling", C.c_int32),
        ("regen_weight", C.c_double * 2),
        ("isoelastic_eta", C.c_double), ("energy_cost", C.c_double),
        ("energy_warmup_constant", C.c_double), ("energy_warmup_auto", C.c_int32),
        ("planner_reward_type", C.c_int32), ("mixing_weight_gini_vs_coin", C.c_double),
        ("build_payment", C.c_double), ("build_labor", C.c_double),
        ("move_labor", C.c_double), ("collect_labor", C.c_double),
        ("max_bid_ask", C.c_int32), ("order_duration", C.c_int32), ("max_num_orders", C.c_int32),
        ("order_labor", C.c_double),
        ("tax_model", C.c_int32), ("disable_taxes", C.c_int32), ("period", C.c_int32),
        ("n_brackets", C.c_int32), ("n_disc_rates", C.c_int32),
        ("bracket_cutoffs", C.c_double * MAX_BRACKETS), ("disc_rates", C.c_double * MAX_RATES),
        ("fixed_rates", C.c_double * MAX_BRACKETS),
        ("tax_annealing", C.c_int32), ("annealing_warmup", C.c_double),
        ("annealing_slope", C.c_double), ("rate_max", C.c_double),
        ("single_action_planner", C.c_int32), ("regen_halfwidth", C.c_int32 * 2),
        ("full_observability", C.c_int32),
    ]


class OrcDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ["n_map_ch", "win", "flat_a", "flat_p", "flat_pa", "mask_a", "mask_p",
                 "n_act_a", "n_act_p", "book_cap", "a_map_elems", "a_idx_elems"]]


def build(force=False):
    """Compile the oracle (gcc, a few hundred ms)."""
    src = os.path.join(_HERE, "foundation_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src),
                                                   os.path.getmtime(os.path.join(_HERE, "foundation_oracle.h")))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig), C.c_int32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_dims_from_config.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcDims)]
        L.orc_load_env.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 13 + [C.c_int32, C.c_int32]
        L.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_get_obs.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 12
        L.orc_get_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 20
        L.orc_get_book.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_get_book.restype = C.c_int32
        L.orc_get_stats.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_stats.restype = C.c_int32
        L.orc_rng_words.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_rng_rand.argtypes = [C.c_void_p, C.c_int32]
        L.orc_rng_rand.restype = C.c_double
        L.orc_rng_permutation.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        _lib = L
    return _lib


def config_from_spec(spec):
    """spec: plain dict (see ai_economist_b200 EnvSpec.to_dict() / ref_harness.spec_from_reference_env)."""
    cfg = OrcConfig()
    for k in ["n_agents", "height", "width", "episode_length", "multi_action_agents", "has_water",
              "obs_range", "planner_gets_spatial_info", "allow_observation_scaling", "isoelastic_eta",
              "energy_cost", "energy_warmup_constant", "energy_warmup_auto", "planner_reward_type",
              "mixing_weight_gini_vs_coin", "build_payment", "build_labor", "move_labor", "collect_labor",
              "max_bid_ask", "order_duration", "max_num_orders", "order_labor", "tax_model",
              "disable_taxes", "period", "n_brackets", "n_disc_rates", "tax_annealing",
              "annealing_warmup", "annealing_slope", "rate_max"]:
        setattr(cfg, k, spec[k])
    comps = [COMP[c] for c in spec["components"]]
    cfg.n_comp = len(comps)
    for i, c in enumerate(comps):
        cfg.comp[i] = c
    cfg.regen_weight[0], cfg.regen_weight[1] = spec["regen_weight"]
    for i, v in enumerate(spec["bracket_cutoffs"]):
        cfg.bracket_cutoffs[i] = v
    for i, v in enumerate(spec["disc_rates"]):
        cfg.disc_rates[i] = v
    for i, v in enumerate(spec["fixed_rates"]):
        cfg.fixed_rates[i] = v
    cfg.single_action_planner = int(spec.get("single_action_planner", 0))
    cfg.regen_halfwidth[0], cfg.regen_halfwidth[1] = spec.get("regen_halfwidth", [0, 0])
    cfg.full_observability = int(spec.get("full_observability", 0))
    return cfg


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """E independent CPU envs stepping the restated reference algorithm."""

    def __init__(self, spec, n_envs):
        self.spec = dict(spec)
        self.cfg = config_from_spec(spec)
        self.dims = OrcDims()
        lib().orc_dims_from_config(C.byref(self.cfg), C.byref(self.dims))
        self.n_envs = int(n_envs)
        self._h = lib().orc_create(C.byref(self.cfg), self.n_envs)
        self.A, self.H, self.W = spec["n_agents"], spec["height"], spec["width"]
        self.P = spec["max_bid_ask"] + 1

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_env(self, e, st):
        """st: post-reset host state dict (u8 maps [H,W], loc i16 [A,2], f64 [A] vectors, mt_key, mt_pos)."""
        A = self.A
        st = dict(st)
        st.setdefault("inv_stone", np.zeros(A, np.int32))
        st.setdefault("inv_wood", np.zeros(A, np.int32))
        g = lambda k, dt: np.ascontiguousarray(st[k], dtype=dt)
        arrs = [g("stone", np.uint8), g("wood", np.uint8), g("stone_src", np.uint8), g("wood_src", np.uint8),
                g("water", np.uint8), g("loc", np.int16), g("coin", np.float64),
                g("inv_stone", np.int32), g("inv_wood", np.int32), g("build_payment", np.float64),
                g("build_skill", np.float64), g("bonus_gather_prob", np.float64), g("mt_key", np.uint32)]
        rc = lib().orc_load_env(self._h, e, *[_p(a) for a in arrs],
                                int(st["mt_pos"]), int(st.get("completions", 0)))
        assert rc == 0
        return rc

    def step(self, actions_a=None, actions_p=None, n_threads=1):
        """actions_a: int32 [E, A, n_act_a] (or [E, A] single-action); actions_p: int32 [E, n_act_p]."""
        aa = None if actions_a is None else np.ascontiguousarray(actions_a, dtype=np.int32)
        ap = None if actions_p is None else np.ascontiguousarray(actions_p, dtype=np.int32)
        if aa is not None:
            assert aa.size == self.n_envs * self.A * self.dims.n_act_a, (aa.shape, self.dims.n_act_a)
        if ap is not None:
            assert ap.size == self.n_envs * self.dims.n_act_p
        lib().orc_step(self._h, _p(aa), _p(ap), int(n_threads))

    def obs(self, e):
        d, A, H, W = self.dims, self.A, self.H, self.W
        full = bool(self.spec.get("full_observability", 0))
        out = dict(
            a_map=np.zeros((A, d.n_map_ch, H, W) if full else (A, d.n_map_ch + 1, d.win, d.win), np.float32),
            a_idx=np.zeros((A, 2, H, W) if full else (A, 2, d.win, d.win), np.int16),
            a_flat=np.zeros((A, d.flat_a), np.float32),
            a_mask=np.zeros((A, d.mask_a), np.float32),
            p_map=np.zeros((d.n_map_ch, H, W), np.float32),
            p_idx=np.zeros((2, H, W), np.int16),
            p_flat=np.zeros((d.flat_p,), np.float32),
            p_agents=np.zeros((A, d.flat_pa), np.float32),
            p_mask=np.zeros((d.mask_p,), np.float32),
            time=np.zeros((1,), np.float32),
            rew=np.zeros((A + 1,), np.float64),
            done=np.zeros((1,), np.int32),
        )
        lib().orc_get_obs(self._h, e, *[_p(out[k]) for k in
                                        ["a_map", "a_idx", "a_flat", "a_mask", "p_map", "p_idx", "p_flat",
                                         "p_agents", "p_mask", "time", "rew", "done"]])
        return out

    def masks(self):
        """(a_mask [E, A, mask_a], p_mask [E, mask_p]) of every env, one C call."""
        d = self.dims
        ma = np.zeros((self.n_envs, self.A, d.mask_a), np.float32)
        mp = np.zeros((self.n_envs, d.mask_p), np.float32)
        lib().orc_get_masks(self._h, _p(ma), _p(mp))
        return ma, mp

    def state(self, e):
        A, H, W, P = self.A, self.H, self.W, self.P
        B = max(1, self.spec["n_brackets"])
        out = dict(
            cell=np.zeros((H, W), np.uint8), owner=np.zeros((H, W), np.int8), loc=np.zeros((A, 2), np.int16),
            coin=np.zeros(A), esc_coin=np.zeros(A), labor=np.zeros(A),
            inv=np.zeros((A, 2), np.int32), esc=np.zeros((A, 2), np.int32),
            n_orders=np.zeros((2, A), np.int32), bid_hist=np.zeros((2, A, P), np.int32),
            ask_hist=np.zeros((2, A, P), np.int32), price_hist=np.zeros((2, A, P)),
            tax_pos=np.zeros(1, np.int32), rate_idx=np.zeros(B, np.int32),
            last_coin=np.zeros(A), last_income=np.zeros(A), last_marg=np.zeros(A),
            mt_key=np.zeros(624, np.uint32), mt_pos=np.zeros(1, np.int32), t=np.zeros(1, np.int32),
        )
        lib().orc_get_state(self._h, e, *[_p(out[k]) for k in
                                          ["cell", "owner", "loc", "coin", "esc_coin", "labor", "inv", "esc",
                                           "n_orders", "bid_hist", "ask_hist", "price_hist", "tax_pos",
                                           "rate_idx", "last_coin", "last_income", "last_marg", "mt_key",
                                           "mt_pos", "t"]])
        has_tax = "PeriodicBracketTax" in self.spec["components"]
        n_stats = 1 + A + 8 * A + ((35 + 2 * A) if has_tax else 0)
        out["stats"] = np.zeros(n_stats)
        out["util_prev"] = np.zeros(A + 1)
        out["auto_warmup"] = np.zeros(1, np.int32)
        n = lib().orc_get_stats(self._h, e, _p(out["stats"]), _p(out["util_prev"]), _p(out["auto_warmup"]))
        assert n == n_stats, (n, n_stats)
        return out

    def book(self, e, c, side):
        cap = self.dims.book_cap + self.A
        rows = np.zeros((cap, 3), np.int32)
        n = lib().orc_get_book(self._h, e, c, side, _p(rows), cap)
        return rows[:n]

    # L0 RNG helpers
    def rng_words(self, e, n):
        out = np.zeros(n, np.uint32)
        lib().orc_rng_words(self._h, e, _p(out), n)
        return out

    def rng_rand(self, e):
        return lib().orc_rng_rand(self._h, e)

    def rng_permutation(self, e, n):
        out = np.zeros(n, np.int32)
        lib().orc_rng_permutation(self._h, e, n, _p(out))
        return out

"""Dense logs of one env replica, in the reference's structure (BaseEnvironment.dense_log, base_env.py:355-356, 763-814,
984-1019): {"world": [...], "states": [...], "actions": [...], "rewards": [...], "Build": [...], "Trade": [...],
"Gather": [...], "PeriodicTax": [...]} with one entry per timestep (+ the final world / states snapshot).

The reference appends to Python lists inside every component_step; here the step kernel writes the step's events of
the logged replica into a small device buffer (aie_buffers.events: builds, gathers, trades) and this host-side logger
- active only while an episode is being logged - reads that buffer and the replica's state record around each step.
Dense logging is a slow diagnostic path in the reference too (deepcopies of the whole world every step); the other
replicas are unaffected.
"""
import numpy as np

_COMMODITY = ("Stone", "Wood")


def world_dict(spec, st):
    """Maps.state_dict (world.py:327-329) after recursive_cast: entity -> [H][W] lists, House -> {owner, health}."""
    cell = np.asarray(st["cell"]).astype(np.int64)
    bit = lambda b: ((cell >> b) & 1).astype(np.float64).tolist()
    out = {"Stone": bit(0), "Wood": bit(1),
           "House": {"owner": np.asarray(st["owner"]).astype(np.int64).tolist(), "health": bit(5)}}
    if spec["has_water"]:
        out["Water"] = bit(4)
    out["StoneSourceBlock"] = bit(2)
    out["WoodSourceBlock"] = bit(3)
    return out


def states_dict(spec, st, static):
    """{agent idx: deepcopy(agent.state)} (base_env.py:795-798); `static` = per-agent build payment / skills."""
    A = int(spec["n_agents"])
    out = {}
    for a in range(A):
        s = {"loc": [int(st["loc"][a, 0]), int(st["loc"][a, 1])],
             "inventory": {"Coin": float(st["coin"][a]), "Stone": int(st["inv"][a, 0]), "Wood": int(st["inv"][a, 1])},
             "escrow": {"Coin": float(st["esc_coin"][a]), "Stone": int(st["esc"][a, 0]), "Wood": int(st["esc"][a, 1])},
             "endogenous": {"Labor": float(st["labor"][a])}}
        if "Build" in spec["components"]:
            s["build_payment"] = float(static["build_payment"][a])
            s["build_skill"] = float(static["build_skill"][a])
        if "Gather" in spec["components"]:
            s["bonus_gather_prob"] = float(static["bonus_gather_prob"][a])
        out[str(a)] = s
    out["p"] = {"inventory": {"Coin": 0, "Stone": 0, "Wood": 0}, "escrow": {"Coin": 0, "Stone": 0, "Wood": 0},
                "endogenous": {}}
    return out


class DenseLogger:
    """Collects one episode of replica `e` of a BatchedFoundationEnv."""

    def __init__(self, env, e=0, world_every=50):
        self.env, self.e, self.world_every = env, int(e), int(world_every)
        self.spec = env.spec
        self.log = {"world": [], "states": [], "actions": [], "rewards": []}
        logging = {"Build", "Gather", "ContinuousDoubleAuction"} | (set() if self.spec["disable_taxes"] else {"PeriodicBracketTax"})
        self.comp = {c.shorthand: [] for c in env.components if c.name in logging}  # the others return None
        self._paid_prev = None
        self._static = None

    # -- helpers -----------------------------------------------------------------------------------------
    def _static_fields(self):
        st = self.env.stepper
        if self._static is None:
            g = lambda k: st.to_numpy(st.state_view(k))[self.e]
            self._static = {"build_payment": g("build_payment"), "build_skill": g("build_skill"),
                            "bonus_gather_prob": g("bonus_gather_prob")}
        return self._static

    def _actions_entry(self, act_a, act_p):
        """{idx: {subspace name: value > 0}} (base_env.py:1000-1005) from the raw action rows of this replica."""
        out = {}
        for ag in self.env.all_agents:
            names = list(ag._action_names)
            if ag.idx == "p" and ag.multi_action_mode:
                row = [] if act_p is None else [int(v) for v in np.asarray(act_p).reshape(-1)][:len(names)]
                out["p"] = {n: v for n, v in zip(names, row) if v > 0}
                continue
            if ag.idx == "p":   # single-action planner: one index, decoded like the agents' below
                row = np.zeros(1, np.int64) if act_p is None or not names else np.asarray(act_p).reshape(-1)
            else:
                row = np.asarray(act_a[int(ag.idx)]).reshape(-1)
            if ag.multi_action_mode:
                out[str(ag.idx)] = {n: int(v) for n, v in zip(names, row) if int(v) > 0}
            else:
                g, d, lo = int(row[0]), {}, 1
                for n in names:  # single_action_map (base_agent.py:109-114)
                    k = int(ag.action_dim[n])
                    if lo <= g < lo + k:
                        d[n] = g - lo + 1
                    lo += k
                out[str(ag.idx)] = d
        return out

    # -- per-step hooks ----------------------------------------------------------------------------------
    def before_step(self, act_a, act_p):
        st = self.env.stepper.read_state(self.e)
        t = int(st["t"][0])
        self.log["world"].append(world_dict(self.spec, st) if t % self.world_every == 0 else {})
        self.log["states"].append(states_dict(self.spec, st, self._static_fields()))
        self.log["actions"].append(self._actions_entry(act_a, act_p))
        if self._paid_prev is None and self.spec_has_tax():
            self._paid_prev = self._tax_paid(st)

    def spec_has_tax(self):
        return "PeriodicBracketTax" in self.spec["components"]

    def _tax_paid(self, st):
        A = int(self.spec["n_agents"])
        t0 = 1 + A + 8 * A
        return np.array(st["stats"][t0 + 35 + A: t0 + 35 + 2 * A], np.float64)

    def after_step(self, rew_row, post_state):
        """rew_row: float64 [A+1]; post_state: the replica's record after the step's dynamics (the episode_final
        snapshot when the step ended the episode under auto-reset)."""
        A = int(self.spec["n_agents"])
        self.log["rewards"].append({**{str(a): float(rew_row[a]) for a in range(A)}, "p": float(rew_row[A])})
        builds, gathers, trades = [], [], []
        bpay = self._static_fields()["build_payment"]
        for ev in self.env.stepper.read_events(self.e):
            if ev[0] == 1:
                builds.append({"builder": ev[1], "loc": [ev[2], ev[3]], "income": float(bpay[ev[1]])})
            elif ev[0] == 2:
                gathers.append({"agent": ev[1], "resource": _COMMODITY[ev[2]], "n": ev[3], "loc": [ev[4], ev[5]]})
            elif ev[0] == 3:
                seller, buyer, c, ask, bid, alife, blife = ev[1:8]
                price = ask if blife <= alife else bid  # continuous_double_auction.py:297-304
                trades.append({"commodity": _COMMODITY[c], "buyer": buyer, "bid": bid, "bid_lifetime": blife,
                               "seller": seller, "ask": ask, "ask_lifetime": alife, "price": price, "cost": price,
                               "income": price})
        for short, rows in (("Build", builds), ("Gather", gathers), ("Trade", trades)):
            if short in self.comp:
                self.comp[short].append(rows)
        if "PeriodicTax" in self.comp:
            self.comp["PeriodicTax"].append(self._tax_entry(post_state))

    def _tax_entry(self, st):
        """redistribution.py:853-905: [] except on the step that closed a tax period."""
        paid_now = self._tax_paid(st)
        paid = paid_now - self._paid_prev
        self._paid_prev = paid_now
        if int(st["tax_pos"][0]) != 1:  # tax_cycle_pos restarts at 1 right after taxes were enacted
            return []
        spec, A, B = self.spec, int(self.spec["n_agents"]), int(self.spec["n_brackets"])
        if spec["tax_model"] == 0:
            rates = [float(spec["disc_rates"][int(i)]) for i in np.asarray(st["rate_idx"])[:B]]
        else:
            rates = [float(v) for v in list(spec["fixed_rates"])[:B]]
            if spec["tax_model"] == 1 and int(spec.get("tax_annealing", 0)):
                # a fixed schedule under a tax_annealing_schedule is clipped by this episode's annealed maximum
                # (redistribution.py:390-413; components/utils.py:10-57)
                done_eps = int(np.asarray(st["completions"]).reshape(-1)[0])
                vis = max(0.0, min(1.0, float(spec["annealing_slope"]) * (done_eps - float(spec["annealing_warmup"]))))
                rates = [min(r, vis * float(spec["rate_max"])) for r in rates]
        d = {"schedule": rates, "cutoffs": [float(v) for v in list(spec["bracket_cutoffs"])[:B]]}
        lump = float(np.sum(paid) / A)
        for a in range(A):
            inc = float(st["last_income"][a])
            d[str(a)] = {"income": inc, "tax_paid": float(paid[a]), "marginal_rate": float(st["last_marg"][a]),
                         "effective_rate": float(paid[a] / max(0.000001, inc)), "lump_sum": lump}
        return d

    def finalize(self, final_state):
        """base_env.py:795-814: final world / states, then every component's dense log under its shorthand."""
        self.log["world"].append(world_dict(self.spec, final_state))
        self.log["states"].append(states_dict(self.spec, final_state, self._static_fields()))
        for short, rows in self.comp.items():
            self.log[short] = rows
        return self.log

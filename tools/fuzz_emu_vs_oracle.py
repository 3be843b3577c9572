"""Randomised configuration fuzz (CPU): the device source (1-lane emulation) against the C oracle through the public API
on configurations drawn at random from the supported option space.  Not part of the test suite (unbounded); run it for a
while after touching the step / observation code:  python tools/fuzz_emu_vs_oracle.py [n_configs] [seed]"""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_economist_b200 import foundation  # noqa: E402
from oracle.oracle import OracleBatch  # noqa: E402
from tests import batch_utils as bu  # noqa: E402
from tests.emu.emu_stepper import emu_factory  # noqa: E402

LAYOUTS = {(15, 15): "env-pure_and_mixed-15x15.txt", (25, 25): "quadrant_25x25_20each_30clump.txt",
           (40, 40): "quadrant_40x40_50each.txt"}


def random_config(rng):
    fam = rng.choice(["layout", "layout", "uniform", "quadrant", "multi_zone"])
    A = int(rng.choice([2, 3, 4, 5, 7, 10, 13]))
    comps = []
    build = ("Build", dict(skill_dist=str(rng.choice(["none", "pareto", "lognormal"])),
                           payment_max_skill_multiplier=int(rng.randint(1, 4)), build_labor=float(rng.choice([10.0, 2.5]))))
    cda = ("ContinuousDoubleAuction", dict(max_num_orders=int(rng.choice([1, 2, 5, 9])), order_duration=int(rng.choice([1, 3, 50])),
                                           max_bid_ask=int(rng.choice([3, 10, 17])), order_labor=float(rng.choice([0.25, 0.0]))))
    gather = ("Gather", dict(skill_dist=str(rng.choice(["none", "pareto", "lognormal"])), move_labor=float(rng.choice([1.0, 0.5]))))
    order = [build, cda, gather]
    rng.shuffle(order)
    keep = [c for c in order if rng.rand() < 0.85 or c[0] == "Gather"]
    comps += keep
    if rng.rand() < 0.6:
        model = str(rng.choice(["model_wrapper", "model_wrapper", "us-federal-single-filer-2018-scaled", "fixed-bracket-rates"]))
        kw = dict(period=int(rng.choice([1, 3, 10, 25])), tax_model=model, bracket_spacing="us-federal")
        if model == "fixed-bracket-rates":
            kw["fixed_bracket_rates"] = [round(float(x), 2) for x in np.sort(rng.rand(7))]
        if model == "model_wrapper":
            kw["rate_disc"] = float(rng.choice([0.05, 0.1, 0.25]))
        if rng.rand() < 0.4:   # planner-mask annealing (model_wrapper) / per-episode clipping of a fixed schedule
            kw["tax_annealing_schedule"] = [int(rng.choice([-100, 0, -2])), float(rng.choice([0.001, 0.5, 0.25]))]
        comps.append(("PeriodicBracketTax", kw))
    elif rng.rand() < 0.3:
        comps.append(("WealthRedistribution", {}))
    kw = dict(components=comps, n_agents=A, episode_length=int(rng.choice([30, 100])),
              multi_action_mode_agents=bool(rng.rand() < 0.4), multi_action_mode_planner=bool(rng.rand() < 0.7),
              flatten_observations=True, flatten_masks=True, starting_agent_coin=float(rng.choice([0, 5, 40])),
              mobile_agent_observation_range=int(rng.choice([0, 2, 5, 7])), planner_gets_spatial_info=bool(rng.rand() < 0.6),
              full_observability=bool(rng.rand() < 0.2), allow_observation_scaling=bool(rng.rand() < 0.8),
              isoelastic_eta=float(rng.choice([0.0, 0.23, 0.7])), energy_cost=float(rng.choice([0.21, 1.0])),
              energy_warmup_constant=float(rng.choice([0, 0, 3])), energy_warmup_method=str(rng.choice(["decay", "auto"])),
              planner_reward_type=str(rng.choice(["coin_eq_times_productivity", "inv_income_weighted_coin_endowments",
                                                  "inv_income_weighted_utility"])),
              mixing_weight_gini_vs_coin=float(rng.choice([0.0, 0.3])))
    if fam == "layout":
        size = list(LAYOUTS)[rng.randint(len(LAYOUTS))]
        kw.update(world_size=list(size), env_layout_file=LAYOUTS[size], resource_regen_prob=float(rng.choice([0.01, 0.3, 1.0])))
        name = "layout_from_file/simple_wood_and_stone"
    else:
        H, W = int(rng.randint(9, 30)), int(rng.randint(9, 30))
        if fam == "quadrant":
            W = H
        kw.update(world_size=[H, W], starting_wood_coverage=0.1, starting_stone_coverage=0.1,
                  wood_regen_weight=float(rng.choice([0.01, 0.5])), stone_regen_weight=float(rng.choice([0.01, 0.5])),
                  wood_regen_halfwidth=int(rng.choice([0, 0, 1, 3])), stone_regen_halfwidth=int(rng.choice([0, 0, 2])))
        if fam == "multi_zone":
            kw.update(num_partitions_row=3, num_partitions_col=3, num_wood_zones=3, num_stone_zones=3, num_wood_and_stone_zones=2)
        name = fam + "/simple_wood_and_stone"
    return name, kw


def run_one(name, kw, seed, steps=120):
    env = foundation.make_env_instance(name, n_envs=2, stepper_factory=emu_factory, auto_reset=False, seed=seed, **kw)
    host = env.host_reset_arrays()
    env.load_host_state(host)
    orc = OracleBatch(env.spec, 2)
    for e in range(2):
        orc.load_env(e, {k: v[e] for k, v in host.items()})
    for e in range(2):
        bu.compare_env(orc, env.stepper, e, "reset", spatial=bool(env.spec["planner_gets_spatial_info"]))
    bu.run_pair(env, orc, min(steps, kw["episode_length"]), np.random.RandomState(seed), check_every=15)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for i in range(n):
        name, kw = random_config(rng)
        try:
            run_one(name, kw, seed=1 + i)
        except (AssertionError, Exception) as ex:   # noqa: BLE001
            if isinstance(ex, (NotImplementedError, TimeoutError)) or "assert" in type(ex).__name__.lower() and "coverage" in str(ex):
                print("[%d] skipped: %r" % (i, ex))
                continue
            bad += 1
            print("[%d] FAILED %s %r\n    %s" % (i, name, kw, "".join(traceback.format_exception_only(type(ex), ex)).strip()[:600]))
    print("%d configs, %d failures" % (n, bad))

"""The BASELINE.json workloads as reference-style kwargs (what `foundation.make_env_instance` takes), plus the small
host-side helpers a caller needs to drive them with a random policy (per-subspace slices of the flattened action masks).

Sources (reference repo): tutorials/economic_simulation_basic.ipynb cell 11 (configs 1/2),
tutorials/rllib/phase2/config.yaml:7-51 (config 3, at 10 agents / 40x40), tests/run_covid19_cpu_gpu_consistency_checks.py:
44-81 (config 4), BASELINE.json configs[4] (config 5: there is no 64x64 layout file, so `uniform/...`).
"""
import numpy as np

_GTB = [("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
        ("ContinuousDoubleAuction", dict(max_num_orders=5)),
        ("Gather", dict())]

BASELINE_CONFIGS = {
    # configs[0] / [1]: tutorial basic, 4 agents, 25x25
    "c1_tutorial": dict(
        scenario_name="layout_from_file/simple_wood_and_stone", components=_GTB,
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=10,
        fixed_four_skill_and_loc=True, n_agents=4, world_size=[25, 25], episode_length=1000,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # configs[2]: paper config (phase 2) at 10 agents / 40x40
    "c3_paper_tax": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=100, rate_disc=0.05,
                                                tax_model="model_wrapper"))],
        env_layout_file="quadrant_40x40_50each.txt", starting_agent_coin=0,
        fixed_four_skill_and_loc=True, n_agents=10, world_size=[40, 40], episode_length=1000,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        isoelastic_eta=0.23, energy_cost=0.21, energy_warmup_constant=0, planner_gets_spatial_info=False,
        mixing_weight_gini_vs_coin=0.0, planner_reward_type="coin_eq_times_productivity"),
    # configs[4]: ContinuousDoubleAuction stress - 64 agents, 64x64, deep book (K=50), multi-action agents
    "c5_full": dict(
        scenario_name="uniform/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=50)),
                    ("Gather", dict(skill_dist="pareto"))],
        n_agents=64, world_size=[64, 64], episode_length=150,
        multi_action_mode_agents=True, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        starting_agent_coin=100, starting_wood_coverage=0.10, starting_stone_coverage=0.10),
}

# configs[3]: the COVID-19 scenario with the reference's consistency-check settings
COVID_KWARGS = dict(
    episode_length=540, start_date="2020-03-22", pop_between_age_18_65=0.6, infection_too_sick_to_work_rate=0.1,
    risk_free_interest_rate=0.03, economic_reward_crra_eta=2, health_priority_scaling_agents=0.3,
    health_priority_scaling_planner=0.45, action_cooldown_period=28, subsidy_interval=90, num_subsidy_levels=20,
    max_annual_subsidy_per_person=20000, daily_vaccines_per_million_people=3000, delivery_interval=1,
    vaccine_delivery_start_date="2021-01-12")


def covid_reference_config(kw=None):
    """The reference-style env config (scenario kwargs + component list) of the COVID workload."""
    kw = dict(COVID_KWARGS if kw is None else kw)
    return {
        "scenario_name": "CovidAndEconomySimulation",
        "components": [
            {"ControlUSStateOpenCloseStatus": {"action_cooldown_period": kw["action_cooldown_period"]}},
            {"FederalGovernmentSubsidy": {"num_subsidy_levels": kw["num_subsidy_levels"],
                                          "subsidy_interval": kw["subsidy_interval"],
                                          "max_annual_subsidy_per_person": kw["max_annual_subsidy_per_person"]}},
            {"VaccinationCampaign": {"daily_vaccines_per_million_people": kw["daily_vaccines_per_million_people"],
                                     "delivery_interval": kw["delivery_interval"],
                                     "vaccine_delivery_start_date": kw["vaccine_delivery_start_date"]}},
        ],
        "use_real_world_data": False, "use_real_world_policies": False, "start_date": kw["start_date"],
        "path_to_data_and_fitted_params": "", "economic_reward_crra_eta": kw["economic_reward_crra_eta"],
        "health_priority_scaling_agents": kw["health_priority_scaling_agents"],
        "health_priority_scaling_planner": kw["health_priority_scaling_planner"],
        "infection_too_sick_to_work_rate": kw["infection_too_sick_to_work_rate"],
        "pop_between_age_18_65": kw["pop_between_age_18_65"], "risk_free_interest_rate": kw["risk_free_interest_rate"],
        "n_agents": 51, "world_size": [1, 1], "episode_length": kw["episode_length"],
        "multi_action_mode_agents": False, "multi_action_mode_planner": False, "flatten_observations": False,
        "flatten_masks": True, "collate_agent_step_and_reset_data": True,
    }


def product_kwargs(cfg_name, table=None):
    """(scenario_name, kwargs) of one named configuration, ready for foundation.make_env_instance."""
    kw = dict((table or BASELINE_CONFIGS)[cfg_name])
    name = kw.pop("scenario_name")
    return name, kw


def mask_segments(spec, who):
    """Lengths of the per-subspace slices of the flattened action mask (base_agent.py:440-460); who: "a" | "p"."""
    P = spec["max_bid_ask"] + 1
    if who == "p":
        planner_acts = ("PeriodicBracketTax" in spec["components"] and spec["tax_model"] == 0
                        and not spec["disable_taxes"])
        if planner_acts and spec.get("single_action_planner", 0):   # one index over [NO-OP] ++ every bracket's rates
            return [1 + spec["n_disc_rates"] * spec["n_brackets"]]
        return [1 + spec["n_disc_rates"]] * spec["n_brackets"] if planner_acts else []
    sizes = []
    for c in spec["components"]:
        if c == "Build":
            sizes += [1]
        elif c == "ContinuousDoubleAuction":
            sizes += [P] * 4
        elif c == "Gather":
            sizes += [4]
    if spec["multi_action_agents"]:
        return [s + 1 for s in sizes]
    return [1 + sum(sizes)]


def sample_from_masks(mask, seg, rng):
    """mask [..., L] (0/1) -> int32 [..., len(seg)]: uniform over the unmasked entries of each segment."""
    out = np.zeros(mask.shape[:-1] + (len(seg),), np.int32)
    off = 0
    for i, n in enumerate(seg):
        m = mask[..., off:off + n]
        out[..., i] = np.argmax(m * (rng.random_sample(m.shape) + 1e-3), axis=-1)
        off += n
    return out

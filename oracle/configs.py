"""TEST INFRASTRUCTURE: the env configurations the parity ladder runs (reference kwargs).

Sources: BASELINE.json configs; tutorials/economic_simulation_basic.ipynb cell 11 (c1/c2);
tutorials/rllib/phase2/config.yaml:7-51 (c3); tests/test_env.py:27-62 (ref_unit_test).
"""

from ai_economist_b200.workloads import BASELINE_CONFIGS, _GTB  # the BASELINE workloads live with the product

CONFIGS = {
    # c1 / c2: tutorial basic; c3: paper config (phase 2) at 10 agents / 40x40 (see ai_economist_b200/workloads.py)
    "c1_tutorial": BASELINE_CONFIGS["c1_tutorial"],
    "c3_paper_tax": BASELINE_CONFIGS["c3_paper_tax"],
    # c3 variant: short tax period, spatial planner, random (non-fixed-four) placement, tax annealing
    "c3_short_period": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=10, rate_disc=0.05,
                                                tax_model="model_wrapper",
                                                tax_annealing_schedule=[-100, 0.001]))],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5,
        fixed_four_skill_and_loc=False, n_agents=6, world_size=[25, 25], episode_length=200,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # fixed us-federal schedule (planner has no actions)
    "tax_us_federal": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="none")),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict()),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=20,
                                                tax_model="us-federal-single-filer-2018-scaled"))],
        env_layout_file="env-pure_and_mixed-15x15.txt", starting_agent_coin=20,
        n_agents=5, world_size=[15, 15], episode_length=120,
        multi_action_mode_agents=True, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # the reference's own unit-test config (tests/test_env.py:27-62)
    "ref_unit_test": dict(
        scenario_name="uniform/simple_wood_and_stone",
        components=[("Build", {}), ("ContinuousDoubleAuction", {"max_num_orders": 5}), ("Gather", {})],
        n_agents=4, world_size=[15, 15], episode_length=100,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        starting_agent_coin=10, starting_wood_coverage=0.10, starting_stone_coverage=0.10),
    # c5-like (small): multi-action agents, deep book, uniform scenario, >=30 agents (sorted gini branch)
    "c5_small": dict(
        scenario_name="uniform/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=50)),
                    ("Gather", dict(skill_dist="pareto"))],
        n_agents=32, world_size=[32, 32], episode_length=150,
        multi_action_mode_agents=True, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        starting_agent_coin=100, starting_wood_coverage=0.10, starting_stone_coverage=0.10),
    # c5 at the BASELINE size: 64 agents, 64x64, deep book (K=50), multi-action agents
    "c5_full": BASELINE_CONFIGS["c5_full"],
    # WealthRedistribution (components/redistribution.py:21-75) as the last component; pareto gather skills, 6 agents
    "wealth_redistribution": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("WealthRedistribution", dict())],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=10,
        fixed_four_skill_and_loc=False, n_agents=9, world_size=[25, 25], episode_length=400,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # Saez tax model (device/host hybrid): short tax period and 10 agents so that the 500-sample buffer fills after 500
    # steps; multi-episode, so the estimator state has to survive resets
    "saez_reset": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=10, tax_model="saez",
                                                usd_scaling=10000.0))],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5,
        fixed_four_skill_and_loc=False, n_agents=10, world_size=[25, 25], episode_length=100,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # the same under a tax_annealing_schedule: the maximum rate of an episode (warm-up draws, the formula's clip, the rates
    # in force and observed) follows the completed-episode count - 0.3, 0.45, 0.6, ... of rate_max over the trace's episodes
    "saez_annealed_reset": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=10, tax_model="saez",
                                                usd_scaling=10000.0, tax_annealing_schedule=[-2, 0.15]))],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5,
        fixed_four_skill_and_loc=False, n_agents=10, world_size=[25, 25], episode_length=100,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # short episodes for the multi-episode (device-side reset) traces
    "c1_reset": dict(
        scenario_name="layout_from_file/simple_wood_and_stone", components=_GTB,
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=10,
        fixed_four_skill_and_loc=True, n_agents=4, world_size=[25, 25], episode_length=25,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    "c3_reset": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=10, rate_disc=0.05,
                                                tax_model="model_wrapper",
                                                tax_annealing_schedule=[-100, 0.001]))],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5,
        fixed_four_skill_and_loc=False, n_agents=6, world_size=[25, 25], episode_length=30,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True, energy_warmup_constant=5, energy_warmup_method="decay"),
    # the other scenario variants of the family (host-side layouts; the step path is the same)
    "multi_zone": dict(
        scenario_name="multi_zone/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict(skill_dist="pareto"))],
        n_agents=5, world_size=[20, 20], episode_length=120,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        num_partitions_row=4, num_partitions_col=4, num_wood_zones=4, num_stone_zones=4, num_wood_and_stone_zones=3,
        starting_agent_coin=10, starting_wood_coverage=0.10, starting_stone_coverage=0.10,
        wood_regen_weight=0.05, stone_regen_weight=0.03),
    "quadrant": dict(
        scenario_name="quadrant/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="lognormal", payment_max_skill_multiplier=2)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict(skill_dist="lognormal"))],
        n_agents=6, world_size=[21, 21], episode_length=120,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        starting_agent_coin=10, starting_wood_coverage=0.08, starting_stone_coverage=0.08,
        wood_regen_weight=0.04, stone_regen_weight=0.04, checker_source_blocks=True),
    # split_layout draws its skill table inside the constructor: `seed` is a constructor kwarg here
    "split_layout": dict(
        scenario_name="split_layout/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict())],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=10, seed=77,
        skill_rank_of_top_agents=[0, 2], n_agents=5, world_size=[25, 25], episode_length=120,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # single-action planner (multi_action_mode_planner=False): one index over [NO-OP] ++ every bracket's rates
    "tax_single_planner": dict(
        scenario_name="layout_from_file/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                    ("Gather", dict(skill_dist="pareto")),
                    ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=5, rate_disc=0.05,
                                                tax_model="model_wrapper"))],
        env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5,
        fixed_four_skill_and_loc=False, n_agents=5, world_size=[25, 25], episode_length=150,
        multi_action_mode_agents=False, multi_action_mode_planner=False,
        flatten_observations=True, flatten_masks=True),
    # regen_halfwidth > 0: the respawn probability of a source cell counts the source cells in its d x d window
    "uniform_halfwidth": dict(
        scenario_name="uniform/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                    ("Gather", dict(skill_dist="pareto"))],
        n_agents=5, world_size=[18, 18], episode_length=150,
        multi_action_mode_agents=False, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True,
        starting_agent_coin=10, starting_wood_coverage=0.12, starting_stone_coverage=0.12,
        wood_regen_halfwidth=2, wood_regen_weight=0.6, stone_regen_halfwidth=1, stone_regen_weight=0.4),
}

# a FIXED tax schedule under a tax_annealing_schedule: the schedule is clipped by the annealed maximum of each episode
# (0.5, 0.75, 1.0 x rate_max over the first three episodes)
CONFIGS["us_federal_annealed_reset"] = dict(
    scenario_name="layout_from_file/simple_wood_and_stone",
    components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                ("Gather", dict(skill_dist="pareto")),
                ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=5, rate_max=0.3,
                                            tax_model="us-federal-single-filer-2018-scaled",
                                            tax_annealing_schedule=[-2, 0.25]))],
    env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=30,
    fixed_four_skill_and_loc=False, n_agents=5, world_size=[25, 25], episode_length=20,
    multi_action_mode_agents=False, multi_action_mode_planner=True, flatten_observations=True, flatten_masks=True)

# device-side reset of split_layout (rank table per replica, drawn in the constructor: `seed` is a constructor kwarg)
CONFIGS["split_reset"] = dict(
    scenario_name="split_layout/simple_wood_and_stone",
    components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                ("Gather", dict(skill_dist="pareto"))],
    env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5, seed=77,
    skill_rank_of_top_agents=[0, 3], n_agents=5, world_size=[25, 25], episode_length=20,
    multi_action_mode_agents=False, multi_action_mode_planner=True, flatten_observations=True, flatten_masks=True)

# device-side reset with lognormal skills: numpy's legacy Gaussian cache carries over between components and resets
CONFIGS["lognormal_reset"] = dict(
    scenario_name="layout_from_file/simple_wood_and_stone",
    # 5 agents, Build alone lognormal: 5 Gaussians per reset, so every other reset starts with the cached second variate
    # of the previous reset's last pair; Gather's Pareto draws sit in between
    components=[("Build", dict(skill_dist="lognormal", payment_max_skill_multiplier=3)),
                ("ContinuousDoubleAuction", dict(max_num_orders=3, order_duration=7)),
                ("Gather", dict(skill_dist="pareto"))],
    env_layout_file="quadrant_25x25_20each_30clump.txt", starting_agent_coin=5,
    fixed_four_skill_and_loc=False, n_agents=5, world_size=[25, 25], episode_length=22,
    multi_action_mode_agents=False, multi_action_mode_planner=True, flatten_observations=True, flatten_masks=True)

# full_observability=True: agents get the whole map (no window, no loc scalars); p<i> carry only the tax entries
CONFIGS["full_obs_tax"] = dict(
    scenario_name="layout_from_file/simple_wood_and_stone",
    components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                ("ContinuousDoubleAuction", dict(max_num_orders=3)),
                ("Gather", dict(skill_dist="pareto")),
                ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=10, rate_disc=0.05,
                                            tax_model="model_wrapper"))],
    env_layout_file="env-pure_and_mixed-15x15.txt", starting_agent_coin=8, full_observability=True,
    n_agents=5, world_size=[15, 15], episode_length=100,
    multi_action_mode_agents=False, multi_action_mode_planner=True, flatten_observations=True, flatten_masks=True)

# Edge-of-range configurations (tests/test_edge_configs.py): smallest / largest sizes and degenerate options.  They are
# checked live against the imported reference in the build container, and oracle <-> device code everywhere.
_LFF = "layout_from_file/simple_wood_and_stone"
_BASE = dict(multi_action_mode_agents=False, multi_action_mode_planner=True, flatten_observations=True,
             flatten_masks=True)
EDGE_CONFIGS = {
    # two agents (the minimum), view radius 0 (1x1 window), K = 1 order, one price level above zero
    "two_agents_w0": dict(
        scenario_name=_LFF, components=[("Build", {}), ("ContinuousDoubleAuction", dict(max_num_orders=1, max_bid_ask=1,
                                                                                       order_duration=1)),
                                        ("Gather", {})],
        env_layout_file="env-pure_and_mixed-15x15.txt", n_agents=2, world_size=[15, 15], episode_length=40,
        mobile_agent_observation_range=0, starting_agent_coin=3, **_BASE),
    # window wider than the world (radius 9 on 15x15), planner without spatial info, no auction component
    "wide_window_no_cda": dict(
        scenario_name=_LFF, components=[("Gather", dict(skill_dist="pareto")), ("Build", dict(skill_dist="lognormal"))],
        env_layout_file="env-pure_and_mixed-15x15.txt", n_agents=3, world_size=[15, 15], episode_length=40,
        mobile_agent_observation_range=9, planner_gets_spatial_info=False, **_BASE),
    # gather only; inventory scaling off; auto energy warm-up; full regeneration probability
    "gather_only_unscaled": dict(
        scenario_name=_LFF, components=[("Gather", dict(move_labor=0.5, collect_labor=2.0))],
        env_layout_file="quadrant_25x25_20each_30clump.txt", n_agents=7, world_size=[25, 25], episode_length=40,
        allow_observation_scaling=False, resource_regen_prob=1.0, energy_warmup_constant=3,
        energy_warmup_method="auto", **_BASE),
    # auction + gather (the reference cannot build its observations without a Gather or Build component),
    # multi-action agents, long-lived orders, 32 price levels (the ABI maximum)
    "cda_wide_prices": dict(
        scenario_name=_LFF, components=[("ContinuousDoubleAuction", dict(max_num_orders=4, max_bid_ask=31,
                                                                         order_duration=200, order_labor=0.0)),
                                        ("Gather", {})],
        env_layout_file="env-pure_and_mixed-15x15.txt", n_agents=4, world_size=[15, 15], episode_length=40,
        starting_agent_coin=50, multi_action_mode_agents=True, multi_action_mode_planner=True,
        flatten_observations=True, flatten_masks=True),
    # tax every step (period 1), taxes disabled, the other two social welfare functions
    "tax_period_one": dict(
        scenario_name=_LFF, components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=2)),
                                        ("Gather", {}),
                                        ("PeriodicBracketTax", dict(period=1, bracket_spacing="linear", n_brackets=3,
                                                                    top_bracket_cutoff=30, rate_disc=0.25,
                                                                    tax_model="model_wrapper"))],
        env_layout_file="env-pure_and_mixed-15x15.txt", n_agents=3, world_size=[15, 15], episode_length=40,
        starting_agent_coin=20, planner_reward_type="inv_income_weighted_utility", **_BASE),
    "tax_disabled_log_brackets": dict(
        scenario_name=_LFF, components=[("Gather", {}), ("Build", {}),
                                        ("PeriodicBracketTax", dict(period=7, bracket_spacing="log", n_brackets=5,
                                                                    top_bracket_cutoff=100, disable_taxes=True))],
        env_layout_file="env-pure_and_mixed-15x15.txt", n_agents=3, world_size=[15, 15], episode_length=40,
        starting_agent_coin=20, planner_reward_type="inv_income_weighted_coin_endowments",
        mixing_weight_gini_vs_coin=0.5, isoelastic_eta=0.0, **_BASE),
    # eta just below 1 (the reference's eta == 1 branch raises, rewards.py:38), fixed bracket rates, tax every 2nd step
    "eta_high_fixed_rates": dict(
        scenario_name=_LFF, components=[("Build", {}), ("ContinuousDoubleAuction", dict(max_num_orders=2)),
                                        ("Gather", {}),
                                        ("PeriodicBracketTax", dict(period=2, bracket_spacing="us-federal",
                                                                    tax_model="fixed-bracket-rates",
                                                                    fixed_bracket_rates=[0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6]))],
        env_layout_file="env-pure_and_mixed-15x15.txt", n_agents=4, world_size=[15, 15], episode_length=40,
        starting_agent_coin=15, isoelastic_eta=0.99, **_BASE),
    # full observability without a tax component: the planner has no p<i> vectors at all; non-square world
    "full_obs_plain": dict(
        scenario_name="uniform/simple_wood_and_stone",
        components=[("Build", {}), ("ContinuousDoubleAuction", dict(max_num_orders=2)), ("Gather", {})],
        n_agents=3, world_size=[9, 14], episode_length=40, starting_agent_coin=5, full_observability=True,
        starting_wood_coverage=0.10, starting_stone_coverage=0.10, **_BASE),
    # non-square world (uniform family), 33 agents: one more than a warp, both gini branches straddled
    "nonsquare_33_agents": dict(
        scenario_name="uniform/simple_wood_and_stone",
        components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                    ("ContinuousDoubleAuction", dict(max_num_orders=7)), ("Gather", dict(skill_dist="pareto"))],
        n_agents=33, world_size=[12, 37], episode_length=30, starting_agent_coin=30,
        starting_wood_coverage=0.10, starting_stone_coverage=0.10, **_BASE),
}

# device-side reset of the dynamic-layout scenarios: every reset draws a new clumped layout (np.random.rand thinning, then
# randn + convolve2d growth) and places the agents in a random order (dynamic_layout.py:313-429)
CONFIGS["uniform_reset"] = dict(
    scenario_name="uniform/simple_wood_and_stone",
    components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                ("Gather", dict(skill_dist="pareto"))],
    n_agents=5, world_size=[18, 22], episode_length=30,
    multi_action_mode_agents=False, multi_action_mode_planner=True,
    flatten_observations=True, flatten_masks=True,
    starting_agent_coin=10, starting_wood_coverage=0.12, starting_stone_coverage=0.10,
    wood_regen_weight=0.3, stone_regen_weight=0.2)
# Quadrant: normalised probability maps, water lines, checkered sources, lognormal skills (the Gaussian cache is shared
# between the layout generator and the skill draws)
CONFIGS["quadrant_reset"] = dict(
    scenario_name="quadrant/simple_wood_and_stone",
    components=[("Build", dict(skill_dist="lognormal", payment_max_skill_multiplier=2)),
                ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                ("Gather", dict(skill_dist="lognormal"))],
    n_agents=6, world_size=[21, 21], episode_length=28,
    multi_action_mode_agents=False, multi_action_mode_planner=True,
    flatten_observations=True, flatten_masks=True,
    starting_agent_coin=10, starting_wood_coverage=0.08, starting_stone_coverage=0.08,
    wood_regen_weight=0.04, stone_regen_weight=0.04, checker_source_blocks=True)

# MultiZone: the region -> zone-type vector is re-shuffled (np.random.shuffle) before every layout
CONFIGS["multi_zone_reset"] = dict(
    scenario_name="multi_zone/simple_wood_and_stone",
    components=[("Build", dict(skill_dist="pareto", payment_max_skill_multiplier=3)),
                ("ContinuousDoubleAuction", dict(max_num_orders=5)),
                ("Gather", dict(skill_dist="none"))],
    n_agents=4, world_size=[20, 17], episode_length=26,
    multi_action_mode_agents=False, multi_action_mode_planner=True,
    flatten_observations=True, flatten_masks=True,
    starting_agent_coin=5, starting_wood_coverage=0.1, starting_stone_coverage=0.1,
    num_partitions_row=4, num_partitions_col=3, num_wood_zones=3, num_stone_zones=3, num_wood_and_stone_zones=2)

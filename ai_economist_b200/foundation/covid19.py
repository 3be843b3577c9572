"""COVID-19 and economy scenario (BASELINE config 4): host-side mirror.

reference: scenarios/covid19/covid19_env.py (CovidAndEconomyEnvironment), components/covid19_components.py
(ControlUSStateOpenCloseStatus, FederalGovernmentSubsidy, VaccinationCampaign).

`build_covid_params` resolves the scenario + component kwargs and the fitted-parameter / real-world data files
into plain arrays: everything the per-timestep path needs.  The derived constants (daily production per
worker, maximum productivity, reward norms and weightages) follow the reference constructor expression by
expression, including its float32 / int32 / float64 dtypes (covid19_env.py:96-385, 1517-1625), because the
step arithmetic downstream is dtype-sensitive.
"""
import json
import os
from datetime import datetime

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "covid19_data")
F32, I32 = np.float32, np.int32


def _softplus(x, beta=1, threshold=20):  # covid19_env.py:1398-1406
    return 1 / beta * np.log(1 + np.exp(beta * x)) * (beta * x <= threshold) + x * (beta * x > threshold)


def build_covid_params(episode_length=540, start_date="2020-03-22", pop_between_age_18_65=0.6,
                       infection_too_sick_to_work_rate=0.1, risk_free_interest_rate=0.03,
                       economic_reward_crra_eta=2, health_priority_scaling_agents=1,
                       health_priority_scaling_planner=1, reward_normalization_factor=1,
                       path_to_data_and_fitted_params="", action_cooldown_period=28, n_stringency_levels=10,
                       subsidy_interval=90, num_subsidy_levels=20, max_annual_subsidy_per_person=20000,
                       daily_vaccines_per_million_people=4500, delivery_interval=1,
                       vaccine_delivery_start_date="2020-12-22", allow_observation_scaling=True):
    path = path_to_data_and_fitted_params or DATA_DIR
    rw = dict(np.load(os.path.join(path, "real_world_data.npz")))
    mc = json.load(open(os.path.join(path, "model_constants.json")))
    fp = json.load(open(os.path.join(path, "fitted_params.json")))
    p = {}
    date_format = mc["DATE_FORMAT"]
    pop = I32(mc["US_STATE_POPULATION"])
    us_pop = I32(mc["US_POPULATION"])
    S = len(pop)
    p["n_states"] = S
    p["episode_length"] = int(episode_length)
    p["population"] = pop
    p["num_stringency_levels"] = int(mc["NUM_STRINGENCY_LEVELS"])
    assert int(n_stringency_levels) == p["num_stringency_levels"], \
        "For the given model fit, the number of stringency levels must be {}".format(p["num_stringency_levels"])
    p["death_rate"] = F32(mc["SIR_MORTALITY"])
    p["gamma"] = F32(mc["SIR_GAMMA"])
    gdp_per_capita = F32(mc["GDP_PER_CAPITA"])
    policy_start = datetime.strptime(fp["POLICY_START_DATE"], date_format)
    start = datetime.strptime(start_date, date_format)
    assert start >= policy_start
    sdi = (start - policy_start).days
    assert 0 <= sdi < len(rw["policy"])
    p["start_date_index"] = sdi
    p["value_of_life"] = I32(fp["VALUE_OF_LIFE"])
    p["beta_delay"] = int(fp["BETA_DELAY"])
    p["beta_slopes"] = np.array(fp["BETA_SLOPES"], dtype=F32)
    p["beta_intercepts"] = np.array(fp["BETA_INTERCEPTS"], dtype=F32)
    for k in ["MIN_MARGINAL_AGENT_HEALTH_INDEX", "MAX_MARGINAL_AGENT_HEALTH_INDEX",
              "MIN_MARGINAL_AGENT_ECONOMIC_INDEX", "MAX_MARGINAL_AGENT_ECONOMIC_INDEX"]:
        p[k.lower()] = np.array(fp[k], dtype=F32)
    for k in ["MIN_MARGINAL_PLANNER_HEALTH_INDEX", "MAX_MARGINAL_PLANNER_HEALTH_INDEX",
              "MIN_MARGINAL_PLANNER_ECONOMIC_INDEX", "MAX_MARGINAL_PLANNER_ECONOMIC_INDEX"]:
        p[k.lower()] = F32(fp[k])
    inferred_w_agent = np.array(fp["INFERRED_WEIGHTAGE_ON_AGENT_HEALTH_INDEX"], dtype=F32)
    inferred_w_planner = F32(fp["INFERRED_WEIGHTAGE_ON_PLANNER_HEALTH_INDEX"])
    L = int(fp["FILTER_LEN"])
    p["filter_len"] = L
    conv_lambdas = np.array(fp["CONV_LAMBDAS"], dtype=F32)
    F = len(conv_lambdas)
    p["num_filters"] = F
    p["unemployment_bias"] = np.array(fp["UNEMPLOYMENT_BIAS"], dtype=F32)
    gw = np.array(fp["GROUPED_CONVOLUTIONAL_FILTER_WEIGHTS"], dtype=F32)
    p["conv_weights"] = gw.reshape(S, F)                                      # float32 [S, F]
    f_ts = np.tile(np.flip(np.arange(L), (0,))[None, None], (1, F, 1)).astype(F32)
    p["conv_filters"] = np.exp(-f_ts / conv_lambdas[None, :, None])[0]        # float32 [F, L]
    p["risk_free_interest_rate"] = F32(risk_free_interest_rate)
    # unemployment at closure policy "all ones" (timestep 0: zero stringency deltas), covid19_env.py:260-262, 1407-1441
    delta0 = np.zeros((L, S))
    w_x = delta0[None].transpose(2, 0, 1) * np.repeat(p["conv_weights"][:, :, None], L, axis=-1)
    excess = _softplus(np.sum(w_x * p["conv_filters"][None], axis=(1, 2)), beta=1)
    unemployed_level_1 = (excess + p["unemployment_bias"]) * pop / 100
    workforce = (us_pop * pop_between_age_18_65 - np.sum(unemployed_level_1)).astype(I32)
    workers_per_capita = (workforce / us_pop).astype(F32)
    gdp_per_worker = (gdp_per_capita / workers_per_capita).astype(F32)
    p["daily_production_per_worker"] = (gdp_per_worker / 365).astype(F32)
    p["infection_too_sick_to_work_rate"] = F32(infection_too_sick_to_work_rate)
    p["pop_between_age_18_65"] = F32(pop_between_age_18_65)
    # economy_step with nobody sick, dead (covid19_env.py:283-292, 1444-1475)
    zero = np.zeros(S, dtype=I32)
    incapacitated = (p["infection_too_sick_to_work_rate"] * zero) + zero
    cant_work = (incapacitated * p["pop_between_age_18_65"]) + unemployed_level_1
    can_work = np.maximum(0, pop * p["pop_between_age_18_65"] - cant_work)
    p["maximum_productivity"] = (can_work * p["daily_production_per_worker"]).astype(F32)
    p["crra_eta"] = F32(economic_reward_crra_eta)
    p["agents_health_norm"] = p["maximum_productivity"] * 365
    p["planner_health_norm"] = np.sum(p["agents_health_norm"])
    p["agents_economic_norm"] = p["maximum_productivity"] * 365
    p["planner_economic_norm"] = np.sum(p["agents_economic_norm"])

    def scale(h, alphas):  # covid19_env.py:312-321
        z = alphas / (1 - alphas)
        sz = h * z
        return sz / (1 + sz)

    p["w_agent_health"] = scale(health_priority_scaling_agents, inferred_w_agent)
    p["w_agent_econ"] = 1 - p["w_agent_health"]
    p["w_planner_health"] = scale(health_priority_scaling_planner, inferred_w_planner)
    p["w_planner_econ"] = 1 - p["w_planner_health"]
    p["reward_normalization_factor"] = reward_normalization_factor
    # components
    p["action_cooldown_period"] = int(action_cooldown_period)
    p["subsidy_interval"] = int(subsidy_interval)
    p["num_subsidy_levels"] = int(num_subsidy_levels)
    p["max_daily_subsidy_per_state"] = pop * float(max_annual_subsidy_per_person) / 365   # float64 [S]
    vstart = datetime.strptime(vaccine_delivery_start_date, "%Y-%m-%d")
    p["time_when_vaccine_delivery_begins"] = (vstart - start).days
    p["delivery_interval"] = int(delivery_interval)
    daily_vaccines = (pop / 1e6) * int(daily_vaccines_per_million_people)
    p["num_vaccines_per_delivery"] = np.array(np.floor(int(delivery_interval) * daily_vaccines), dtype=I32)
    t_first = int(p["time_when_vaccine_delivery_begins"])
    while t_first % p["delivery_interval"] != 0:
        t_first += 1
    p["t_first_delivery"] = t_first
    p["time_scale"] = float(episode_length) if allow_observation_scaling else 1.0
    # real-world series used at reset and for lagged stringency before the episode has enough history
    p["rw_policy"] = np.asarray(rw["policy"])                  # int64 [D, S]
    # initial state (covid19_env.py:1175-1215)
    p["init"] = dict(
        susceptible=rw["susceptible"][sdi], infected=rw["infected"][sdi], recovered=rw["recovered"][sdi],
        deaths=rw["recovered"][sdi] * p["death_rate"], unemployed=rw["unemployed"][sdi], vaccinated=rw["vaccinated"][sdi],
        stringency=rw["policy"][sdi],
        stringency_history=np.pad(rw["policy"][: sdi + 1], [(L, 0), (0, 0)], constant_values=1)[-(L + 1):],
    )
    return p


# ---------------------------------------------------------------------------------------------------------
# Registry entries + the batched env (Gym-style surface of CovidAndEconomyEnvironment)
# ---------------------------------------------------------------------------------------------------------
from .components import BaseComponent, component_registry  # noqa: E402
from .scenarios import BaseScenario, scenario_registry  # noqa: E402


@component_registry.add
class ControlUSStateOpenCloseStatus(BaseComponent):
    """reference: components/covid19_components.py:43-96"""
    name = "ControlUSStateOpenCloseStatus"
    agent_subclasses = ["BasicMobileAgent"]

    def __init__(self, *a, n_stringency_levels=10, action_cooldown_period=28, **k):
        super().__init__(*a, **k)
        self.n_stringency_levels = int(n_stringency_levels)
        assert self.n_stringency_levels >= 2
        self.action_cooldown_period = action_cooldown_period

    def get_n_actions(self, agent_cls_name):
        return self.n_stringency_levels if agent_cls_name == "BasicMobileAgent" else None

    def spec_fields(self):
        return dict(n_stringency_levels=self.n_stringency_levels, action_cooldown_period=self.action_cooldown_period)


@component_registry.add
class FederalGovernmentSubsidy(BaseComponent):
    """reference: components/covid19_components.py:241-314"""
    name = "FederalGovernmentSubsidy"
    agent_subclasses = ["BasicPlanner"]

    def __init__(self, *a, subsidy_interval=90, num_subsidy_levels=20, max_annual_subsidy_per_person=20000, **k):
        super().__init__(*a, **k)
        self.subsidy_interval = int(subsidy_interval)
        assert self.subsidy_interval >= 1
        self.num_subsidy_levels = int(num_subsidy_levels)
        assert self.num_subsidy_levels >= 1
        self.max_annual_subsidy_per_person = float(max_annual_subsidy_per_person)
        assert self.max_annual_subsidy_per_person >= 0

    def get_n_actions(self, agent_cls_name):
        return self.num_subsidy_levels if agent_cls_name == "BasicPlanner" else None

    def spec_fields(self):
        return dict(subsidy_interval=self.subsidy_interval, num_subsidy_levels=self.num_subsidy_levels,
                    max_annual_subsidy_per_person=self.max_annual_subsidy_per_person)


@component_registry.add
class VaccinationCampaign(BaseComponent):
    """reference: components/covid19_components.py:465-560 (passive component)"""
    name = "VaccinationCampaign"
    agent_subclasses = ["BasicMobileAgent"]

    def __init__(self, *a, daily_vaccines_per_million_people=4500, delivery_interval=1,
                 vaccine_delivery_start_date="2020-12-22", observe_rate=False, **k):
        super().__init__(*a, **k)
        self.daily_vaccines_per_million_people = int(daily_vaccines_per_million_people)
        assert 0 <= self.daily_vaccines_per_million_people <= 1e6
        self.delivery_interval = int(delivery_interval)
        assert 1 <= self.delivery_interval <= 5000
        self.vaccine_delivery_start_date = vaccine_delivery_start_date
        self.observe_rate = bool(observe_rate)   # adds the `next_vaccination_rate` observation (covid19_components.py:659-661)

    def spec_fields(self):
        return dict(daily_vaccines_per_million_people=self.daily_vaccines_per_million_people,
                    delivery_interval=self.delivery_interval, vaccine_delivery_start_date=self.vaccine_delivery_start_date)


class CovidBatchedEnv:
    """reset()/step() over E replicas of the COVID-19 + economy simulation on one GPU.

    Observations come collated exactly like the reference's mandatory `collate_agent_step_and_reset_data=True`
    layout: obs["a"][key] has the agent (US state) axis last, obs["p"][key] is the planner's; every tensor has
    a leading env axis.  Values are float32 (the reference's own CUDA path also uses float32 placeholders).
    """

    def __init__(self, components=None, n_agents=51, world_size=(1, 1), episode_length=540,
                 multi_action_mode_agents=False, multi_action_mode_planner=False, flatten_observations=False,
                 flatten_masks=True, collate_agent_step_and_reset_data=True, allow_observation_scaling=True,
                 use_real_world_data=False, use_real_world_policies=False, n_envs=1, device="cuda:0", auto_reset=True,
                 stepper_factory=None, seed=None, seeds=None, dense_log_frequency=None, world_dense_log_frequency=50,
                 **scenario_kwargs):
        assert collate_agent_step_and_reset_data, \
            "The env. config 'collate_agent_step_and_reset_data' should be set to True."
        if use_real_world_data or use_real_world_policies:
            raise NotImplementedError("replaying real-world data/policies is a host-side mode outside the GPU hot path")
        if multi_action_mode_agents or multi_action_mode_planner or not flatten_masks:
            raise NotImplementedError("the COVID path uses single-action agents/planner and flattened masks")
        kw = dict(episode_length=episode_length, allow_observation_scaling=allow_observation_scaling)
        self._components = []
        for spec in components or []:
            cname, ckw = (spec if isinstance(spec, (tuple, list)) else list(spec.items())[0])
            comp = component_registry.get(cname)(n_agents, episode_length, **ckw)
            self._components.append(comp)
            kw.update(comp.spec_fields())
        names = [c.name for c in self._components]
        assert names == ["ControlUSStateOpenCloseStatus", "FederalGovernmentSubsidy", "VaccinationCampaign"], \
            "the GPU path implements the reference's three COVID components in their canonical order"
        for k in ["start_date", "pop_between_age_18_65", "infection_too_sick_to_work_rate", "risk_free_interest_rate",
                  "economic_reward_crra_eta", "health_priority_scaling_agents", "health_priority_scaling_planner",
                  "reward_normalization_factor", "path_to_data_and_fitted_params"]:
            if k in scenario_kwargs:
                kw[k] = scenario_kwargs.pop(k)
        assert not scenario_kwargs, "unknown kwargs: %s" % list(scenario_kwargs)
        self.params = build_covid_params(**kw)
        assert n_agents == self.params["n_states"], \
            "n_agents should be set to the number of US states, i.e., {}.".format(self.params["n_states"])
        self.name = "CovidAndEconomySimulation"
        self.n_agents, self.n_envs = n_agents, int(n_envs)
        self.num_agents = n_agents + 1
        self.world_size = list(world_size)
        self._episode_length = int(episode_length)
        if stepper_factory is None:
            from ..covid_stepper import CudaCovidStepper
            self._stepper = CudaCovidStepper(self.params, self.n_envs, device=device, auto_reset=auto_reset)
        else:
            self._stepper = stepper_factory(self.params, self.n_envs, auto_reset)
        self._loaded = False
        self._build_views()
        self._build_rate_observation()

    @property
    def episode_length(self):
        return self._episode_length

    @property
    def components(self):
        return self._components

    @property
    def stepper(self):
        return self._stepper

    def seed(self, seed):  # the scenario is deterministic
        return None

    @property
    def action_buffers(self):
        """(agent actions [E, S], planner actions [E]): write in place and call step(env.action_buffers) for zero copies."""
        return self._stepper.buf["actions_agent"], self._stepper.buf["actions_planner"]

    def _build_views(self):
        b = self._stepper.buf
        S = self.n_agents
        sc = b["obs_scalars"]
        exp = (lambda v: v[:, None].expand(-1, S)) if hasattr(sc, "expand") else (lambda v: np.repeat(v[:, None], S, axis=1))
        self.obs = {
            "a": {"world-agent_state": b["obs_agent_state"], "world-agent_postsubsidy_productivity": b["obs_postsubsidy"],
                  "world-lagged_stringency_level": b["obs_lagged_stringency"], "time": exp(sc[:, 0]),
                  "ControlUSStateOpenCloseStatus-agent_policy_indicators": b["obs_policy_indicators"],
                  "FederalGovernmentSubsidy-t_until_next_subsidy": exp(sc[:, 1]),
                  "FederalGovernmentSubsidy-current_subsidy_level": exp(sc[:, 2]),
                  "VaccinationCampaign-t_until_next_vaccines": exp(sc[:, 3]), "action_mask": b["mask_agent"]},
            "p": {"world-agent_state": b["obs_agent_state"], "world-agent_postsubsidy_productivity": b["obs_postsubsidy"],
                  "world-lagged_stringency_level": b["obs_lagged_stringency"], "time": sc[:, 0:1],
                  "ControlUSStateOpenCloseStatus-agent_policy_indicators": b["obs_policy_indicators"],
                  "FederalGovernmentSubsidy-t_until_next_subsidy": sc[:, 1],
                  "FederalGovernmentSubsidy-current_subsidy_level": sc[:, 2],
                  "VaccinationCampaign-t_until_next_vaccines": sc[:, 3], "action_mask": b["mask_planner"]},
        }
        self.rew = {"a": b["reward_agent"], "p": b["reward_planner"]}
        self.done = {"__all__": b["done"]}
        self.info = {"a": {}, "p": {}}

    def _build_rate_observation(self):
        """VaccinationCampaign(observe_rate=True) (covid19_components.py:629-661): the vaccination rate of the NEXT timestep, 0
        until deliveries have begun.  It depends on the timestep alone, so it is a table lookup by the replica's timestep after
        every step / reset (the reference's own CUDA path does not emit it; no kernel involved)."""
        vac = self._components[2]
        self._rate_tab = None
        if not vac.observe_rate:
            return
        T, t_first = self._episode_length, int(self.params["t_first_delivery"])
        rate = np.float32(vac.daily_vaccines_per_million_people / 1e6)
        tab = np.array([np.float32(0.0) if t + 1 <= t_first else rate for t in range(T + 1)], np.float32)
        sc = self._stepper.buf["obs_scalars"]
        E, S = self.n_envs, self.n_agents
        if isinstance(sc, np.ndarray):
            self._rate_tab, self._rate_a, self._rate_p = tab, np.zeros((E, S), np.float32), np.zeros(E, np.float32)
        else:
            import torch
            self._rate_tab = torch.as_tensor(tab, device=sc.device)
            self._rate_a, self._rate_p = torch.zeros((E, S), dtype=torch.float32, device=sc.device), torch.zeros(E, dtype=torch.float32, device=sc.device)
        self.obs["a"]["VaccinationCampaign-next_vaccination_rate"] = self._rate_a
        self.obs["p"]["VaccinationCampaign-next_vaccination_rate"] = self._rate_p

    def _refresh_rate_observation(self):
        if self._rate_tab is None:
            return
        sc = self._stepper.buf["obs_scalars"]   # [:, 0] = timestep / time_scale (episode_length under observation scaling)
        scale = float(self.params["time_scale"])
        if isinstance(sc, np.ndarray):
            t = np.rint(sc[:, 0].astype(np.float64) * scale).astype(np.int64)
            self._rate_p[...] = self._rate_tab[t]
            self._rate_a[...] = self._rate_p[:, None]
        else:
            import torch
            t = torch.round(sc[:, 0].double() * scale).long()
            self._rate_p.copy_(self._rate_tab[t])
            self._rate_a.copy_(self._rate_p[:, None].expand_as(self._rate_a))

    def reset(self):
        self._stepper.reset()
        self._loaded = True
        self._refresh_rate_observation()
        return self.obs

    def step(self, actions=None):
        """actions: None | {"a": int tensor [E, 51], "p": int tensor [E]} | (agent_actions, planner_actions)."""
        assert self._loaded, "call reset() first"
        ba, bp = self._stepper.buf["actions_agent"], self._stepper.buf["actions_planner"]
        if actions is None:
            ba[...] = 0
            bp[...] = 0
        else:
            a, p = (actions.get("a"), actions.get("p")) if isinstance(actions, dict) else actions
            for buf, v in ((ba, a), (bp, p)):
                if v is None:
                    buf[...] = 0
                elif v is not buf:
                    if isinstance(buf, np.ndarray):
                        buf[...] = np.asarray(v, dtype=buf.dtype).reshape(buf.shape)
                    else:
                        import torch
                        buf[...] = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v,
                                                   device=buf.device).to(buf.dtype).reshape(buf.shape)
        self._stepper.step()
        self._refresh_rate_observation()
        return self.obs, self.rew, self.done, self.info


@scenario_registry.add
class CovidAndEconomySimulation(BaseScenario):
    """Registry entry; make_env_instance routes this name to CovidBatchedEnv."""
    name = "CovidAndEconomySimulation"
    env_class = CovidBatchedEnv

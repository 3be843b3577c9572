"""TEST / BASELINE INFRASTRUCTURE: the reference's own COVID-19 CUDA kernels (oracle/_ref/libref_covid_cuda.so, built
by oracle/build_ref_covid.py from the sources under /root/reference) driven the way the reference's WarpDrive wrapper
drives them: one block of n_agents (51 states + planner) threads per env replica, five launches per env.step().

Only tests/ and bench.py's baseline leg use this; the product never does.  The data pushed to the device is the
reference's data dictionary (scenarios/covid19/covid19_env.py:388-636; components/covid19_components.py:110-136,
327-352, 562-584), built here from ai_economist_b200.foundation.covid19.build_covid_params (the same constants, already
pinned against the reference's Python path by the golden trace).  Arrays the reference marks
`save_copy_and_apply_at_reset` are restored by reset().
"""
import ctypes as C
import os

import numpy as np

from . import build_ref_covid

F32, I32 = np.float32, np.int32

_PTR_FIELDS = [
    "susceptible", "infected", "recovered", "deaths", "vaccinated", "unemployed", "subsidy", "productivity",
    "postsubsidy_productivity", "stringency_level", "subsidy_level", "beta", "incapacitated", "cant_work",
    "num_people_that_can_work", "delta_stringency_level", "signal", "action_in_cooldown_until",
    "num_vaccines_available_t", "timestep", "done", "default_agent_action_mask", "no_op_agent_action_mask",
    "default_planner_action_mask", "no_op_planner_action_mask", "max_daily_subsidy_per_state",
    "num_vaccines_per_delivery", "us_state_population", "real_world_stringency_policy_history", "beta_slopes",
    "beta_intercepts", "grouped_convolutional_filter_weights", "unemp_conv_filters", "unemployment_bias",
    "maximum_productivity", "min_marginal_agent_health_index", "max_marginal_agent_health_index",
    "min_marginal_agent_economic_index", "max_marginal_agent_economic_index",
    "weightage_on_marginal_agent_health_index", "weightage_on_marginal_agent_economic_index", "agents_health_norm",
    "agents_economic_norm", "actions_a", "actions_p", "obs_a_policy_indicators", "obs_a_action_mask",
    "obs_p_policy_indicators", "obs_a_t_until_next_subsidy", "obs_a_current_subsidy_level", "obs_p_t_until_next_subsidy",
    "obs_p_current_subsidy_level", "obs_p_action_mask", "obs_a_t_until_next_vaccines", "obs_p_t_until_next_vaccines",
    "obs_a_agent_state", "obs_a_postsubsidy", "obs_a_lagged", "obs_a_time", "obs_p_agent_state", "obs_p_postsubsidy",
    "obs_p_lagged", "obs_p_time", "rewards_a", "rewards_p"]
_INT_FIELDS = ["action_cooldown_period", "num_stringency_levels", "subsidy_interval", "num_subsidy_levels",
               "delivery_interval", "time_when_vaccine_delivery_begins", "beta_delay", "filter_len", "num_filters",
               "num_days_in_an_year", "value_of_life", "n_agents", "episode_length", "n_envs"]
_FLOAT_FIELDS = ["gamma", "death_rate", "infection_too_sick_to_work_rate", "population_between_age_18_65",
                 "daily_production_per_worker", "risk_free_interest_rate", "economic_reward_crra_eta",
                 "min_marginal_planner_health_index", "max_marginal_planner_health_index",
                 "min_marginal_planner_economic_index", "max_marginal_planner_economic_index",
                 "weightage_on_marginal_planner_health_index", "weightage_on_marginal_planner_economic_index",
                 "planner_health_norm", "planner_economic_norm"]


class RefCovidArgs(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in _PTR_FIELDS] + [(n, C.c_int) for n in _INT_FIELDS] +
                [(n, C.c_float) for n in _FLOAT_FIELDS])


def load():
    path = build_ref_covid.build()
    if path is None or not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_covid_step.argtypes = [C.POINTER(RefCovidArgs), C.c_void_p]
    lib.ref_covid_step.restype = C.c_int
    lib.ref_covid_args_size.restype = C.c_int
    assert lib.ref_covid_args_size() == C.sizeof(RefCovidArgs), "RefCovidArgs layout mismatch"
    return lib


class RefCovidCuda:
    """E replicas of the reference's GPU COVID env on `device`."""

    def __init__(self, p, n_envs, device="cuda:0"):
        import torch

        self.torch, self.p, self.E = torch, p, int(n_envs)
        self.device = torch.device(device)
        self.lib = load()
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libref_covid_cuda.so is not built (needs /root/reference at build time)")
        E, S, T, L, Fn = self.E, p["n_states"], p["episode_length"], p["filter_len"], p["num_filters"]
        NS, NL = p["num_stringency_levels"], p["num_subsidy_levels"]
        ini = p["init"]
        dev = self.device

        def const(a, dt):
            return torch.as_tensor(np.ascontiguousarray(np.asarray(a), dtype=dt), device=dev)

        def timedep(v0, dt):   # [E, T+1, S], index 0 = the state at reset
            a = np.zeros((T + 1, S), dt)
            if v0 is not None:
                a[0] = np.asarray(v0).astype(dt)
            return torch.as_tensor(a, device=dev).unsqueeze(0).repeat(E, 1, 1).contiguous()

        z = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)
        hist = np.asarray(ini["stringency_history"])
        sdi, bd = p["start_date_index"], p["beta_delay"]
        t = dict(
            susceptible=timedep(ini["susceptible"], F32), infected=timedep(ini["infected"], F32),
            recovered=timedep(ini["recovered"], F32), deaths=timedep(ini["deaths"], F32),
            vaccinated=timedep(ini["vaccinated"], F32), unemployed=timedep(ini["unemployed"], F32),
            subsidy=timedep(None, F32), productivity=timedep(None, F32), postsubsidy_productivity=timedep(None, F32),
            stringency_level=timedep(ini["stringency"], I32), subsidy_level=timedep(None, I32),
            beta=z(E, S), incapacitated=z(E, S), cant_work=z(E, S), num_people_that_can_work=z(E, S),
            delta_stringency_level=const(np.tile((hist[1:] - hist[:-1]).astype(I32)[None], (E, 1, 1)), I32),
            signal=z(E, S, Fn, L), action_in_cooldown_until=z(E, S, dt=torch.int32),
            num_vaccines_available_t=z(E, S, dt=torch.int32), timestep=z(E, dt=torch.int32), done=z(E, dt=torch.int32),
            default_agent_action_mask=const([1] * (NS + 1), I32), no_op_agent_action_mask=const([1] + [0] * NS, I32),
            default_planner_action_mask=const([1] * (NL + 1), I32), no_op_planner_action_mask=const([1] + [0] * NL, I32),
            max_daily_subsidy_per_state=const(p["max_daily_subsidy_per_state"], F32),
            num_vaccines_per_delivery=const(p["num_vaccines_per_delivery"], I32),
            us_state_population=const(p["population"], I32),
            real_world_stringency_policy_history=const(p["rw_policy"][sdi - bd + 1: sdi], I32),
            beta_slopes=const(p["beta_slopes"], F32), beta_intercepts=const(p["beta_intercepts"], F32),
            grouped_convolutional_filter_weights=const(p["conv_weights"], F32),
            unemp_conv_filters=const(p["conv_filters"], F32), unemployment_bias=const(p["unemployment_bias"], F32),
            maximum_productivity=const(p["maximum_productivity"], F32),
            min_marginal_agent_health_index=const(p["min_marginal_agent_health_index"], F32),
            max_marginal_agent_health_index=const(p["max_marginal_agent_health_index"], F32),
            min_marginal_agent_economic_index=const(p["min_marginal_agent_economic_index"], F32),
            max_marginal_agent_economic_index=const(p["max_marginal_agent_economic_index"], F32),
            weightage_on_marginal_agent_health_index=const(p["w_agent_health"], F32),
            weightage_on_marginal_agent_economic_index=const(p["w_agent_econ"], F32),
            agents_health_norm=const(p["agents_health_norm"], F32), agents_economic_norm=const(p["agents_economic_norm"], F32),
            actions_a=z(E, S, dt=torch.int32), actions_p=z(E, dt=torch.int32),
            obs_a_policy_indicators=z(E, S), obs_a_action_mask=z(E, NS + 1, S), obs_p_policy_indicators=z(E, S),
            obs_a_t_until_next_subsidy=z(E, S), obs_a_current_subsidy_level=z(E, S), obs_p_t_until_next_subsidy=z(E),
            obs_p_current_subsidy_level=z(E), obs_p_action_mask=z(E, NL + 1), obs_a_t_until_next_vaccines=z(E, S),
            obs_p_t_until_next_vaccines=z(E), obs_a_agent_state=z(E, 6, S), obs_a_postsubsidy=z(E, S), obs_a_lagged=z(E, S),
            obs_a_time=z(E, S), obs_p_agent_state=z(E, 6, S), obs_p_postsubsidy=z(E, S), obs_p_lagged=z(E, S), obs_p_time=z(E),
            rewards_a=z(E, S), rewards_p=z(E))
        self.t = t
        a = RefCovidArgs()
        for n in _PTR_FIELDS:
            setattr(a, n, t[n].data_ptr())
        ints = dict(action_cooldown_period=p["action_cooldown_period"], num_stringency_levels=NS,
                    subsidy_interval=p["subsidy_interval"], num_subsidy_levels=NL, delivery_interval=p["delivery_interval"],
                    time_when_vaccine_delivery_begins=p["time_when_vaccine_delivery_begins"], beta_delay=bd, filter_len=L,
                    num_filters=Fn, num_days_in_an_year=365, value_of_life=int(p["value_of_life"]), n_agents=S + 1,
                    episode_length=T, n_envs=E)
        for n in _INT_FIELDS:
            setattr(a, n, int(ints[n]))
        floats = dict(gamma=p["gamma"], death_rate=p["death_rate"],
                      infection_too_sick_to_work_rate=p["infection_too_sick_to_work_rate"],
                      population_between_age_18_65=p["pop_between_age_18_65"],
                      daily_production_per_worker=p["daily_production_per_worker"],
                      risk_free_interest_rate=p["risk_free_interest_rate"], economic_reward_crra_eta=p["crra_eta"],
                      min_marginal_planner_health_index=p["min_marginal_planner_health_index"],
                      max_marginal_planner_health_index=p["max_marginal_planner_health_index"],
                      min_marginal_planner_economic_index=p["min_marginal_planner_economic_index"],
                      max_marginal_planner_economic_index=p["max_marginal_planner_economic_index"],
                      weightage_on_marginal_planner_health_index=p["w_planner_health"],
                      weightage_on_marginal_planner_economic_index=p["w_planner_econ"],
                      planner_health_norm=p["planner_health_norm"], planner_economic_norm=p["planner_economic_norm"])
        for n in _FLOAT_FIELDS:
            setattr(a, n, float(F32(floats[n])))
        self.args = a
        # what the reference's wrapper restores at reset (save_copy_and_apply_at_reset=True)
        self._saved_names = ["susceptible", "infected", "recovered", "deaths", "unemployed", "vaccinated", "stringency_level",
                             "subsidy_level", "subsidy", "postsubsidy_productivity", "productivity", "incapacitated",
                             "cant_work", "num_people_that_can_work", "beta", "delta_stringency_level", "signal",
                             "action_in_cooldown_until", "num_vaccines_available_t"]
        self._light_reset = ["delta_stringency_level", "action_in_cooldown_until", "num_vaccines_available_t"]
        self._saved = {n: t[n][:1].clone() for n in self._saved_names}   # every replica starts from the same snapshot

    def launches_per_step(self):
        return 5

    def reset(self, light=False):
        """Restore the saved arrays (all of them, as WarpDrive's reset does; `light`: only those whose stale contents
        would change the next episode - the time-indexed arrays are rewritten before they are read)."""
        for n in (self._light_reset if light else self._saved_names):
            self.t[n].copy_(self._saved[n].expand_as(self.t[n]))
        self.t["timestep"].zero_(); self.t["done"].zero_()

    def step(self):
        rc = self.lib.ref_covid_step(C.byref(self.args), C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError("reference COVID kernels: CUDA error %d" % rc)

    def read_obs(self, e):
        """Host copy of env e's outputs in oracle/covid_oracle.py's layout."""
        g = lambda k: self.t[k][e].detach().cpu().numpy()
        return dict(agent_state=g("obs_a_agent_state"), postsubsidy=g("obs_a_postsubsidy"), lagged=g("obs_a_lagged"),
                    policy_ind=g("obs_a_policy_indicators"),
                    scalars=np.array([g("obs_p_time"), g("obs_p_t_until_next_subsidy"), g("obs_p_current_subsidy_level"),
                                      g("obs_p_t_until_next_vaccines")], F32),
                    mask_a=g("obs_a_action_mask"), mask_p=g("obs_p_action_mask"), rew_a=g("rewards_a"),
                    rew_p=np.float64(g("rewards_p")), done=np.int32(g("done")))

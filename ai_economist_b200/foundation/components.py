"""Component configuration carriers.

Same names, kwargs, defaults and validation as the reference components; the dynamics themselves run on the
GPU (csrc/aie_core.cuh).  Each class declares its action subspaces exactly as the reference's get_n_actions."""
import numpy as np

from .registrar import Registry


class BaseComponent:
    name = ""
    component_type = None
    required_entities = []
    agent_subclasses = []

    def __init__(self, n_agents, episode_length, inventory_scale=1):
        self.n_agents = int(n_agents)
        self.episode_length = int(episode_length)
        self.inv_scale = inventory_scale

    @property
    def shorthand(self):
        return self.name if self.component_type is None else self.component_type

    def get_n_actions(self, agent_cls_name):
        return None

    def spec_fields(self):
        return {}


component_registry = Registry(BaseComponent)


@component_registry.add
class Build(BaseComponent):
    """reference: components/build.py:16-68"""
    name = "Build"
    component_type = "Build"
    required_entities = ["Wood", "Stone", "Coin", "House", "Labor"]
    agent_subclasses = ["BasicMobileAgent"]

    def __init__(self, *a, payment=10, payment_max_skill_multiplier=1, skill_dist="none", build_labor=10.0, **k):
        super().__init__(*a, **k)
        self.payment = int(payment)
        assert self.payment >= 0
        self.payment_max_skill_multiplier = int(payment_max_skill_multiplier)
        assert self.payment_max_skill_multiplier >= 1
        self.build_labor = float(build_labor)
        assert self.build_labor >= 0
        self.skill_dist = skill_dist.lower()
        assert self.skill_dist in ["none", "pareto", "lognormal"]

    def get_n_actions(self, agent_cls_name):
        return 1 if agent_cls_name == "BasicMobileAgent" else None

    def sample_skills(self, rs, n):
        """build.py:224-254 -> (build_payment[n], sampled_skill[n]); consumes the stream like the reference."""
        pay, skill = np.zeros(n), np.zeros(n)
        pmsm = self.payment_max_skill_multiplier
        for i in range(n):
            if self.skill_dist == "none":
                s, rate = 1, 1
            elif self.skill_dist == "pareto":
                s = rs.pareto(4)
                rate = np.minimum(pmsm, (pmsm - 1) * s + 1)
            else:
                s = rs.lognormal(-1, 0.5)
                rate = np.minimum(pmsm, (pmsm - 1) * s + 1)
            pay[i], skill[i] = float(rate * self.payment), float(s)
        return pay, skill

    def spec_fields(self):
        return dict(build_payment=float(self.payment), build_labor=self.build_labor)


@component_registry.add
class Gather(BaseComponent):
    """reference: components/move.py:16-62"""
    name = "Gather"
    required_entities = ["Coin", "House", "Labor"]
    agent_subclasses = ["BasicMobileAgent"]

    def __init__(self, *a, move_labor=1.0, collect_labor=1.0, skill_dist="none", **k):
        super().__init__(*a, **k)
        self.move_labor = float(move_labor)
        assert self.move_labor >= 0
        self.collect_labor = float(collect_labor)
        assert self.collect_labor >= 0
        self.skill_dist = skill_dist.lower()
        assert self.skill_dist in ["none", "pareto", "lognormal"]

    def get_n_actions(self, agent_cls_name):
        return 4 if agent_cls_name == "BasicMobileAgent" else None

    def sample_skills(self, rs, n):
        """move.py:193-210"""
        out = np.zeros(n)
        for i in range(n):
            if self.skill_dist == "pareto":
                out[i] = float(np.minimum(2, rs.pareto(3)) / 2)
            elif self.skill_dist == "lognormal":
                out[i] = float(np.minimum(2, rs.lognormal(-2.022, 0.938)) / 2)
        return out

    def spec_fields(self):
        return dict(move_labor=self.move_labor, collect_labor=self.collect_labor)


@component_registry.add
class ContinuousDoubleAuction(BaseComponent):
    """reference: components/continuous_double_auction.py:17-99"""
    name = "ContinuousDoubleAuction"
    component_type = "Trade"
    required_entities = ["Coin", "Labor"]
    agent_subclasses = ["BasicMobileAgent"]
    commodities = ["Stone", "Wood"]

    def __init__(self, *a, max_bid_ask=10, order_labor=0.25, order_duration=50, max_num_orders=None, **k):
        super().__init__(*a, **k)
        self.max_bid_ask = int(max_bid_ask)
        assert self.max_bid_ask >= 1
        self.order_duration = int(order_duration)
        assert self.order_duration >= 1
        self.max_num_orders = int(max_num_orders or self.order_duration)
        assert self.max_num_orders >= 1
        self.order_labor = max(float(order_labor), 0.0)

    def get_n_actions(self, agent_cls_name):
        if agent_cls_name != "BasicMobileAgent":
            return None
        out = []
        for c in self.commodities:
            out.append(("Buy_{}".format(c), 1 + self.max_bid_ask))
            out.append(("Sell_{}".format(c), 1 + self.max_bid_ask))
        return out

    def spec_fields(self):
        return dict(max_bid_ask=self.max_bid_ask, order_duration=self.order_duration,
                    max_num_orders=self.max_num_orders, order_labor=self.order_labor)


@component_registry.add
class PeriodicBracketTax(BaseComponent):
    """reference: components/redistribution.py:78-360; tax_model="saez" runs as a device/host hybrid (foundation/saez.py)"""
    name = "PeriodicBracketTax"
    component_type = "PeriodicTax"
    required_entities = ["Coin"]
    agent_subclasses = ["BasicMobileAgent", "BasicPlanner"]
    US_FEDERAL_2018 = [0.1, 0.12, 0.22, 0.24, 0.32, 0.35, 0.37]

    def __init__(self, *a, disable_taxes=False, tax_model="model_wrapper", period=100, rate_min=0.0, rate_max=1.0,
                 rate_disc=0.05, n_brackets=5, top_bracket_cutoff=100, usd_scaling=1000.0,
                 bracket_spacing="us-federal", fixed_bracket_rates=None, pareto_weight_type="inverse_income",
                 saez_fixed_elas=None, tax_annealing_schedule=None, **k):
        super().__init__(*a, **k)
        self.disable_taxes = bool(disable_taxes)
        self.tax_model = tax_model
        assert tax_model in ["model_wrapper", "us-federal-single-filer-2018-scaled", "saez", "fixed-bracket-rates"]
        self.pareto_weight_type = pareto_weight_type
        assert pareto_weight_type in ["inverse_income", "uniform"]
        self.saez_fixed_elas = None if saez_fixed_elas is None else float(saez_fixed_elas)
        assert self.saez_fixed_elas is None or self.saez_fixed_elas >= 0
        self.period = int(period)
        assert self.period > 0
        self.rate_min = 0.0 if self.disable_taxes else float(rate_min)
        self.rate_max = 0.0 if self.disable_taxes else float(rate_max)
        assert 0 <= self.rate_min <= self.rate_max <= 1.0
        self.rate_disc = float(rate_disc)
        if tax_model == "model_wrapper":
            r = np.arange(self.rate_min, self.rate_max + self.rate_disc, self.rate_disc)
            self.disc_rates = r[r <= self.rate_max]
            assert len(self.disc_rates) > 1 or self.disable_taxes
        else:
            self.disc_rates = np.zeros(0)
        self.n_disc_rates = len(self.disc_rates)
        self.n_brackets = int(n_brackets)
        assert self.n_brackets >= 2
        self.top_bracket_cutoff = float(top_bracket_cutoff)
        assert self.top_bracket_cutoff >= 10
        self.usd_scale = float(usd_scaling)
        assert self.usd_scale > 0
        self.bracket_spacing = bracket_spacing.lower()
        assert self.bracket_spacing in ["linear", "log", "us-federal"]
        if self.bracket_spacing == "linear":
            self.bracket_cutoffs = np.linspace(0, self.top_bracket_cutoff, self.n_brackets)
        elif self.bracket_spacing == "log":
            b0 = self.top_bracket_cutoff / (2 ** (self.n_brackets - 2))
            self.bracket_cutoffs = np.concatenate(
                [[0], 2 ** np.linspace(np.log2(b0), np.log2(self.top_bracket_cutoff), n_brackets - 1)])
        else:
            self.bracket_cutoffs = np.array([0, 9700, 39475, 84200, 160725, 204100, 510300]) / self.usd_scale
            self.n_brackets = len(self.bracket_cutoffs)
            self.top_bracket_cutoff = float(self.bracket_cutoffs[-1])
        assert self.bracket_cutoffs[0] == 0
        if tax_model == "us-federal-single-filer-2018-scaled":
            assert self.bracket_spacing == "us-federal"
            self.fixed_rates = np.minimum(np.array(self.US_FEDERAL_2018), self.rate_max)
        elif tax_model == "fixed-bracket-rates":
            assert isinstance(fixed_bracket_rates, (tuple, list))
            assert np.min(fixed_bracket_rates) >= 0 and np.max(fixed_bracket_rates) <= 1
            assert len(fixed_bracket_rates) == self.n_brackets
            self.fixed_rates = np.minimum(np.array(fixed_bracket_rates, dtype=np.float64), self.rate_max)
        else:
            self.fixed_rates = np.zeros(self.n_brackets)
        self.tax_annealing_schedule = tax_annealing_schedule
        if tax_annealing_schedule is not None:
            assert isinstance(tax_annealing_schedule, (tuple, list))
            self._annealing_warmup, self._annealing_slope = tax_annealing_schedule[0], tax_annealing_schedule[1]
        else:
            self._annealing_warmup = self._annealing_slope = None

    def get_n_actions(self, agent_cls_name):
        if agent_cls_name == "BasicPlanner" and self.tax_model == "model_wrapper" and not self.disable_taxes:
            return [("TaxIndexBracket_{:03d}".format(int(r)), self.n_disc_rates) for r in self.bracket_cutoffs]
        return 0

    def spec_fields(self):
        return dict(
            tax_model={"model_wrapper": 0, "saez": 2}.get(self.tax_model, 1), disable_taxes=int(self.disable_taxes),
            period=self.period, n_brackets=self.n_brackets, n_disc_rates=self.n_disc_rates,
            bracket_cutoffs=[float(x) for x in self.bracket_cutoffs],
            disc_rates=[float(x) for x in self.disc_rates], fixed_rates=[float(x) for x in self.fixed_rates],
            tax_annealing=int(self.tax_annealing_schedule is not None),
            annealing_warmup=float(self._annealing_warmup or 0.0), annealing_slope=float(self._annealing_slope or 0.0),
            rate_max=float(self.rate_max), rate_min=float(self.rate_min))


@component_registry.add
class WealthRedistribution(BaseComponent):
    """reference: components/redistribution.py:21-75.  No actions, observations or masks: every step it equalises the
    agents' total coin (inventory + escrow) by resetting inventory coin.  Should be the last component."""
    name = "WealthRedistribution"
    required_entities = ["Coin"]
    agent_subclasses = ["BasicMobileAgent"]


@component_registry.add
class SimpleLabor(BaseComponent):
    """reference: components/simple_labor.py:16-134.  100 labor actions (hours worked); income = hours x skill.  The skill
    table is the mean of 1000 sorted, clipped Pareto draws, taken in the reference's constructor from the global stream:
    here every replica draws its own table from its own stream (`draw_skills`, called by the one-step-economy scenario
    right after the env has been seeded, like the reference seeds before it builds its components)."""
    name = "SimpleLabor"
    required_entities = ["Coin"]
    agent_subclasses = ["BasicMobileAgent"]
    num_labor_hours = 100

    def __init__(self, *a, mask_first_step=True, payment_max_skill_multiplier=3, pareto_param=4.0, **k):
        super().__init__(*a, **k)
        assert isinstance(mask_first_step, bool)
        self.mask_first_step = mask_first_step
        self.pareto_param = float(pareto_param)
        assert self.pareto_param > 0
        self.payment_max_skill_multiplier = float(payment_max_skill_multiplier)
        self.skills = None   # [n_envs, n_agents], set by draw_skills

    def get_n_actions(self, agent_cls_name):
        return self.num_labor_hours if agent_cls_name == "BasicMobileAgent" else None

    def draw_skills(self, streams):
        """simple_labor.py:72-81 (sic: the shape parameter is the literal 4, not pareto_param)."""
        pmsm, n = self.payment_max_skill_multiplier, self.n_agents
        out = []
        for rs in streams:
            samples = rs.pareto(4, size=(1000, n))
            out.append(np.sort(np.minimum(pmsm, (pmsm - 1) * samples + 1), axis=1).mean(axis=0))
        self.skills = np.stack(out)
        return self.skills

    def spec_fields(self):
        return dict(labor_mask_first_step=int(self.mask_first_step), labor_skill_scale=self.payment_max_skill_multiplier)

"""L1 micro-scenarios (SURVEY §7 parity ladder): hand-built states, scripted actions, one mechanism at a time.

Each scenario runs the product's device code (1-lane host emulation on CPU; the same scenarios run on the CUDA
build under `-m gpu`) next to the pinned C oracle and additionally asserts the outcome the reference is known to
produce (probed against the reference in SURVEY Appendix A/C: CDA tie-breaks, price rule, order window, shared
order cap, NO-OP harvesting, build restrictions, tax with escrow-protected coin).
"""
import numpy as np
import pytest

from oracle.oracle import OracleBatch
from tests import batch_utils as bu

BASE_SPEC = dict(
    components=["Build", "ContinuousDoubleAuction", "Gather"], n_agents=4, height=7, width=7, episode_length=100,
    multi_action_agents=1, has_water=1, obs_range=2, planner_gets_spatial_info=1, allow_observation_scaling=1,
    regen_weight=[0.0, 0.0], isoelastic_eta=0.23, energy_cost=0.21, energy_warmup_constant=0.0, energy_warmup_auto=0,
    planner_reward_type=0, mixing_weight_gini_vs_coin=0.0, build_payment=10.0, build_labor=10.0, move_labor=1.0,
    collect_labor=1.0, max_bid_ask=10, order_duration=3, max_num_orders=2, order_labor=0.25, tax_model=0,
    disable_taxes=0, period=100, n_brackets=0, n_disc_rates=0, bracket_cutoffs=[], disc_rates=[], fixed_rates=[],
    tax_annealing=0, annealing_warmup=0.0, annealing_slope=0.0, rate_max=1.0)


def blank_state(spec, locs, coin=None, stone=None, wood=None, seed=5):
    H, W, A = spec["height"], spec["width"], spec["n_agents"]
    z = lambda: np.zeros((H, W), np.uint8)
    key = np.random.RandomState(seed).get_state()
    return dict(stone=z(), wood=z(), stone_src=z(), wood_src=z(), water=z(), loc=np.array(locs, np.int16),
                coin=np.array(coin if coin is not None else [10.0] * A, np.float64),
                inv_stone=np.array(stone if stone is not None else [0] * A, np.int32),
                inv_wood=np.array(wood if wood is not None else [0] * A, np.int32),
                build_payment=np.full(A, 10.0), build_skill=np.ones(A), bonus_gather_prob=np.zeros(A),
                mt_key=np.array(key[1], np.uint32), mt_pos=int(key[2]), completions=0)


class Pair:
    """Product stepper + oracle loaded with the same hand-built state."""

    def __init__(self, make_stepper, spec, state):
        self.spec = spec
        self.s = make_stepper(spec, 1)
        self.s.load_state({k: (np.asarray(v)[None] if not np.isscalar(v) else np.array([v])) for k, v in state.items()})
        self.o = OracleBatch(spec, 1)
        self.o.load_env(0, state)
        self.n_sub = self.s.dims.n_act_agent

    def step(self, acts=None, planner=None):
        """acts: {agent: {subspace_index: value}} in multi-action mode (0 Build, 1 Buy_Stone, 2 Sell_Stone,
        3 Buy_Wood, 4 Sell_Wood, 5 Gather)."""
        A = self.spec["n_agents"]
        a = np.zeros((1, A, self.n_sub), np.int32)
        for ag, d in (acts or {}).items():
            for k, v in d.items():
                a[0, ag, k] = v
        p = None if planner is None else np.asarray(planner, np.int32)[None]
        ba, bp = self.s.buf["actions_agent"], self.s.buf["actions_planner"]
        if isinstance(ba, np.ndarray):
            ba[...] = a
            if p is not None:
                bp[...] = p
        else:
            import torch
            ba.copy_(torch.as_tensor(a))
            if p is not None:
                bp.copy_(torch.as_tensor(p))
        self.s.step()
        self.o.step(a, p)
        bu.compare_env(self.o, self.s, 0, "micro")
        return self.s.read_state(0)


def emu(spec, n):
    from tests.emu.emu_stepper import EmuStepper
    return EmuStepper(spec, n)


def cuda(spec, n):
    from ai_economist_b200.stepper import CudaStepper
    return CudaStepper(spec, n, device="cuda:0", auto_reset=False)


BACKENDS = [pytest.param(emu, id="emu"), pytest.param(cuda, id="cuda", marks=pytest.mark.gpu)]
BUILD, BUY_S, SELL_S, BUY_W, SELL_W, MOVE = 0, 1, 2, 3, 4, 5
LOCS = [[0, 0], [0, 6], [6, 0], [6, 6]]


@pytest.mark.parametrize("mk", BACKENDS)
def test_cda_same_step_tie_goes_to_lowest_agent_and_ask_price_rules(mk):
    st = blank_state(BASE_SPEC, LOCS, stone=[0, 0, 0, 1])
    p = Pair(mk, BASE_SPEC, st)
    # agents 0,1,2 bid 5 for Stone, agent 3 asks 3 in the same step: lowest index (0) trades, at the ask price (same age)
    s = p.step({0: {BUY_S: 6}, 1: {BUY_S: 6}, 2: {BUY_S: 6}, 3: {SELL_S: 4}})
    assert s["inv"][0, 0] == 1 and s["inv"][1, 0] == 0 and s["inv"][3, 0] == 0
    assert s["coin"][0] == 10 - 3 and s["coin"][3] == 10 + 3
    assert s["esc_coin"][1] == 5 and s["esc_coin"][2] == 5 and s["n_orders"][0].tolist() == [0, 1, 1, 0]


@pytest.mark.parametrize("mk", BACKENDS)
def test_cda_older_bid_sets_the_price_and_has_priority(mk):
    st = blank_state(BASE_SPEC, LOCS, stone=[0, 0, 0, 2])
    p = Pair(mk, BASE_SPEC, st)
    p.step({1: {BUY_S: 8}})                       # agent 1 bids 7 at t=1
    s = p.step({0: {BUY_S: 8}, 3: {SELL_S: 3}})   # t=2: agent 0 bids 7 too; ask 2 -> older bid (agent 1) wins, pays its bid 7
    assert s["inv"][1, 0] == 1 and s["inv"][0, 0] == 0
    assert s["coin"][3] == 10 + 7 and s["coin"][1] == 10 - 7 and s["esc_coin"][0] == 7


@pytest.mark.parametrize("mk", BACKENDS)
def test_cda_no_self_trade_order_cap_and_unaffordable_bid(mk):
    st = blank_state(BASE_SPEC, LOCS, coin=[4.0, 10, 10, 10], stone=[2, 0, 0, 0])
    p = Pair(mk, BASE_SPEC, st)
    s = p.step({0: {BUY_S: 6, SELL_S: 2}})        # bid 5 with 4 coin: rejected; ask 1 accepted
    assert s["n_orders"][0, 0] == 1 and s["esc_coin"][0] == 0 and s["esc"][0, 0] == 1
    s = p.step({0: {BUY_S: 3}})                    # bid 2 >= own ask 1, but an agent never trades with itself
    assert s["inv"][0, 0] == 1 and s["esc"][0, 0] == 1 and s["esc_coin"][0] == 2 and s["n_orders"][0, 0] == 2
    s = p.step({0: {SELL_S: 1}})                   # cap (max_num_orders=2) is shared by bids and asks: rejected
    assert s["n_orders"][0, 0] == 2 and s["inv"][0, 0] == 1


@pytest.mark.parametrize("mk", BACKENDS)
def test_cda_order_lives_duration_plus_one_steps_then_refunds(mk):
    st = blank_state(BASE_SPEC, LOCS)
    p = Pair(mk, BASE_SPEC, st)
    s = p.step({2: {BUY_W: 5}})                    # t=1, order_duration=3: matchable at t=1..4
    assert s["esc_coin"][2] == 4
    for _ in range(2):
        s = p.step()
        assert s["esc_coin"][2] == 4
    s = p.step()                                    # end of t=4: lifetime 4 > 3 -> expired, coin refunded
    assert s["esc_coin"][2] == 0 and s["coin"][2] == 10 and s["n_orders"][1, 2] == 0


@pytest.mark.parametrize("mk", BACKENDS)
def test_gather_noop_harvests_bonus_and_blocking(mk):
    st = blank_state(BASE_SPEC, [[3, 3], [3, 4], [0, 0], [6, 6]])
    st["stone"][3, 3] = 1; st["stone_src"][3, 3] = 1     # agent 0 stands on stone
    st["wood"][3, 5] = 1; st["wood_src"][3, 5] = 1
    st["water"][2, 3] = 1
    st["bonus_gather_prob"] = np.array([1.0, 0.0, 0.0, 0.0])
    p = Pair(mk, BASE_SPEC, st)
    s = p.step({0: {MOVE: 2}, 1: {MOVE: 2}})      # 0 -> right into 1's cell (blocked unless 1 left first); 1 -> right onto wood
    assert s["inv"][1, 1] == 1 and s["loc"][1].tolist() == [3, 5]
    # agent 0 harvested 2 stone (bonus prob 1) wherever it ended up this step only if it stayed on [3,3]
    if s["loc"][0].tolist() == [3, 3]:
        assert s["inv"][0, 0] == 2
    s = p.step({0: {MOVE: 3}})                     # up into water from [3,3] (or not adjacent any more): never enters water
    assert s["loc"][0].tolist() != [2, 3]
    s = p.step({2: {MOVE: 1}, 3: {MOVE: 4}})      # border moves are no-ops
    assert s["loc"][2].tolist() == [0, 0] and s["loc"][3].tolist() == [6, 6]


@pytest.mark.parametrize("mk", BACKENDS)
def test_build_rules_and_house_access(mk):
    st = blank_state(BASE_SPEC, [[3, 3], [3, 4], [0, 0], [6, 6]], stone=[1, 2, 0, 1], wood=[1, 2, 1, 1])
    st["stone_src"][3, 4] = 1                      # agent 1 stands on an (empty) source block: cannot build there
    p = Pair(mk, BASE_SPEC, st)
    s = p.step({0: {BUILD: 1}, 1: {BUILD: 1}, 2: {BUILD: 1}})
    assert s["owner"][3, 3] == 0 and s["coin"][0] == 20 and s["labor"][0] == 10          # built
    assert s["owner"][3, 4] == -1 and s["inv"][1].tolist() == [2, 2]                      # source block: refused
    assert s["inv"][2].tolist() == [0, 1] and s["coin"][2] == 10                          # no stone: refused
    s = p.step({0: {BUILD: 1}, 1: {MOVE: 1}})     # cannot build twice on the same cell; 1 cannot enter 0's house
    assert s["coin"][0] == 20 and s["loc"][1].tolist() == [3, 4]
    s = p.step({0: {MOVE: 3}})                     # owner steps off ...
    s = p.step({1: {MOVE: 1}, 0: {MOVE: 4}})      # ... and only the owner may step back on
    assert s["loc"][1].tolist() == [3, 4]


TAX_SPEC = dict(BASE_SPEC, components=["Build", "ContinuousDoubleAuction", "Gather", "PeriodicBracketTax"], period=3,
                n_brackets=7, n_disc_rates=21, bracket_cutoffs=[0, 9.7, 39.475, 84.2, 160.725, 204.1, 510.3],
                disc_rates=[float(x) for x in np.arange(0, 1.05, 0.05)[np.arange(0, 1.05, 0.05) <= 1.0]],
                fixed_rates=[0.0] * 7)


@pytest.mark.parametrize("mk", BACKENDS)
def test_tax_day_collects_from_inventory_only_and_redistributes(mk):
    st = blank_state(TAX_SPEC, LOCS, coin=[10.0, 10, 10, 10], stone=[0, 1, 0, 0], wood=[0, 1, 0, 0])
    p = Pair(mk, TAX_SPEC, st)
    rates = [21] * 7                                # planner sets every bracket to 100 % on the first day
    p.step({1: {BUILD: 1}, 0: {BUY_S: 11}}, planner=rates)   # agent 1 earns 10; agent 0 escrows all 10 coin in a bid
    p.step()
    s = p.step()                                    # tax day (period 3)
    # agent 1: income 10 -> due 9.7*1.0 + 0.3*1.0 = 10, paid from inventory; agent 0: income 0 -> pays nothing,
    # its escrowed coin is untouched; everyone receives the lump sum 10 / 4
    assert np.isclose(s["last_income"][1], 10) and np.isclose(s["coin"][1], 20 - 10 + 2.5)
    assert s["esc_coin"][0] == 10 and np.isclose(s["coin"][0], 2.5)
    assert int(s["tax_pos"][0]) == 1

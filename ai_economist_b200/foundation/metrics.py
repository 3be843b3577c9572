"""`env.metrics` for the batched gather-trade-build env, computed from one env's state record.

Mirrors BaseEnvironment.metrics (ai_economist/foundation/base/base_env.py:421-432): the scenario's
scenario_metrics() (scenarios/simple_wood_and_stone/layout_from_file.py:595-650) merged with every component's
get_metrics() under its shorthand (build.py:198-222 "Build", continuous_double_auction.py:585-641 "Trade",
redistribution.py:1141-1186 "PeriodicTax"; Gather defines none).

The event logs the reference walks through (Build.builds, ContinuousDoubleAuction.executed_trades,
PeriodicBracketTax._schedules/_occupancy/all_effective_tax_rates/taxes) are kept on the device as running sums in
the "stats" section of the state record (include/aie_b200.h, aie_dims.n_stats), so the metrics of any env - or of the
episode an auto-reset env just finished (episode_final snapshots) - are a pure function of its record.
"""
import numpy as np

_COMMODITIES = ("Stone", "Wood")


def gini(x):
    """scenarios/utils/social_metrics.py:10-46 (both branches)."""
    x = np.asarray(x, np.float64)
    n = len(x)
    if n < 30:
        diff = np.sum(np.abs(x.reshape((n, 1)) - x.reshape((1, n))))
        norm = 2 * n * x.sum(axis=0)
        return (diff / (norm + 1e-10)) / ((n - 1) / n)
    s = np.sort(x)
    return 1 - (2 / (n + 1)) * np.sum(np.cumsum(s) / (np.sum(s) + 1e-10))


def energy_weight(spec, completions, warmup_integrator):
    """layout_from_file.py:249-267: the labor-cost annealing weight ("decay" counts completions, "auto" the integrator)."""
    warm = float(spec.get("energy_warmup_constant", 0) or 0)
    if warm <= 0.0:
        return 1.0
    x = warmup_integrator if int(spec.get("energy_warmup_auto", 0)) else completions
    return float(1.0 - np.exp(-x / warm))


def metrics_from_state(spec, st, saez_elasticity=None):
    """spec: the env spec (scenario_spec_fields + components); st: dict with coin, esc_coin, inv, esc, labor, util_prev,
    auto_warmup, completions, cell, stats (BatchStepper.read_state layout).  Returns the reference's metrics dict."""
    A = int(spec["n_agents"])
    coin = np.asarray(st["coin"], np.float64) + np.asarray(st["esc_coin"], np.float64)  # total_endowment("Coin")
    util = np.asarray(st["util_prev"], np.float64)  # curr_optimization_metric: agents then planner
    m = {}
    if int(spec.get("scenario_kind", 0)) == 1:
        # one-step-economy (scenarios/one_step_economy/one_step_economy.py:210-262): averages instead of per-agent entries;
        # the pre-tax incomes (production) travel in the record's build_payment field
        prod = np.asarray(st["build_payment"], np.float64)
        m["social/productivity"] = float(np.sum(coin))
        m["social/equality"] = float(1 - gini(coin))
        m["social_welfare/coin_eq_times_productivity"] = float((1 - gini(coin)) * (np.sum(coin) / A))
        w = 1 / np.maximum(prod, 1)
        w = w / np.sum(w)
        m["social_welfare/inv_income_weighted_utility"] = float(np.sum(util[:A] * w))
        m["endow/avg_agent/Coin"] = float(np.mean(coin))
        m["endogenous/avg_agent/Labor"] = float(np.mean(np.asarray(st["labor"], np.float64)))
        m["util/avg_agent"] = float(np.mean(util[:A]))
        m["endow/p/Coin"] = 0
        m["util/p"] = float(util[A])
        _component_metrics(m, spec, st, A, coin, saez_elasticity)
        return m
    m["social/productivity"] = float(np.sum(coin))
    m["social/equality"] = float(1 - gini(coin))
    m["social_welfare/coin_eq_times_productivity"] = float((1.0 * (1 - gini(coin)) + 0.0) * (np.sum(coin) / A))
    w = 1 / np.maximum(coin, 1)
    w = w / np.sum(w)
    m["social_welfare/inv_income_weighted_coin_endow"] = float(np.sum(coin * w))
    m["social_welfare/inv_income_weighted_utility"] = float(np.sum(util[:A] * w))
    inv = np.asarray(st["inv"]).reshape(A, 2) + np.asarray(st["esc"]).reshape(A, 2)
    for a in range(A):
        m["endow/%d/Coin" % a] = float(coin[a])
        m["endow/%d/Stone" % a] = float(inv[a, 0])
        m["endow/%d/Wood" % a] = float(inv[a, 1])
        m["endogenous/%d/Labor" % a] = float(st["labor"][a])
        m["util/%d" % a] = float(util[a])
    for r in ("Coin", "Stone", "Wood"):  # the planner holds nothing in these scenarios
        m["endow/p/%s" % r] = 0.0
    m["util/p"] = float(util[A])
    completions = int(np.asarray(st["completions"]).reshape(-1)[0])
    warm_int = int(np.asarray(st["auto_warmup"]).reshape(-1)[0])
    ew = energy_weight(spec, completions, warm_int)
    m["labor/weighted_cost"] = float(spec["energy_cost"]) * ew
    m["labor/warmup_integrator"] = warm_int

    _component_metrics(m, spec, st, A, coin, saez_elasticity)
    return m


def _component_metrics(m, spec, st, A, coin, saez_elasticity):
    """Every component's get_metrics() under its shorthand, from the running sums in the record's stats section."""
    stats = np.asarray(st["stats"], np.float64)
    comps = list(spec["components"])
    st_trade = 1 + A
    for name in comps:
        if name == "Build":
            for a in range(A):
                m["Build/%d/n_builds" % a] = int(stats[1 + a])
            m["Build/total_builds"] = int(np.sum((np.asarray(st["cell"]) & 32) != 0))  # houses standing on the map
        elif name == "ContinuousDoubleAuction":
            for a in range(A):
                for ci, cn in enumerate(_COMMODITIES):
                    for side, prefix in ((0, "Sell"), (1, "Buy")):
                        n, psum = stats[st_trade + ((a * 2 + ci) * 2 + side) * 2: st_trade + ((a * 2 + ci) * 2 + side) * 2 + 2]
                        v = float("nan") if n == 0 else float(psum / n)
                        for k in ("price", "cost", "income"):  # cost == income == price (:305-309)
                            m["Trade/%d/%s%s/%s" % (a, prefix, cn, k)] = v
                        m["Trade/%d/%s%s/n_sales" % (a, prefix, cn)] = int(n)
            m["Trade/n_trades"] = int(stats[0])
        elif name == "PeriodicBracketTax":
            t0 = st_trade + 8 * A
            periods, collected, eff_sum = stats[t0], stats[t0 + 1], stats[t0 + 2]
            sched, occ = stats[t0 + 3: t0 + 19], stats[t0 + 19: t0 + 35]
            inc, paid = stats[t0 + 35: t0 + 35 + A], stats[t0 + 35 + A: t0 + 35 + 2 * A]
            cutoffs = list(spec["bracket_cutoffs"])[: int(spec["n_brackets"])]
            n_obs = max(1.0, float(np.sum(occ)))
            groups = {}   # the reference keys its logs by "%03d" % int(cutoff): brackets that share a key are pooled
            for b, c in enumerate(cutoffs):
                groups.setdefault("%03d" % int(c), []).append(b)
            for k, bs in groups.items():
                m["PeriodicTax/avg_bracket_rate/%s" % k] = (float(sum(sched[b] for b in bs) / (periods * len(bs)))
                                                            if periods else float("nan"))
                m["PeriodicTax/bracket_occupancy/%s" % k] = float(sum(occ[b] for b in bs) / n_obs)
            if not spec.get("disable_taxes", False):
                m["PeriodicTax/avg_effective_tax_rate"] = float(eff_sum / (periods * A)) if periods else float("nan")
                m["PeriodicTax/total_collected_taxes"] = float(collected)
                for i, tag in ((int(np.argmin(coin)), "poorest"), (int(np.argmax(coin)), "richest")):
                    m["PeriodicTax/avg_tax_rate/%s" % tag] = float(paid[i] / max(0.001, inc[i]))
                if int(spec["tax_model"]) == 2:  # the running elasticity estimate lives in the host-side estimator
                    m["PeriodicTax/saez/estimated_elasticity"] = float(saez_elasticity)

"""foundation/saez.py: the replica-batched Saez estimator against the per-replica one (which restates
redistribution.py:437-823 call by call and is pinned on a reference trace in tests/test_device_reset.py)."""
import numpy as np
import pytest

from ai_economist_b200.foundation.saez import BUFFER_SIZE, SaezBatch, SaezEstimator

CUTOFFS = [0, 9.7, 39.475, 84.2, 160.725, 204.1, 510.3]


@pytest.mark.parametrize("weights,fixed", [("inverse_income", None), ("uniform", None), ("inverse_income", 0.4)])
def test_batched_estimator_tracks_the_per_replica_estimator(weights, fixed):
    E, A, periods = 12, 10, 70
    rng = np.random.RandomState(3)
    one = [SaezEstimator(CUTOFFS, 0.0, 1.0, weights, fixed) for _ in range(E)]
    bat = SaezBatch(E, CUTOFFS, 0.0, 1.0, weights, fixed)
    scale = rng.choice([5.0, 60.0, 400.0], size=E)          # poor / middling / rich replicas: empty bins, open top bin
    for t in range(periods):
        inc = rng.gamma(1.2, scale[:, None], size=(E, A)) - rng.rand(E, A) * (t % 7 == 0)   # some negative incomes
        if t % 11 == 0:
            inc[::3] = 0.0                                   # nobody earns: degenerate regressions
        tau = np.clip(rng.rand(E, 1) * 0.6 + 0.1 * rng.randn(E, A), 0, 1.0) * (t % 13 != 0)  # sometimes constant rates
        rows = np.arange(E) if t % 5 else np.arange(0, E, 2)  # not every replica has a tax day every time
        for e in rows:
            one[e].add_samples(inc[e], tau[e])
        bat.add_samples(rows, inc[rows], tau[rows])
        assert np.array_equal(bat.count, [min(len(o.buffer), BUFFER_SIZE) for o in one])
        ready = np.array([o.ready for o in one])
        assert np.array_equal(bat.ready(np.arange(E)), ready)
        sel = np.nonzero(ready)[0]
        if len(sel) == 0:
            continue
        caps = None if t % 3 else rng.choice([0.3, 0.45, 1.0], size=len(sel))   # per-replica curr_rate_max (tax annealing)
        want = np.stack([one[e].new_period_rates(None if caps is None else caps[i]) for i, e in enumerate(sel)])
        got = bat.new_period_rates(sel, caps)
        if caps is not None:
            assert np.all(got <= caps[:, None] + 1e-15)
        assert np.allclose(want, got, rtol=1e-10, atol=1e-12), (t, np.abs(want - got).max())
        assert np.allclose([one[e].elas_t for e in sel], bat.elas_t[sel], rtol=1e-10, atol=1e-13)
        assert np.allclose(np.stack([one[e].running_avg for e in sel]), bat.running_avg[sel], rtol=1e-10, atol=1e-13)
    assert sum(o.ready for o in one) == E


def test_saez_host_uses_the_batched_estimator_above_64_replicas_and_agrees_with_the_exact_loop():
    """Plumbing through BatchedFoundationEnv: 66 replicas (batched estimator) vs 2 replicas (per-replica loop) with the
    same seeds and actions; the first bracket rates the Saez formula produces must agree."""
    from ai_economist_b200 import foundation
    from ai_economist_b200.foundation.saez import SaezBatch, SaezLoop
    from oracle.configs import CONFIGS
    from tests import batch_utils as bu
    from tests.emu.emu_stepper import emu_factory
    kw = dict(CONFIGS["saez_reset"])
    name = kw.pop("scenario_name")
    kw["components"] = [(n, dict(k, period=5) if n == "PeriodicBracketTax" else k) for n, k in kw["components"]]
    big = foundation.make_env_instance(name, n_envs=66, stepper_factory=emu_factory, seeds=list(range(100, 166)), **kw)
    small = foundation.make_env_instance(name, n_envs=2, stepper_factory=emu_factory, seeds=[100, 101], **kw)
    assert isinstance(big._saez.batch, SaezBatch) and isinstance(small._saez.batch, SaezLoop)
    big.reset(); small.reset()
    seg = bu.segments(small.spec, "a")
    rng = np.random.RandomState(1)
    first = None
    for t in range(1, 300):
        aa = bu.sample_from_masks(small.stepper.to_numpy(small.stepper.buf["mask_agent"]), seg, rng)
        full = np.zeros((66,) + aa.shape[1:], np.int32)
        full[:2] = aa
        small.step((aa, None)); big.step((full, None))
        rs = small.stepper.to_numpy(small.stepper.state_view("saez_rates"))
        rb = big.stepper.to_numpy(big.stepper.state_view("saez_rates"))
        if bool(small._saez.est[0].ready) and first is None and np.any(small._saez.batch.running_avg[0] != 0):
            first = t
            assert np.allclose(rs[:2], rb[:2], rtol=1e-9, atol=1e-12), t
            break
    assert first is not None and 240 <= first <= 260        # 500 samples = 50 tax days of 10 agents, period 5
    assert np.all(np.isfinite(rb)) and rb.min() >= 0.0 and rb.max() <= 1.0

"""Shared by the CPU (emulation) and GPU metrics tests: replay a golden action trace through a BatchStepper and compare
`metrics_from_state` with the `env.metrics` the unmodified reference produced (tests/golden_metrics/*.json)."""
import glob
import json
import math
import os

import numpy as np

from ai_economist_b200.foundation.metrics import metrics_from_state
from tests import golden_utils as gu
from tests.stepper_adapters import GoldenStepperAdapter

HERE = os.path.dirname(os.path.abspath(__file__))


def metric_files():
    return sorted(glob.glob(os.path.join(HERE, "golden_metrics", "*.json")))


def compare_metrics(ref, got, label, rtol=1e-6):
    assert set(ref) == set(got), "%s: key sets differ: %s" % (label, sorted(set(ref) ^ set(got))[:8])
    for k, rv in ref.items():
        gv = got[k]
        if rv is None:
            assert isinstance(gv, float) and math.isnan(gv), "%s: %s expected NaN, got %r" % (label, k, gv)
        elif isinstance(rv, int):
            assert int(gv) == rv and float(gv) == float(rv), "%s: %s = %r, reference %r" % (label, k, gv, rv)
        else:
            assert abs(gv - rv) <= rtol * max(1.0, abs(rv)), "%s: %s = %r, reference %r" % (label, k, gv, rv)


def replay_and_check(json_path, make_batch_stepper, max_steps=None):
    ref = json.load(open(json_path))
    z, meta, init = gu.load_fixture(os.path.join(HERE, "golden", ref["fixture"]))
    spec = meta["spec"]
    ad = GoldenStepperAdapter(make_batch_stepper(spec), init)
    want = dict(zip(ref["steps"], ref["metrics"]))
    n = int(meta["n_steps"]) if max_steps is None else min(int(meta["n_steps"]), max_steps)
    checked = 0
    for t in range(1, n + 1):
        ap = z["act_p"][t - 1].astype(np.int32)
        ad.step(z["act_a"][t - 1].astype(np.int32), ap if ap.size else None)
        if t in want:
            compare_metrics(want[t], metrics_from_state(spec, ad.state()), "%s t=%d" % (ref["fixture"], t))
            checked += 1
    return checked

"""The reference's OWN COVID-19 CUDA kernels on the same GPU (BASELINE config 4 / VERDICT r1 item 5).

oracle/_ref/libref_covid_cuda.so is built in the build container from the sources under /root/reference
(oracle/build_ref_covid.py; nothing copied) and travels to the GPU box as a built artefact.  Here:

  * the reference CUDA path replays the golden trace recorded from the reference's PYTHON path (the comparison the
    reference's tests/run_covid19_cpu_gpu_consistency_checks.py makes through WarpDrive's EnvironmentCPUvsGPU, with
    num_envs = 3), and
  * aie_covid_step_kernel replays the same trace next to it: both against the golden, and against each other.

Tolerances.  The reference's CUDA path is a float32 re-implementation of a Python path that promotes to float64 in
places, so the two differ in the last digits (that is why WarpDrive's checker compares with a tolerance rather than
exactly); REF_RTOL / REF_ATOL below are what the reference's own CUDA path needs against its own Python path on this
trace (measured maxima are printed).  aie_covid_step_kernel follows the Python path and keeps the suite's 1e-6.
"""
import json
import os

import numpy as np
import pytest

from ai_economist_b200.foundation.covid19 import build_covid_params
from oracle import build_ref_covid

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_covid")
OBS_KEYS = ["agent_state", "postsubsidy", "lagged", "policy_ind", "scalars"]
REF_RTOL, REF_ATOL = 2e-3, 2e-4   # reference CUDA (float32) vs reference Python (float64 promotions), see module docstring
RTOL, ATOL = 1e-6, 1e-9           # aie_covid_step_kernel vs reference Python

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not build_ref_covid.available(),
                                                  reason="oracle/_ref/libref_covid_cuda.so not built (needs /root/reference)")]


def _load(name="covid_seed3.npz"):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    meta = json.loads(str(z["meta_json"]))
    return z, meta, build_covid_params(**meta["kwargs"])


def _maxdev(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-3))) if a.size else 0.0


def test_reference_cuda_kernels_replay_the_python_golden_trace_next_to_ours():
    import torch
    from ai_economist_b200.covid_stepper import CudaCovidStepper
    from oracle.ref_covid_cuda import RefCovidCuda

    z, meta, p = _load()
    E = 3
    ref = RefCovidCuda(p, E)
    ours = CudaCovidStepper(p, E, auto_reset=False)
    ours.reset()
    worst = {}
    for t in range(1, meta["n_steps"] + 1):
        a = torch.as_tensor(z["act_a"][t - 1].astype(np.int32), device="cuda")
        pl = int(z["act_p"][t - 1])
        ref.t["actions_a"][:] = a; ref.t["actions_p"][:] = pl
        ours.buf["actions_agent"][:] = a; ours.buf["actions_planner"][:] = pl
        ref.step(); ours.step()
        if t % 7 and t < meta["n_steps"] - 2 and t > 3:
            continue   # full comparison on a subset of days (every day costs two D2H round trips per field)
        for e in (0, E - 1):
            r, o = ref.read_obs(e), ours.read_obs(e)
            for k in OBS_KEYS + ["rew_a"]:
                g = z[k][t] if k != "rew_a" else z["rew_a"][t - 1]
                worst[k] = max(worst.get(k, 0.0), _maxdev(r[k], g))
                assert np.allclose(r[k], g, rtol=REF_RTOL, atol=REF_ATOL), "reference CUDA vs Python golden, day %d: %s" % (t, k)
                assert np.allclose(o[k], g, rtol=RTOL, atol=ATOL), "ours vs Python golden, day %d: %s" % (t, k)
                assert np.allclose(o[k], r[k], rtol=REF_RTOL, atol=REF_ATOL), "ours vs reference CUDA, day %d: %s" % (t, k)
            worst["rew_p"] = max(worst.get("rew_p", 0.0), _maxdev(r["rew_p"], z["rew_p"][t - 1]))
            assert np.isclose(float(r["rew_p"]), float(z["rew_p"][t - 1]), rtol=REF_RTOL, atol=REF_ATOL), "day %d: rew_p" % t
            assert np.isclose(float(o["rew_p"]), float(z["rew_p"][t - 1]), rtol=RTOL, atol=ATOL)
            # masks are exact in all three
            assert np.array_equal(r["mask_a"], z["mask_a"][t]) and np.array_equal(r["mask_p"], z["mask_p"][t]), "day %d: masks" % t
            assert np.array_equal(o["mask_a"], z["mask_a"][t]) and np.array_equal(o["mask_p"], z["mask_p"][t])
            assert int(r["done"]) == int(z["done"][t - 1]) == int(o["done"])
    print("max relative deviation of the reference CUDA path from its Python path:", {k: "%.2e" % v for k, v in worst.items()})


def test_reference_cuda_reset_restores_the_saved_arrays():
    import torch
    from oracle.ref_covid_cuda import RefCovidCuda

    z, meta, p = _load()
    ref = RefCovidCuda(p, 2)
    first = None
    for ep in range(2):
        for t in range(1, 31):
            ref.t["actions_a"][:] = torch.as_tensor(z["act_a"][t - 1].astype(np.int32), device="cuda")
            ref.t["actions_p"][:] = int(z["act_p"][t - 1])
            ref.step()
        snap = {k: v.copy() for k, v in ref.read_obs(1).items()}
        if first is None:
            first = snap
        else:
            for k in first:
                assert np.array_equal(first[k], snap[k]), k
        ref.reset()

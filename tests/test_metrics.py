"""`env.metrics`: the device-side episode statistics + metrics_from_state against the
`env.metrics` of the unmodified reference on the golden action traces (CPU: 1-lane emulation of the device source)."""
import numpy as np
import pytest

from tests import metrics_utils as mu
from tests.emu.emu_stepper import EmuStepper


@pytest.mark.parametrize("path", mu.metric_files(), ids=lambda p: p.split("/")[-1])
def test_metrics_match_reference_env_metrics(path):
    assert mu.replay_and_check(path, lambda spec: EmuStepper(spec, 1)) >= 3


def test_oracle_statistics_match_emulated_device_statistics():
    """The C oracle keeps the same running sums from its own restatement of the reference's logs."""
    from oracle.oracle import OracleBatch
    from tests import golden_utils as gu
    from tests.stepper_adapters import GoldenStepperAdapter
    import os
    path = os.path.join(mu.HERE, "golden", "c3_short_period_seed1001.npz")
    z, meta, init = gu.load_fixture(path)
    spec = meta["spec"]
    ad = GoldenStepperAdapter(EmuStepper(spec, 1), init)
    orc = OracleBatch(spec, 1)
    orc.load_env(0, init)
    for t in range(1, int(meta["n_steps"]) + 1):
        aa, ap = z["act_a"][t - 1].astype(np.int32), z["act_p"][t - 1].astype(np.int32)
        ad.step(aa, ap if ap.size else None)
        orc.step(aa[None], ap[None] if ap.size else None)
    so, sp = orc.state(0), ad.state()
    assert so["stats"][0] > 0 and np.allclose(so["stats"], sp["stats"], rtol=1e-9, atol=1e-12)
    assert np.allclose(so["util_prev"], sp["util_prev"], rtol=1e-9) and so["auto_warmup"][0] == sp["auto_warmup"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("path", mu.metric_files(), ids=lambda p: p.split("/")[-1])
def test_cuda_metrics_match_reference_env_metrics(path):
    """The CUDA kernels' episode statistics -> env.metrics of the unmodified reference, through the C-ABI."""
    from ai_economist_b200.stepper import CudaStepper
    assert mu.replay_and_check(path, lambda spec: CudaStepper(spec, 1, auto_reset=False)) >= 3

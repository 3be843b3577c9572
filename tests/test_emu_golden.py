"""CPU: the device source (csrc/aie_core.cuh + the C-ABI host code), compiled for the host with a 1-lane warp
(tests/emu), against the golden traces recorded from the unmodified reference.  A logic check of the kernels
on the GPU-less build container; the `-m gpu` tests repeat it on the real CUDA build."""
import pytest

from tests import golden_utils as gu
from tests.emu.emu_stepper import EmuStepper
from tests.stepper_adapters import GoldenStepperAdapter


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: p.split("/")[-1])
def test_emulated_device_code_matches_reference_golden_trace(path):
    def make(spec, init):
        return GoldenStepperAdapter(EmuStepper(spec, 1), init)

    assert gu.replay(path, make) >= 50


@pytest.mark.parametrize("nt", [32, 128])
@pytest.mark.parametrize("path", [p for p in gu.golden_files() if any(k in p for k in
                                  ("c1_tutorial_seed1", "c3_paper_tax", "c3_short_period", "c5_small", "full_obs", "tax_us_federal"))],
                         ids=lambda p: p.split("/")[-1])
def test_observation_pass_thread_layouts_match_golden_trace(path, nt, monkeypatch):
    """The observation pass is written for NT cooperating threads (32: one warp per env; 128: one CTA per env for large
    records).  The emulation walks every phase for thread index 0..NT-1 in turn (phases only read what earlier phases
    wrote), which checks each thread's slice of the staging / transpose / concat / stream loops - unaligned heads and
    tails, plane boundaries, chunk boundaries - against the reference's observations."""
    monkeypatch.setenv("AIE_EMU_NT", str(nt))

    def make(spec, init):
        return GoldenStepperAdapter(EmuStepper(spec, 1), init)

    assert gu.replay(path, make) >= 50


@pytest.mark.parametrize("cfg,E,steps", [
    ("tax_single_planner", 4, 60), ("uniform_halfwidth", 4, 60), ("quadrant", 3, 40), ("multi_zone", 3, 40),
    ("split_layout", 2, 40),
])
def test_emulated_batch_matches_oracle(cfg, E, steps):
    """Distinct seeds per replica through the public API (host reset -> upload -> step) against the C oracle: the
    CPU twin of tests/test_gpu_parity.py::test_cuda_batch_matches_oracle for the options added after round-1 (e)."""
    import numpy as np

    from ai_economist_b200 import foundation
    from oracle.oracle import OracleBatch
    from tests import batch_utils as bu
    from tests.emu.emu_stepper import emu_factory

    name, kw = bu.product_kwargs(cfg)
    kw.pop("seed", None)
    env = foundation.make_env_instance(name, n_envs=E, stepper_factory=emu_factory, auto_reset=False, seed=4000, **kw)
    host = env.host_reset_arrays()
    env.load_host_state(host)
    orc = OracleBatch(env.spec, E)
    for e in range(E):
        orc.load_env(e, {k: v[e] for k, v in host.items()})
    bu.run_pair(env, orc, steps, np.random.RandomState(9), check_every=20)

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

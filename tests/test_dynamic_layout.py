"""Device-side layout generation of the dynamic scenarios (uniform / quadrant): pieces pinned on the CPU.

The growth step of Uniform.reset_starting_layout (dynamic_layout.py:354-366) thresholds a scipy.signal.convolve2d
output at zero, so the device has to accumulate each output's float64 sum in scipy's own order to land on the same
map.  `conv7_same` in csrc/aie_core.cuh restates that order; here the same order, written in Python, is compared
bit for bit with the scipy of this environment (the host reset path calls scipy itself).  If a future scipy changes
its loop, this test - not a silent one-cell difference in a layout - says so.
The full path (generator + placement + skills across episodes) is pinned by tests/golden_reset/uniform_reset_* and
quadrant_reset_* (tests/test_device_reset.py) and by tests/test_fuzz_subsets.py against the live reference.
"""
import numpy as np
import pytest


def conv7_same_reference_order(x, kern):
    """out[m, n] = sum_{j, k} x[m - j + 3, n - k + 3] * kern[j, k]: rows ascending; first four taps as one expression,
    then taps 4, 5, 6 one by one (csrc/aie_core.cuh: conv7_same)."""
    H, W = x.shape
    out = np.zeros((H, W))
    for m in range(H):
        for n in range(W):
            acc = 0.0
            for j in range(7):
                r = m - j + 3
                t = [(x[r, n - k + 3] if (0 <= r < H and 0 <= n - k + 3 < W and kern[j, k]) else 0.0) for k in range(7)]
                acc += ((t[0] + t[1]) + t[2]) + t[3]
                acc += t[4]
                acc += t[5]
                acc += t[6]
            out[m, n] = acc
    return out


@pytest.mark.parametrize("shape", [(9, 9), (18, 22), (25, 25), (7, 31)])
def test_convolution_accumulation_order_matches_scipy(shape):
    signal = pytest.importorskip("scipy.signal")
    rs = np.random.RandomState(shape[0] * 100 + shape[1])
    for _ in range(3):
        x = (rs.rand(*shape) < 0.1).astype(np.float64) + 0.2 * rs.randn(*shape) - 0.25
        kern = rs.randn(7, 7) > 0
        want = signal.convolve2d(x, kern.astype(np.float32), "same")
        got = conv7_same_reference_order(x, kern)
        assert np.array_equal(want, got), "scipy.signal.convolve2d accumulates in a different order than csrc/aie_core.cuh: conv7_same"


def test_dynamic_scenarios_select_the_device_generator():
    from ai_economist_b200 import foundation
    from tests.emu.emu_stepper import emu_factory
    base = dict(components=[("Build", {}), ("Gather", {})], n_agents=3, world_size=[12, 12], episode_length=5,
                flatten_observations=True, flatten_masks=True, stepper_factory=emu_factory, n_envs=1, seed=3)
    for name, kind in [("uniform/simple_wood_and_stone", 1), ("quadrant/simple_wood_and_stone", 2)]:
        env = foundation.make_env_instance(name, **base)
        assert env.spec["reset_mode"] == 1 and env.spec["dyn_layout"] == kind
    mz = foundation.make_env_instance("multi_zone/simple_wood_and_stone", num_partitions_row=3, num_partitions_col=3,
                                      num_wood_zones=3, num_stone_zones=3, num_wood_and_stone_zones=2, **base)
    assert mz.spec["reset_mode"] == 1 and mz.spec["dyn_layout"] == 3 and mz.spec["mz_zones"] == [3, 3, 2]
    snap = foundation.make_env_instance("uniform/simple_wood_and_stone", device_reset="snapshot", **base)
    assert snap.spec["reset_mode"] == 0

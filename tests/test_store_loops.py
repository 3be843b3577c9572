"""The per-lane output writers of the device source (16-byte front-to-back stores with unaligned heads / tails, plane
boundaries inside a group) walked with 32 lanes on the host, lane by lane, against element-wise references - for the
default build and for the tuning variants kept behind macros (tests/emu/planes_check.cpp)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("flags", [[], ["-DAIE_PLANES_V2=1"]], ids=["default", "planes_v2"])
def test_store_loops_lane_by_lane(flags, tmp_path):
    exe = str(tmp_path / "planes_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17"] + flags + ["-o", exe, os.path.join(HERE, "emu", "planes_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and " 0 failures" in out.stdout, out.stdout[-2000:]

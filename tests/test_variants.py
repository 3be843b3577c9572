"""Tuning variants of the device source that are kept behind macros (tools/build_variants.py) must stay correct while the
default build moves on: each is compiled for the 1-lane emulation here and checked on the CPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _build_emu(tmp_path, flags):
    so = str(tmp_path / "libaie_emu_variant.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unused-function"] + flags +
                          ["-o", so, os.path.join(HERE, "emu", "aie_emu.cpp"), "-lm", "-lpthread"])
    return so


def test_fused_policy_variant_draws_only_unmasked_actions(tmp_path):
    so = _build_emu(tmp_path, ["-DAIE_FUSED_POLICY=1"])
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(HERE, "variants", "fused_policy_check.py"), so], capture_output=True,
                         text=True, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.count(" ok: ") == 4, out.stdout[-1500:] + out.stderr[-1500:]


def test_planes_v2_variant_replays_a_golden_trace(tmp_path):
    """The leaner bit-plane writer inside the whole step (its per-lane loop is checked with 32 lanes in test_store_loops)."""
    so = _build_emu(tmp_path, ["-DAIE_PLANES_V2=1"])
    code = ("import sys; from ai_economist_b200 import _abi; from tests.emu import emu_stepper as es; "
            "es._lib = _abi.load_library(sys.argv[1]); from tests import golden_utils as gu; "
            "from tests.stepper_adapters import GoldenStepperAdapter; "
            "p = [f for f in gu.golden_files() if 'c3_short_period' in f][0]; "
            "print(gu.replay(p, lambda spec, init: GoldenStepperAdapter(es.EmuStepper(spec, 1), init)))")
    out = subprocess.run([sys.executable, "-c", code, so], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT),
                         cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip().endswith("200"), out.stdout[-800:] + out.stderr[-1500:]

"""CPU: the C oracle (oracle/foundation_oracle.c) against the golden traces recorded from the
unmodified reference.  This is what "pins" the oracle (task brief ③)."""
import numpy as np
import pytest

from oracle.oracle import OracleBatch
from tests import golden_utils as gu


class OracleStepper:
    def __init__(self, spec, init):
        self.b = OracleBatch(spec, 1)
        self.b.load_env(0, init)

    def step(self, act_a, act_p):
        self.b.step(act_a[None], None if act_p is None else act_p[None])

    def obs(self):
        return self.b.obs(0)

    def state(self):
        return self.b.state(0)

    def books(self):
        return {(c, s): self.b.book(0, c, s) for c in (0, 1) for s in (0, 1)}


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: p.split("/")[-1])
def test_oracle_matches_reference_golden_trace(path):
    n = gu.replay(path, OracleStepper)
    assert n >= 50


def test_golden_fixtures_present():
    assert len(gu.golden_files()) >= 6

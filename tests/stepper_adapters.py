"""Adapters that let tests/golden_utils.replay drive a BatchStepper (CUDA product or the 1-lane emulation)."""
import numpy as np


class GoldenStepperAdapter:
    """One env at index 0 of a BatchStepper, loaded from a golden fixture's init_* snapshot."""

    def __init__(self, stepper, init, env_index=0):
        self.s = stepper
        self.e = env_index
        st = {k: np.asarray(v)[None] for k, v in init.items() if k not in ("mt_pos", "completions")}
        st["mt_pos"] = np.array([init["mt_pos"]], np.int32)
        st["completions"] = np.array([init.get("completions", 0)], np.int32)
        self.s.load_state(st, env_lo=env_index)

    def step(self, act_a, act_p):
        ba, bp = self.s.buf["actions_agent"], self.s.buf["actions_planner"]
        a = np.asarray(act_a, np.int32).reshape(ba.shape[1:])
        if isinstance(ba, np.ndarray):
            ba[self.e] = a
            if act_p is not None:
                bp[self.e] = np.asarray(act_p, np.int32)
        else:
            import torch
            ba[self.e] = torch.as_tensor(a, device=ba.device)
            if act_p is not None:
                bp[self.e] = torch.as_tensor(np.asarray(act_p, np.int32), device=bp.device)
        self.s.step()

    def obs(self):
        return self.s.read_obs(self.e)

    def state(self):
        self._st = self.s.read_state(self.e)
        return self._st

    def books(self):
        return self.s.read_state(self.e)["books"]

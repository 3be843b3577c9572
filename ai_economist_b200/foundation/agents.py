"""Agent descriptors (reference: base/base_agent.py, agents/mobiles.py, agents/planners.py).

In the batched stepper the agents' state lives in device tensors; these objects only describe the action
interface (subspace names/sizes in registration order, base_agent.py:97-222) and give read access to the
state of one (env, agent) pair for callers such as plotting code."""
import numpy as np

from .registrar import Registry


class BaseAgent:
    name = ""

    def __init__(self, idx, multi_action_mode):
        self._idx = idx
        self.multi_action_mode = bool(multi_action_mode)
        self._action_names = []
        self.action_dim = {}
        self._env = None

    @property
    def idx(self):
        return self._idx

    def _register(self, components):
        for comp in components:
            n = comp.get_n_actions(self.name)
            if n is None:
                continue
            subs = [(comp.name, n)] if isinstance(n, int) else [("{}.{}".format(comp.name, s), k) for s, k in n]
            for sub_name, k in subs:
                if k == 0:
                    continue
                self._action_names.append(sub_name)
                self.action_dim[sub_name] = k + (1 if self.multi_action_mode else 0)

    @property
    def action_spaces(self):
        """int (single-action: 1 + sum of subspace sizes) or array of per-subspace sizes (multi-action)."""
        if self.multi_action_mode:
            if not self._action_names:
                return np.array([1])  # PassiveAgentPlaceholder (base_agent.py:163-166)
            return np.array([self.action_dim[k] for k in self._action_names])
        return 1 + sum(self.action_dim.values())

    def get_random_action(self, rng=np.random):
        if self.multi_action_mode:
            return [int(rng.randint(0, self.action_dim[k])) for k in self._action_names]
        return int(rng.randint(0, self.action_spaces))

    # read-only state mirrors of env replica `env_index` (default 0)
    def state_of(self, env_index=0):
        return self._env._agent_state(self._idx, env_index)

    @property
    def state(self):
        return self.state_of(0)

    @property
    def loc(self):
        return self.state["loc"]

    @property
    def inventory(self):
        return self.state["inventory"]

    @property
    def escrow(self):
        return self.state["escrow"]

    def total_endowment(self, resource):
        s = self.state
        return s["inventory"][resource] + s["escrow"][resource]


agent_registry = Registry(BaseAgent)


@agent_registry.add
class BasicMobileAgent(BaseAgent):
    name = "BasicMobileAgent"


@agent_registry.add
class BasicPlanner(BaseAgent):
    name = "BasicPlanner"

    def __init__(self, multi_action_mode):
        super().__init__("p", multi_action_mode)

    @property
    def loc(self):
        raise AttributeError("BasicPlanner agents do not occupy a location.")

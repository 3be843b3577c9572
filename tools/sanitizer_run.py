"""Small product run for compute-sanitizer (racecheck / memcheck): tax config, tutorial config and a large-record config
(one CTA of four warps per env), a few envs and steps each, with the random policy fused into the step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_economist_b200 import foundation  # noqa: E402
from tests import batch_utils as bu  # noqa: E402

for cfg, E, steps in [("c3_short_period", 6, 40), ("c1_tutorial", 6, 30), ("c5_small", 3, 12)]:
    name, kw = bu.product_kwargs(cfg)
    env = foundation.make_env_instance(name, n_envs=E, device="cuda:0", seed=9, **kw)
    env.reset()
    for t in range(steps):
        if t == steps // 2:
            env.stepper.set_fused_policy(5)   # second half: the step kernel draws the actions itself
        if t < steps // 2:
            env.stepper.sample_random_actions(seed=t)
        env.stepper.step()
    import torch
    torch.cuda.synchronize()
    print(cfg, "ok", env.stepper.read_state(0)["t"])

# device-side layout generation (uniform: an auto-reset inside the run), the one-step-economy kernels, and the compacted
# host transfer (pack kernel + slices)
import ctypes as C  # noqa: E402
import numpy as np  # noqa: E402
from oracle import configs  # noqa: E402

kw = dict(configs.CONFIGS["uniform_reset"]); name = kw.pop("scenario_name"); kw["episode_length"] = 8
env = foundation.make_env_instance(name, n_envs=4, device="cuda:0", seed=3, auto_reset=True, **kw)
env.reset(); env.stepper.set_fused_policy(7)
for t in range(18):
    env.stepper.step()
torch.cuda.synchronize()
print("uniform_reset ok", env.stepper.read_state(0)["completions"])
env = foundation.make_env_instance("one-step-economy", n_envs=5, device="cuda:0", seed=2, auto_reset=True,
                                   components=[("SimpleLabor", {}), ("PeriodicBracketTax", dict(bracket_spacing="us-federal", period=2))],
                                   n_agents=6, world_size=[1, 1], episode_length=2)
env.reset(); env.stepper.set_fused_policy(9)
for t in range(6):
    env.stepper.step()
torch.cuda.synchronize()
print("one-step-economy ok")
name, kw = bu.product_kwargs("c1_tutorial")
env = foundation.make_env_instance(name, n_envs=40, device="cuda:0", seed=4, **kw)
env.reset()
st = env.stepper
out = {n: np.zeros(tuple(st.buf[n].shape), st.to_numpy(st.buf[n][:1]).dtype) for n in
       ["obs_agent_map", "obs_agent_idx", "obs_agent_flat", "mask_agent", "obs_planner_map", "obs_planner_idx", "obs_planner_flat",
        "obs_planner_agents", "mask_planner", "obs_time", "reward", "done"] if n in st.buf}
ptrs = {n: a.ctypes.data_as(C.c_void_p) for n, a in out.items()}
aa = np.zeros(tuple(st.buf["actions_agent"].shape), np.int32)
for t in range(3):
    st.step_host(aa.ctypes.data_as(C.c_void_p), None, ptrs, compact=True, n_threads=4)
print("compact host step ok")

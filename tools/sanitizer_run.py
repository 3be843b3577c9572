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

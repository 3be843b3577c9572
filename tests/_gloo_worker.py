"""Worker for tests/test_sharding_gloo.py (one process per rank, gloo, CPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_shard(seeds, steps):
    from ai_economist_b200 import foundation
    from tests import batch_utils as bu
    from tests.emu.emu_stepper import emu_factory

    name, kw = bu.product_kwargs("c1_tutorial")
    env = foundation.make_env_instance(name, n_envs=len(seeds), seeds=seeds, stepper_factory=emu_factory,
                                       auto_reset=False, **kw)
    env.reset()
    for t in range(steps):
        # actions are a deterministic function of the GLOBAL env seed, so shards and the full run agree
        acts = np.stack([np.random.RandomState(s * 1000 + t).randint(0, 50, size=(4, 1)) for s in seeds]).astype(np.int32)
        env.step((acts, None))
    st = env.stepper
    return np.stack([st.read_state(e)["loc"] for e in range(len(seeds))]), np.array(st.buf["reward"])


if __name__ == "__main__":
    import torch.distributed as dist

    from ai_economist_b200.sharding import max_over_ranks, shard_seeds

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = shard_seeds(1000, rank, world, 3)
    loc, rew = run_shard(seeds, 12)
    t = max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    np.savez(sys.argv[1] % rank, loc=loc, rew=rew, tmax=t, seeds=np.array(seeds))
    dist.destroy_process_group()

"""Debug aid (GPU box): one configuration of tests/test_fuzz_subsets.py::test_fuzz_cuda_matches_oracle, step by step, printing
the first difference in rewards / utility trackers between the CUDA build and the C oracle.  usage: python tools/debug_fuzz_case.py I"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_emu_vs_oracle as fz  # noqa: E402
from ai_economist_b200 import foundation  # noqa: E402
from oracle.oracle import OracleBatch  # noqa: E402
from tests import batch_utils as bu  # noqa: E402

i = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.RandomState(20260923)
cfgs = [fz.random_config(rng) for _ in range(20)]
name, kw = cfgs[i]
E = 3
dev = {} if "--emu" not in sys.argv else None
if dev is None:
    from tests.emu.emu_stepper import emu_factory
    env = foundation.make_env_instance(name, n_envs=E, stepper_factory=emu_factory, auto_reset=False, seed=1 + i, **kw)
else:
    env = foundation.make_env_instance(name, n_envs=E, device="cuda:0", auto_reset=False, seed=1 + i, **kw)
host = env.host_reset_arrays()
env.load_host_state(host)
orc = OracleBatch(env.spec, E)
for e in range(E):
    orc.load_env(e, {k: v[e] for k, v in host.items()})
st = env.stepper
spec = env.spec
seg_a, seg_p = bu.segments(spec, "a"), bu.segments(spec, "p")
r = np.random.RandomState(1 + i)
np.set_printoptions(precision=17, linewidth=200)
for t in range(1, 31):
    aa = bu.sample_from_masks(st.to_numpy(st.buf["mask_agent"]), seg_a, r)
    ap = bu.sample_from_masks(st.to_numpy(st.buf["mask_planner"]), seg_p, r) if seg_p else None
    env.step((aa, ap)); orc.step(aa, ap)
    for e in range(E):
        oo, os_ = orc.obs(e), orc.state(e)
        po, ps = st.read_obs(e), st.read_state(e)
        bad = []
        for k in ("rew",):
            if not np.allclose(oo[k], np.asarray(po[k]).reshape(oo[k].shape), rtol=1e-9, atol=1e-12): bad.append(k)
        for k in ("util_prev", "coin", "labor", "auto_warmup", "completions", "esc_coin"):
            if k in os_ and k in ps and not np.allclose(np.asarray(os_[k], float), np.asarray(ps[k], float).reshape(np.asarray(os_[k]).shape), rtol=1e-12, atol=1e-12): bad.append(k)
        if bad:
            print("t=%d env %d differs in %s" % (t, e, bad))
            for k in ("rew",):
                print("  oracle", k, oo[k]); print("  ours  ", k, np.asarray(po[k]))
            for k in ("util_prev", "coin", "labor", "auto_warmup", "esc_coin"):
                if k in os_ and k in ps:
                    print("  oracle", k, np.asarray(os_[k])); print("  ours  ", k, np.asarray(ps[k]))
            sys.exit(0)
print("no difference in 30 steps")

"""TEST INFRASTRUCTURE (run by tests/test_variants.py in a subprocess): the -DAIE_FUSED_POLICY=1 tuning variant of the
device source on the 1-lane emulation - after the first stand-alone sampler call the step's observation pass draws the next
step's random actions; every draw must be open in the masks that same step wrote.  argv[1] = the variant's emulation .so."""
import sys

import numpy as np
from tests.emu import emu_stepper as es
from ai_economist_b200 import _abi, foundation
es._lib = _abi.load_library(sys.argv[1])
from oracle.configs import CONFIGS
from tests import batch_utils as bu
for cfg in ["c1_tutorial", "tax_us_federal", "c3_reset", "tax_single_planner"]:
    kw = dict(CONFIGS[cfg]); name = kw.pop("scenario_name")
    env = foundation.make_env_instance(name, n_envs=6, stepper_factory=es.emu_factory, **kw)
    env.seed(list(range(6))); env.reset()
    st, spec = env.stepper, env.spec
    seg_a, seg_p = bu.segments(spec, "a"), bu.segments(spec, "p")
    st.sample_random_actions(seed=5)      # first call: the stand-alone sampler
    seen, prev = set(), None
    for it in range(40):
        aa = st.to_numpy(st.buf["actions_agent"]).copy(); ma = st.to_numpy(st.buf["mask_agent"])
        off = 0
        for si, n in enumerate(seg_a):
            pick = aa[..., si if spec["multi_action_agents"] else 0]
            assert np.all((pick >= 0) & (pick < n)), (cfg, it)
            assert np.all(np.take_along_axis(ma[..., off:off + n], pick[..., None], axis=-1) == 1.0), (cfg, it, si)
            off += n
            if not spec["multi_action_agents"]: break
        if seg_p:
            ap, mp = st.to_numpy(st.buf["actions_planner"]), st.to_numpy(st.buf["mask_planner"])
            off = 0
            for b, n in enumerate(seg_p):
                assert np.all(np.take_along_axis(mp[:, off:off + n], ap[:, b][:, None], axis=-1) == 1.0), (cfg, it, "planner")
                off += n
        seen.update(np.unique(aa).tolist())
        assert prev is None or not np.array_equal(prev, aa)
        prev = aa
        n0 = st.launch_count()
        st.sample_random_actions(seed=100 + it)   # no-op in the variant apart from refreshing the seed
        assert st.launch_count() == n0
        env.step(env.action_buffers)              # the step's observation pass draws the next actions
    print(cfg, "ok: %d distinct action values" % len(seen))

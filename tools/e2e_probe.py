"""Where the end-to-end (host buffer) step spends its time: aie_step_host_compact at several host-thread counts, with
plain and NUMA-interleaved pinned output tensors.  Run on the GPU box.  usage: python tools/e2e_probe.py [c2|c3|c5]"""
import ctypes as C
import json
import os
import sys
import time

os.environ.setdefault("AIE_E2E_REPEAT_EXPAND", "2")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_economist_b200 import foundation, hostmem, workloads as wl   # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = {"c2": ("c1_tutorial", 8192), "c3": ("c3_paper_tax", 8192), "c5": ("c5_full", 2048)}[key]
name, kw = wl.product_kwargs(cfg[0])
env = foundation.make_env_instance(name, n_envs=cfg[1], device="cuda:0", seed=1000, auto_reset=True, **kw)
env.reset()
st = env.stepper
d = st.dims
names = ["obs_agent_map", "obs_agent_idx", "obs_agent_flat", "mask_agent", "obs_planner_map", "obs_planner_idx",
         "obs_planner_flat", "obs_planner_agents", "mask_planner", "obs_time", "reward", "done"]
seg_a, seg_p = wl.mask_segments(env.spec, "a"), wl.mask_segments(env.spec, "p")
out = {"workload": key, "runs": []}
for interleave, numa_mode in (("plain", "2"), ("split", "2"), ("split", "1")):
    os.environ["AIE_E2E_NUMA"] = numa_mode
    host = {n: hostmem.pinned_empty(st.buf[n].shape, st.buf[n].dtype, numa=(interleave if interleave != "plain" else None))
            for n in names if n in st.buf}
    ptrs = {n: C.c_void_p(t.data_ptr()) for n, t in host.items()}
    act_a = torch.zeros(st.buf["actions_agent"].shape, dtype=torch.int32, pin_memory=True)
    act_p = torch.zeros(st.buf["actions_planner"].shape, dtype=torch.int32, pin_memory=True)
    host["mask_agent"].copy_(st.buf["mask_agent"]); host["mask_planner"].copy_(st.buf["mask_planner"])
    rng = np.random.RandomState(0)
    for threads in (32, 64, 96, 128):
        ts, tim = [], []
        for i in range(8):
            act_a.copy_(torch.from_numpy(wl.sample_from_masks(host["mask_agent"].numpy(), seg_a, rng)))
            if seg_p:
                act_p.copy_(torch.from_numpy(wl.sample_from_masks(host["mask_planner"].numpy(), seg_p, rng)))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st.step_host(C.c_void_p(act_a.data_ptr()), C.c_void_p(act_p.data_ptr()) if d.n_act_planner else None, ptrs,
                         compact=True, n_threads=threads)
            ts.append(time.perf_counter() - t0)
            tim.append(st.host_timing())
        med = sorted(ts[2:])[len(ts[2:]) // 2]
        t = tim[-1]
        r = dict(alloc=interleave, numa_mode=numa_mode, threads=threads, ms=1e3 * med, rate=cfg[1] * env.n_agents / med,
                 before_transfer=t["before_transfer"], first_slice=t["first_slice"], last_slice=t["last_slice"], expanded=t["expanded"],
                 d2h_MB=t["d2h_bytes"] / 1e6, wait_sum=t["wait_sum"], busy_sum=t["busy_sum"], first_dev=t["first_slice_dev"],
                 last_dev=t["last_slice_dev"], expand_only=t["expand_only"])
        out["runs"].append(r)
        print(json.dumps(r), flush=True)
    del host, ptrs
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_probe_%s.json" % key, "w"), indent=1)

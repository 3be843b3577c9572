"""Thin Python driver over the C-ABI: allocates the device buffers (PyTorch-owned), binds them, uploads host
reset snapshots and launches steps.  Plumbing only — all compute is in the CUDA library."""
import ctypes as C

import numpy as np

from . import _abi


class BatchStepper:
    """Backend-agnostic part (buffer bookkeeping + ABI calls).  Subclasses supply allocation.

    Reference counterpart: FoundationEnvWrapper (ai_economist/foundation/env_wrapper.py:84-418), which owns
    the WarpDrive data manager / function manager for the COVID env.
    """

    _DTYPES = {"u8": np.uint8, "i16": np.int16, "i32": np.int32, "f32": np.float32, "f64": np.float64}

    def __init__(self, spec, n_envs, lib, device_index=0, auto_reset=True, event_envs=0):
        self.spec = dict(spec)
        self.n_envs = int(n_envs)
        self.lib = lib
        self.cfg = _abi.config_from_spec(spec, auto_reset=auto_reset)
        h = C.c_void_p()
        self._check(lib.aie_create(C.byref(self.cfg), self.n_envs, int(device_index), C.byref(h)))
        self._h = h
        self.dims = _abi.AieDims()
        self._check(lib.aie_get_dims(self._h, C.byref(self.dims)))
        d, E, A = self.dims, self.n_envs, spec["n_agents"]
        H, W = spec["height"], spec["width"]
        full = bool(spec.get("full_observability", 0))   # agents see the whole map: [M, H, W] / [2, H, W] per agent
        none = bool(spec.get("scenario_kind", 0) == 1)   # one-step-economy: no spatial observations (zero-size tensors)
        assert d.agent_map_elems == (0 if none else d.n_map_channels * H * W if full else (d.n_map_channels + 1) * d.window ** 2)
        shapes = {
            "state": ("u8", (E, d.state_bytes)), "state0": ("u8", (E, d.state_bytes)),
            "actions_agent": ("i32", (E, A, d.n_act_agent)),
            "actions_planner": ("i32", (E, max(1, d.n_act_planner))),
            "obs_agent_map": ("f32", (E, A, 0) if none else (E, A, d.n_map_channels, H, W) if full else
                              (E, A, d.n_map_channels + 1, d.window, d.window)),
            "obs_agent_idx": ("i16", (E, A, 0) if none else (E, A, 2, H, W) if full else (E, A, 2, d.window, d.window)),
            "obs_agent_flat": ("f32", (E, A, d.flat_agent)),
            "mask_agent": ("f32", (E, A, d.mask_agent)),
            "obs_planner_map": ("f32", (E, d.n_map_channels, H, W)),
            "obs_planner_idx": ("i16", (E, 2, H, W)),
            "obs_planner_flat": ("f32", (E, d.flat_planner)),
            "obs_planner_agents": ("f32", (E, A, d.flat_planner_agent)),
            "mask_planner": ("f32", (E, d.mask_planner)),
            "obs_time": ("f32", (E,)),
            "reward": ("f64", (E, A + 1)),
            "done": ("i32", (E,)),
        }
        if auto_reset:  # end-of-episode record snapshots (previous_episode_metrics)
            shapes["episode_final"] = ("u8", (E, d.state_bytes))
        if not spec["planner_gets_spatial_info"]:
            del shapes["obs_planner_map"], shapes["obs_planner_idx"]
        # per-step event log of the first event_envs replicas (dense logs): capacity = every agent builds, gathers both
        # resources and every resting order trades
        self.event_envs = max(0, min(int(event_envs), E))
        K = int(spec.get("max_num_orders", 0) or 0) if "ContinuousDoubleAuction" in spec["components"] else 0
        self.event_cap = int(min(1024, 3 * A + 2 * A * K + 8))
        if self.event_envs:
            shapes["events"] = ("i32", (self.event_envs, self.event_cap + 1, 8))
        self.buf = {k: self._alloc(shape, dt) for k, (dt, shape) in shapes.items()}
        bufs = _abi.AieBuffers()
        for name in _abi._BUF_NAMES:
            ok = name in self.buf and int(np.prod(self.buf[name].shape)) > 0   # zero-size tensors have no storage
            setattr(bufs, name, self._ptr(self.buf[name]) if ok else None)
        bufs.events = self._ptr(self.buf["events"]) if self.event_envs else None
        bufs.event_envs, bufs.event_cap = self.event_envs, self.event_cap
        self._check(lib.aie_bind_buffers(self._h, C.byref(bufs)))
        self._fields = {}

    # -- to be provided by subclasses ------------------------------------------------------------------
    def _alloc(self, shape, dt):
        raise NotImplementedError

    def _ptr(self, buf):
        raise NotImplementedError

    def _stream(self):
        return None

    def to_numpy(self, buf):
        raise NotImplementedError

    # --------------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != _abi.AIE_OK:
            raise _abi.AieError("aie error %d: %s" % (rc, self.lib.aie_last_error().decode()))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aie_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def flat_layout(self, which):
        """[(key, offset, size)] of the fields concatenated into a flat vector (base_env.py:562-589);
        which: "agent" | "planner" | "planner_agent"."""
        idx = {"agent": 0, "planner": 1, "planner_agent": 2}[which]
        n = self.lib.aie_get_flat_layout(self._h, idx, None, 0)
        if n < 0:
            self._check(n)
        arr = (_abi.AieFlatField * max(1, n))()
        self.lib.aie_get_flat_layout(self._h, idx, arr, n)
        return [(arr[i].key.decode(), int(arr[i].offset), int(arr[i].size)) for i in range(n)]

    def load_state(self, host_state, env_lo=0):
        """host_state: dict of numpy arrays with a leading env axis (see aie_host_state in the header)."""
        n = int(np.asarray(host_state["loc"]).shape[0])
        keep = []

        def arr(key, dt, required=True):
            v = host_state.get(key)
            if v is None:
                assert not required, key
                return None
            a = np.ascontiguousarray(v, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        hs = _abi.AieHostState()
        hs.n = n
        for key, dt, req in [("stone", np.uint8, True), ("wood", np.uint8, True), ("stone_src", np.uint8, True),
                             ("wood_src", np.uint8, True), ("water", np.uint8, False), ("loc", np.int16, True),
                             ("coin", np.float64, True), ("inv_stone", np.int32, False), ("inv_wood", np.int32, False),
                             ("build_payment", np.float64, True), ("build_skill", np.float64, True),
                             ("bonus_gather_prob", np.float64, True), ("mt_key", np.uint32, True),
                             ("mt_pos", np.int32, True), ("completions", np.int32, False),
                             ("gauss_has", np.int32, False), ("gauss_val", np.float64, False),
                             ("split_skill", np.float64, False)]:
            setattr(hs, key, arr(key, dt, req))
        self._check(self.lib.aie_load_state(self._h, C.byref(hs), int(env_lo), self._stream()))

    def step(self):
        """Advance every env by one timestep from the bound action buffers (asynchronous on the stream)."""
        self._check(self.lib.aie_step(self._h, self._stream()))

    def observe(self):
        self._check(self.lib.aie_observe(self._h, self._stream()))

    def step_dynamics(self):
        self._check(self.lib.aie_step_dynamics(self._h, self._stream()))

    def sample_random_actions(self, seed=0):
        """Device-side random policy: one uniformly random unmasked action per agent/subspace."""
        self._check(self.lib.aie_sample_random_actions(self._h, C.c_uint64(int(seed)), self._stream()))

    def set_fused_policy(self, seed):
        """seed != 0: every following step() also draws the NEXT step's uniformly random unmasked actions into the action
        buffers inside the observation pass (no sampler launch per step); 0 turns it off (aie_set_fused_policy)."""
        self._check(self.lib.aie_set_fused_policy(self._h, C.c_uint64(int(seed)), self._stream()))

    def step_host(self, actions_agent, actions_planner, out_ptrs, compact=False, n_threads=0):
        """End-to-end step with HOST buffers (aie_step_host).  out_ptrs: dict name -> host pointer / None.
        compact=True: same bytes in the host tensors, bit- / byte-packed over PCIe (aie_step_host_compact)."""
        o = _abi.AieHostOut()
        for name in _abi._OUT_NAMES:
            setattr(o, name, out_ptrs.get(name))
        if compact:
            self._check(self.lib.aie_step_host_compact(self._h, actions_agent, actions_planner, C.byref(o),
                                                       int(n_threads), self._stream()))
        else:
            self._check(self.lib.aie_step_host(self._h, actions_agent, actions_planner, C.byref(o), self._stream()))

    def host_timing(self):
        """Host-clock breakdown (ms) of the last compacted step_host call (aie_get_host_timing)."""
        out = (C.c_double * 16)()
        self.lib.aie_get_host_timing(self._h, out, 16)
        k = ["enqueued", "first_slice", "last_slice", "expanded", "slices", "threads", "d2h_bytes", "before_transfer",
             "wait_sum", "busy_sum", "first_slice_dev", "last_slice_dev", "expand_only", "overflow_envs", "chunks", "staging_node"]
        return dict(zip(k, list(out)))

    def compact_bytes_per_env(self):
        return int(self.lib.aie_compact_bytes_per_env(self._h))

    def launch_count(self):
        return int(self.lib.aie_launch_count(self._h))

    def field(self, name):
        if name not in self._fields:
            f = _abi.AieField()
            self._check(self.lib.aie_get_field(self._h, name.encode(), C.byref(f)))
            self._fields[name] = f
        return self._fields[name]

    def read_state(self, e, final=False):
        """Readback of env e in the test oracle's layout (dict of numpy arrays + 'books').  final=True reads the
        end-of-episode snapshot of the env's last finished episode instead (auto-reset only; no order books)."""
        A, H, W = self.spec["n_agents"], self.spec["height"], self.spec["width"]
        P = self.spec["max_bid_ask"] + 1 if "ContinuousDoubleAuction" in self.spec["components"] else 1
        B = max(1, self.spec["n_brackets"])
        cap = A * max(1, self.spec["max_num_orders"])
        out = dict(
            cell=np.zeros((H, W), np.uint8), owner=np.zeros((H, W), np.int8), loc=np.zeros((A, 2), np.int16),
            coin=np.zeros(A), esc_coin=np.zeros(A), labor=np.zeros(A),
            inv=np.zeros((A, 2), np.int32), esc=np.zeros((A, 2), np.int32),
            n_orders=np.zeros((2, A), np.int32), bid_hist=np.zeros((2, A, P), np.int32),
            ask_hist=np.zeros((2, A, P), np.int32), price_hist=np.zeros((2, A, P)),
            tax_pos=np.zeros(1, np.int32), rate_idx=np.zeros(B, np.int32),
            last_coin=np.zeros(A), last_income=np.zeros(A), last_marg=np.zeros(A),
            mt_key=np.zeros(624, np.uint32), mt_pos=np.zeros(1, np.int32), t=np.zeros(1, np.int32),
            completions=np.zeros(1, np.int32),
            book_rows=np.zeros((2, 2, cap, 3), np.int32), book_count=np.zeros((2, 2), np.int32),
            stats=np.zeros(self.dims.n_stats), util_prev=np.zeros(A + 1), auto_warmup=np.zeros(1, np.int32),
        )
        d = _abi.AieStateDump()
        for k in _abi._DUMP_PTRS + _abi._DUMP_PTRS2:
            setattr(d, k, out[k].ctypes.data_as(C.c_void_p))
        d.book_cap = cap
        fn = self.lib.aie_read_episode_final if final else self.lib.aie_read_state
        self._check(fn(self._h, int(e), C.byref(d)))
        out["books"] = {(c, s): out["book_rows"][c, s, :out["book_count"][c, s]].copy()
                        for c in (0, 1) for s in (0, 1)}
        if hasattr(self, "state_view"):   # skills / payments (one-step-economy: SimpleLabor skill / cumulative production)
            for k in ("build_payment", "build_skill"):
                out[k] = np.asarray(self.to_numpy(self.state_view(k, final=final)[e]), np.float64).copy()
        return out

    def read_events(self, e):
        """Events of replica e's last step (dense logs): list of int rows [kind, ...] (include/aie_b200.h)."""
        blk = self.to_numpy(self.buf["events"][e])
        n, dropped = int(blk[0, 0]), int(blk[0, 2])
        if dropped:
            raise _abi.AieError("event log overflow: %d events dropped (capacity %d)" % (dropped, self.event_cap))
        return [tuple(int(v) for v in blk[1 + i]) for i in range(n)]

    def read_obs(self, e):
        """Host copy of env e's outputs in the oracle's obs layout (tests)."""
        g = lambda k: self.to_numpy(self.buf[k][e]) if k in self.buf else None
        out = dict(a_map=g("obs_agent_map"), a_idx=g("obs_agent_idx"), a_flat=g("obs_agent_flat"),
                   a_mask=g("mask_agent"), p_flat=g("obs_planner_flat"), p_agents=g("obs_planner_agents"),
                   p_mask=g("mask_planner"), time=np.atleast_1d(g("obs_time")), rew=g("reward"),
                   done=np.atleast_1d(g("done")))
        if "obs_planner_map" in self.buf:
            out["p_map"], out["p_idx"] = g("obs_planner_map"), g("obs_planner_idx")
        return out


class CudaStepper(BatchStepper):
    """The product path: torch-owned CUDA tensors, kernels launched on torch's current stream."""

    _TORCH = None

    def __init__(self, spec, n_envs, device="cuda:0", auto_reset=True, lib_path=None, event_envs=0):
        import torch

        if not torch.cuda.is_available():
            raise _abi.AieError("no CUDA device visible: ai_economist_b200 has no CPU fallback")
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.AieError("device must be a CUDA device, got %s" % device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        lib = _abi.load_library(lib_path)
        with torch.cuda.device(self.device):
            super().__init__(spec, n_envs, lib, device_index=idx, auto_reset=auto_reset, event_envs=event_envs)

    def _alloc(self, shape, dt):
        t = self.torch
        m = {"u8": t.uint8, "i16": t.int16, "i32": t.int32, "f32": t.float32, "f64": t.float64}
        return t.zeros(shape, dtype=m[dt], device=self.device)

    def _ptr(self, buf):
        return C.c_void_p(buf.data_ptr())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def to_numpy(self, buf):
        return buf.detach().cpu().numpy()

    def state_view(self, name, final=False):
        """Strided struct-of-arrays view [E, ...] of one state field inside the packed records (final=True: inside the
        end-of-episode snapshots of the last finished episode, auto-reset only)."""
        t = self.torch
        f = self.field(name)
        shape = [f.shape[i] for i in range(f.ndim)]
        n = int(np.prod(shape)) if shape else 1
        dt = {(1, 0, 0): t.uint8, (1, 0, 1): t.int8, (2, 0, 1): t.int16, (4, 0, 1): t.int32, (4, 0, 0): t.int32,
              (8, 1, 1): t.float64}[(f.elem_bytes, f.is_float, f.is_signed)]
        raw = self.buf["episode_final" if final else "state"][:, f.offset:f.offset + n * f.elem_bytes]
        return raw.view(dt).view([self.n_envs] + shape) if shape else raw.view(dt)[:, 0]

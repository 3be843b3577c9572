"""Error behaviour of the C-ABI (include/aie_b200.h): every entry point returns AIE_OK or a negative AIE_E* code and
aie_last_error() says why - no exceptions, no device asserts (the reference's WarpDrive convention has neither return
codes nor messages, env_wrapper.py:230-252).  Runs against the emulation build, which compiles the same aie_abi.inl /
aie_host.h (validation, call-order checks, record packing) as the CUDA library."""
import copy
import ctypes as C

import numpy as np
import pytest

from ai_economist_b200 import _abi
from tests import golden_utils as gu
from tests.emu.emu_stepper import EmuStepper, emu_lib

AIE_EINVAL, AIE_ESTATE = -1, -3


def _spec():
    z, meta, init = gu.load_fixture([p for p in gu.golden_files() if "c3_short_period" in p][0])
    return meta["spec"], init


def _create(spec, n_envs=2, mutate=None):
    L = emu_lib()
    cfg = _abi.config_from_spec(spec, auto_reset=False)
    if mutate:
        mutate(cfg)
    h = C.c_void_p()
    rc = L.aie_create(C.byref(cfg), n_envs, 0, C.byref(h))
    msg = L.aie_last_error().decode()
    if rc == 0:
        L.aie_destroy(h)
    return rc, msg


@pytest.mark.parametrize("mutate,needle", [
    (lambda c: setattr(c, "abi_version", 1), "abi_version"),
    (lambda c: setattr(c, "n_agents", 1), "n_agents"),
    (lambda c: setattr(c, "n_agents", 65), "n_agents"),
    (lambda c: setattr(c, "height", 300), "world size"),
    (lambda c: setattr(c, "episode_length", 0), "episode_length"),
    (lambda c: setattr(c, "n_components", 0), "components"),
    (lambda c: c.components.__setitem__(0, 9), "unknown component"),
    (lambda c: c.components.__setitem__(1, c.components[0]), "duplicate component"),
    (lambda c: setattr(c, "obs_range", 40), "observation_range"),
    (lambda c: c.regen_weight.__setitem__(0, 1.5), "regen weight"),
    (lambda c: c.regen_halfwidth.__setitem__(1, 4), "regen_halfwidth"),
    (lambda c: setattr(c, "isoelastic_eta", 1.5), "isoelastic_eta"),
    (lambda c: setattr(c, "planner_reward_type", 7), "planner_reward_type"),
    (lambda c: setattr(c, "max_bid_ask", 40), "max_bid_ask"),
    (lambda c: setattr(c, "order_duration", 0), "order_duration"),
    (lambda c: setattr(c, "max_num_orders", 300), "max_num_orders"),
    (lambda c: setattr(c, "tax_model", 5), "tax_model"),
    (lambda c: setattr(c, "n_brackets", 1), "n_brackets"),
    (lambda c: setattr(c, "n_disc_rates", 0), "n_disc_rates"),
    (lambda c: setattr(c, "period", 0), "period"),
    (lambda c: setattr(c, "reset_mode", 2), "reset_mode"),
])
def test_create_rejects_bad_configuration_with_a_message(mutate, needle):
    spec, _ = _spec()
    rc, msg = _create(spec, mutate=mutate)
    assert rc == AIE_EINVAL and needle in msg, (rc, msg)


def test_create_rejects_bad_batch_and_null_arguments():
    spec, _ = _spec()
    L = emu_lib()
    assert _create(spec, n_envs=0)[0] == AIE_EINVAL
    cfg = _abi.config_from_spec(spec)
    assert L.aie_create(C.byref(cfg), 1, 0, None) == AIE_EINVAL
    assert L.aie_step(None, None) == AIE_EINVAL and L.aie_get_dims(None, None) == AIE_EINVAL
    assert L.aie_destroy(None) == 0                       # destroying nothing is fine
    assert _create(spec)[0] == 0                          # and the unmodified config is accepted


def test_single_action_planner_mask_longer_than_the_sampler_supports_is_rejected():
    spec, _ = _spec()
    spec = dict(spec, single_action_planner=1)
    rc, msg = _create(spec, mutate=lambda c: setattr(c, "n_disc_rates", 40))   # 1 + 7 * 40 > 160
    assert rc == AIE_EINVAL and "single-action planner" in msg


def test_calls_out_of_order_return_estate():
    spec, init = _spec()
    L = emu_lib()
    cfg = _abi.config_from_spec(spec, auto_reset=False)
    h = C.c_void_p()
    assert L.aie_create(C.byref(cfg), 2, 0, C.byref(h)) == 0
    hs = _abi.AieHostState()
    assert L.aie_load_state(h, C.byref(hs), 0, None) == AIE_ESTATE and b"not bound" in L.aie_last_error()
    assert L.aie_step(h, None) == AIE_ESTATE
    assert L.aie_observe(h, None) == AIE_ESTATE and L.aie_sample_random_actions(h, 1, None) == AIE_ESTATE
    bufs = _abi.AieBuffers()                              # all NULL
    assert L.aie_bind_buffers(h, C.byref(bufs)) == AIE_EINVAL and b"NULL" in L.aie_last_error()
    dump = _abi.AieStateDump()
    assert L.aie_read_state(h, 0, C.byref(dump)) == AIE_ESTATE
    f = _abi.AieField()
    assert L.aie_get_field(h, b"no_such_field", C.byref(f)) == AIE_EINVAL and b"unknown state field" in L.aie_last_error()
    assert L.aie_get_flat_layout(h, 3, None, 0) == AIE_EINVAL
    assert L.aie_get_flat_layout(h, 0, None, 0) > 0
    L.aie_destroy(h)


def test_bound_but_not_loaded_and_bad_load_arguments():
    spec, init = _spec()
    st = EmuStepper(spec, 2)
    L = st.lib
    assert L.aie_step(st._h, None) == AIE_ESTATE and b"load state first" in L.aie_last_error()
    hs = _abi.AieHostState()
    hs.n = 1
    assert L.aie_load_state(st._h, C.byref(hs), 0, None) == AIE_EINVAL and b"NULL" in L.aie_last_error()
    hs.n = 5
    assert L.aie_load_state(st._h, C.byref(hs), 0, None) == AIE_EINVAL and b"out of bounds" in L.aie_last_error()
    dump = _abi.AieStateDump()
    assert L.aie_read_state(st._h, 7, C.byref(dump)) == AIE_EINVAL
    assert L.aie_read_episode_final(st._h, 0, C.byref(dump)) == AIE_ESTATE       # auto_reset off: no snapshot buffer
    # a location outside the world is refused by the record packer instead of corrupting the map
    bad = {k: np.asarray(v)[None].repeat(2, axis=0) if not np.isscalar(v) else np.full(2, v) for k, v in init.items()}
    bad["loc"] = bad["loc"].copy()
    bad["loc"][1, 0] = (99, 0)
    with pytest.raises(_abi.AieError):
        st.load_state(bad)


def _covid_cfg():
    import json, os
    from ai_economist_b200.covid_stepper import covid_config_from_params
    from ai_economist_b200.foundation.covid19 import build_covid_params
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_covid", "covid_seed3.npz"))
    return covid_config_from_params(build_covid_params(**json.loads(str(z["meta_json"]))["kwargs"]), auto_reset=False)


@pytest.mark.parametrize("mutate,needle", [
    (lambda c: setattr(c, "abi_version", 2), "abi_version"),
    (lambda c: setattr(c, "n_states", 65), "n_states"),
    (lambda c: setattr(c, "num_filters", 9), "num_filters"),
    (lambda c: setattr(c, "beta_delay", 10 ** 6), "filter / delay"),
    (lambda c: setattr(c, "subsidy_interval", 0), "schedule"),
    (lambda c: setattr(c, "num_stringency_levels", 1), "schedule"),
    (lambda c: setattr(c, "start_date_index", 10 ** 6), "start_date_index"),
    (lambda c: setattr(c, "conv_filters", None), "NULL"),
])
def test_covid_create_rejects_bad_configuration_with_a_message(mutate, needle):
    L = emu_lib()
    cfg, keep = _covid_cfg()
    mutate(cfg)
    h = C.c_void_p()
    rc = L.aie_covid_create(C.byref(cfg), 2, 0, C.byref(h))
    assert rc == AIE_EINVAL and needle in L.aie_last_error().decode(), L.aie_last_error()


def test_covid_calls_out_of_order_return_estate():
    L = emu_lib()
    cfg, keep = _covid_cfg()
    h = C.c_void_p()
    assert L.aie_covid_create(C.byref(cfg), 0, 0, C.byref(h)) == AIE_EINVAL
    assert L.aie_covid_create(C.byref(cfg), 2, 0, C.byref(h)) == 0
    assert L.aie_covid_reset(h, None) == AIE_ESTATE and L.aie_covid_step(h, None) == AIE_ESTATE
    assert L.aie_covid_sample_random_actions(h, 1, None) == AIE_ESTATE
    bufs = _abi.AieCovidBuffers()
    assert L.aie_covid_bind_buffers(h, C.byref(bufs)) == AIE_EINVAL and b"NULL" in L.aie_last_error()
    assert L.aie_covid_step(None, None) == AIE_EINVAL and L.aie_covid_destroy(None) == 0
    assert L.aie_covid_destroy(h) == 0

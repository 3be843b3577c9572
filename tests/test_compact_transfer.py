"""aie_step_host_compact: the caller's host tensors must receive exactly the bytes aie_step_host delivers, for every
output tensor, whatever the layout (window / full observability, spatial planner or not, no p<i> vectors, odd sizes that
do not fill a 32-bit word).  CPU: emulation build (same aie_abi.inl, pack / expand code and thread pool); GPU: CUDA."""
import ctypes as C

import numpy as np
import pytest

from ai_economist_b200 import _abi, foundation
from oracle import configs
from tests import batch_utils as bu

CASES = ["c1_tutorial", "c3_paper_tax", "tax_us_federal", "full_obs_tax", "tax_single_planner"]


def _env(cfg, n_envs, factory, device):
    allc = dict(configs.CONFIGS)
    allc.update(configs.EDGE_CONFIGS)
    kw = dict(allc[cfg])
    kw.pop("seed", None)
    name = kw.pop("scenario_name")
    extra = dict(stepper_factory=factory) if factory is not None else dict(device=device)
    env = foundation.make_env_instance(name, n_envs=n_envs, auto_reset=True, seed=31, **kw, **extra)
    env.reset()
    return env


def _host_outputs(st):
    out, ptrs = {}, {}
    for nm in _abi._OUT_NAMES:
        if nm in st.buf and int(np.prod(st.buf[nm].shape)) > 0:
            a = np.full(tuple(st.buf[nm].shape), 77, dtype=st.to_numpy(st.buf[nm][:1]).dtype)   # poisoned
            out[nm] = a
            ptrs[nm] = a.ctypes.data_as(C.c_void_p)
    return out, ptrs


def _check(cfg, n_envs, factory=None, device=None, threads=(1, 3, 0)):
    plain, compact = _env(cfg, n_envs, factory, device), _env(cfg, n_envs, factory, device)
    sp, sc = plain.stepper, compact.stepper
    assert 0 < sc.compact_bytes_per_env() < sum(int(np.prod(sp.buf[n].shape[1:])) * sp.to_numpy(sp.buf[n][:1]).itemsize
                                                for n in _abi._OUT_NAMES if n in sp.buf)
    seg_a, seg_p = bu.segments(plain.spec, "a"), bu.segments(plain.spec, "p")
    rng = np.random.RandomState(2)
    op, pp = _host_outputs(sp)
    oc, pc = _host_outputs(sc)
    for t in range(12):
        aa = np.ascontiguousarray(bu.sample_from_masks(sp.to_numpy(sp.buf["mask_agent"]), seg_a, rng), np.int32)
        ap = np.ascontiguousarray(bu.sample_from_masks(sp.to_numpy(sp.buf["mask_planner"]), seg_p, rng), np.int32) \
            if seg_p else None
        pa = aa.ctypes.data_as(C.c_void_p)
        ppn = ap.ctypes.data_as(C.c_void_p) if ap is not None else None
        sp.step_host(pa, ppn, pp)
        sc.step_host(pa, ppn, pc, compact=True, n_threads=threads[t % len(threads)])
        for nm in op:
            assert np.array_equal(op[nm], oc[nm]), "%s step %d: %s differs" % (cfg, t, nm)
            assert np.array_equal(op[nm], sp.to_numpy(sp.buf[nm])), nm      # and both equal the device tensors


@pytest.mark.parametrize("cfg", CASES + ["full_obs_plain", "two_agents_w0", "nonsquare_33_agents"])
def test_emulated_compact_transfer_delivers_the_same_bytes(cfg):
    from tests.emu.emu_stepper import emu_factory
    _check(cfg, 5, factory=emu_factory)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CASES)
def test_cuda_compact_transfer_delivers_the_same_bytes(cfg):
    _check(cfg, 300, device="cuda:0")


def test_emulated_compact_transfer_index_plane_overflow_is_fetched_directly(monkeypatch):
    """An env whose index planes hold more non-zero elements than the compact record carries is completed by a direct
    copy of those planes (capacities forced down to 4 entries here, so every env takes that path)."""
    from tests.emu.emu_stepper import emu_factory
    monkeypatch.setenv("AIE_COMPACT_TINY_CAPS", "1")
    _check("c3_paper_tax", 4, factory=emu_factory, threads=(2,))


@pytest.mark.gpu
def test_cuda_compact_transfer_index_plane_overflow_is_fetched_directly(monkeypatch):
    monkeypatch.setenv("AIE_COMPACT_TINY_CAPS", "1")
    _check("c1_tutorial", 64, device="cuda:0", threads=(4,))


@pytest.mark.parametrize("chunks", [1, 3])
def test_emulated_compact_transfer_chunked_stepping(monkeypatch, chunks):
    """The batch may step in several launches over env ranges (AIE_E2E_CHUNKS) so that early slices go down while the rest
    still steps: same bytes whatever the chunk count (70 envs = 5 work items of 16 envs = 5 slices)."""
    from tests.emu.emu_stepper import emu_factory
    monkeypatch.setenv("AIE_E2E_CHUNKS", str(chunks))
    _check("c1_tutorial", 70, factory=emu_factory, threads=(2,))


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [1, 2, 5, 16])
def test_cuda_compact_transfer_chunked_stepping(monkeypatch, chunks):
    """1 000 envs = 63 work items = 16 slices; ragged last chunk with 5 chunks; one chunk per slice with 16."""
    monkeypatch.setenv("AIE_E2E_CHUNKS", str(chunks))
    _check("c3_paper_tax", 1000, device="cuda:0", threads=(0,))


@pytest.mark.parametrize("n_envs", [5, 70, 300])
def test_hostmem_block_rows_match_the_transfer_slices(n_envs):
    """hostmem.transfer_block_rows restates the library's slicing (work items of 16 envs, at most 16 slices): the NUMA
    blocks of pinned_empty(numa="blocks") must be exactly the transfer slices."""
    from ai_economist_b200 import hostmem
    from tests.emu.emu_stepper import emu_factory
    env = _env("c1_tutorial", n_envs, emu_factory, None)
    st = env.stepper
    out, ptrs = _host_outputs(st)
    aa = np.zeros(tuple(st.buf["actions_agent"].shape), np.int32)
    st.step_host(aa.ctypes.data_as(C.c_void_p), None, ptrs, compact=True, n_threads=2)
    rows = hostmem.transfer_block_rows(n_envs)
    assert rows % 16 == 0
    assert int(st.host_timing()["slices"]) == -(-n_envs // rows)


def test_hostmem_block_placement_covers_the_buffer_and_alternates_nodes():
    """hostmem._placement: page-aligned pieces covering the whole mapping; with numa="blocks" piece b holds block b's rows
    (up to half a page at either edge) and goes to node b % nodes; tiny tensors degrade to whole pages."""
    from ai_economist_b200 import hostmem as h
    page, row = 4096, 13552                                   # c2 agent map: 8 192 rows of 13 552 bytes, slices of 512 rows
    size = (8192 * row + page - 1) // page * page
    pc = h._placement(size, page, 2, "blocks", 8192, row, 512)
    assert len(pc) == 16 and [k for _, _, k in pc] == [b % 2 for b in range(16)]
    assert pc[0][0] == 0 and pc[-1][1] == size and all(a[1] == b[0] for a, b in zip(pc, pc[1:]))
    assert all(lo % page == 0 and abs(lo - b * 512 * row) <= page // 2 for b, (lo, _, _) in enumerate(pc))
    for args in [(32768, page, 2, "blocks", 8192, 4, 512), (4096, page, 2, "blocks", 300, 1, 32),
                 (5 * page, page, 4, "blocks", 70, 290, 16), (1 << 20, page, 2, "split", 8192, 128, 512)]:
        pc = h._placement(*args)
        assert pc[0][0] == 0 and pc[-1][1] == args[0] and all(a[1] == b[0] for a, b in zip(pc, pc[1:]))
        assert all(0 <= k < args[2] and hi > lo and lo % page == 0 for lo, hi, k in pc)

"""Pinned host tensors for the end-to-end path (aie_step_host / aie_step_host_compact), spread over all NUMA nodes.

The e2e step ends in a pure write stream into the caller's host tensors (c2: 296 MB per step).  On a two-socket host
a tensor allocated the usual way lives on the socket of the thread that first touched it, and the stream is then limited
by ONE socket's memory controllers.  `pinned_empty(..., numa="blocks")` binds blocks of env replicas (one transfer slice
each) to alternating nodes with mbind(2) (no libnuma needed) and pins the buffer in place; `numa="split"` makes one block
per node; `interleave=True` sets MPOL_INTERLEAVE for the duration of the allocation so that 4 KB pages alternate.  Falls
back to a plain pinned allocation where the syscalls are unavailable.
"""
import ctypes
import os

_SYS_SET_MEMPOLICY = 238      # x86_64
_MPOL_DEFAULT, _MPOL_INTERLEAVE = 0, 3


def numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return [0]


def _set_mempolicy(mode, nodes):
    libc = ctypes.CDLL(None, use_errno=True)
    if not nodes:
        return libc.syscall(_SYS_SET_MEMPOLICY, ctypes.c_int(mode), None, ctypes.c_ulong(0)) == 0
    mask = ctypes.c_ulong(sum(1 << n for n in nodes))
    return libc.syscall(_SYS_SET_MEMPOLICY, ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(max(nodes) + 2)) == 0


class interleaved:
    """Context manager: allocations first touched inside it are interleaved over all NUMA nodes."""

    def __enter__(self):
        nodes = numa_nodes()
        self.on = len(nodes) > 1 and os.uname().machine == "x86_64" and _set_mempolicy(_MPOL_INTERLEAVE, nodes)
        return self

    def __exit__(self, *exc):
        if self.on:
            _set_mempolicy(_MPOL_DEFAULT, [])
        return False


_SYS_MBIND = 237              # x86_64
_MPOL_BIND = 2
_keep = []                    # mmap objects backing node-split tensors (registered with CUDA; never unmapped)


def transfer_block_rows(n_envs, max_slices=16, item_envs=16):
    """Envs per transfer slice of aie_step_host_compact (csrc/aie_abi.inl: work items of 16 envs, at most 16 slices): the
    block size to pass to pinned_empty(numa="blocks") so that consecutive slices land on alternating NUMA nodes."""
    n_items = (int(n_envs) + item_envs - 1) // item_envs
    n_slices = min(n_items, max_slices)
    return max(1, (n_items + n_slices - 1) // n_slices) * item_envs


def _placement(size, page, n_nodes, numa, rows, row_bytes, block_rows):
    """[(byte lo, byte hi, node index)] covering [0, size) in page-aligned pieces.  "split": one piece per node.  "blocks":
    block b of `block_rows` rows goes to node b % n_nodes; a block boundary is rounded to the nearest page, and a piece takes
    the node of the block that owns its middle byte (tiny tensors: several blocks share a page)."""
    if numa == "split":
        per = (size // n_nodes) // page * page
        cuts = [i * per for i in range(n_nodes)] + [size]
        return [(cuts[i], cuts[i + 1], i) for i in range(n_nodes) if cuts[i + 1] > cuts[i]]
    block_bytes = max(1, block_rows * row_bytes)
    n_blocks = (rows + block_rows - 1) // block_rows
    cuts = sorted({min(size, (b * block_bytes + page // 2) // page * page) for b in range(n_blocks)} | {0, size})
    return [(lo, hi, min(n_blocks - 1, ((lo + hi) // 2) // block_bytes) % n_nodes) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


def pinned_empty(shape, dtype, interleave=False, numa=None, block_rows=None):
    """A pinned host tensor.
    numa="blocks": blocks of `block_rows` rows of the leading axis (env replicas) alternate between the NUMA nodes (bound
    with mbind(2) before first touch, then pinned in place with cudaHostRegister).  With block_rows =
    transfer_block_rows(n_envs) consecutive transfer slices of aie_step_host_compact belong to alternating nodes, so the
    node-pinned expansion threads of every socket have work from the first slice on (the placement the bench uses).
    numa="split": one contiguous block of rows per node.  interleave=True (or numa="interleave"): 4 KB pages alternate
    between the nodes.  Default: plain pinned allocation."""
    import mmap

    import numpy as np
    import torch

    if numa == "interleave":
        interleave = True
    nodes = numa_nodes()
    if numa in ("split", "blocks") and len(nodes) > 1 and os.uname().machine == "x86_64":
        t0 = torch.empty(0, dtype=dtype)
        n_el = int(np.prod(shape))
        nbytes = max(n_el * t0.element_size(), 1)
        page = mmap.PAGESIZE
        size = (nbytes + page - 1) // page * page
        mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        buf = (ctypes.c_char * size).from_buffer(mm)
        base = ctypes.addressof(buf)
        libc = ctypes.CDLL(None, use_errno=True)
        rows = int(shape[0]) if len(shape) else 1
        pieces = _placement(size, page, len(nodes), numa, rows, nbytes // max(rows, 1),
                            int(block_rows or transfer_block_rows(rows)))
        ok = True
        for lo, hi, k in pieces:
            mask = ctypes.c_ulong(1 << nodes[k])
            ok &= libc.syscall(_SYS_MBIND, ctypes.c_void_p(base + lo), ctypes.c_ulong(hi - lo), ctypes.c_int(_MPOL_BIND),
                               ctypes.byref(mask), ctypes.c_ulong(max(nodes) + 2), ctypes.c_uint(0)) == 0
        t = torch.frombuffer(buf, dtype=dtype, count=n_el).reshape(shape)
        t.zero_()                                             # first touch: pages land on their nodes
        rc = torch.cuda.cudart().cudaHostRegister(base, size, 0)
        if int(rc) == 0:     # pinned in place (ok False: some block could not be bound - still valid, just not placed)
            _keep.append((mm, buf))
            return t
        del t                # registering failed: fall through to a plain pinned allocation
    if not interleave:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    with interleaved():
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.zero_()   # first touch under the interleave policy (cudaHostAlloc normally touches the pages itself)
    return t

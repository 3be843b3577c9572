"""Pinned host tensors for the end-to-end path (aie_step_host / aie_step_host_compact), spread over all NUMA nodes.

The e2e step ends in a pure write stream into the caller's host tensors (c2: 296 MB per step).  On a two-socket host
a tensor allocated the usual way lives on the socket of the thread that first touched it, and the stream is then limited
by ONE socket's memory controllers.  `pinned_empty(..., interleave=True)` sets the calling thread's memory policy to
MPOL_INTERLEAVE over all nodes (Linux set_mempolicy(2), no libnuma needed) for the duration of the allocation, so the
pages of the pinned buffer alternate between the sockets; the policy is restored afterwards.  Falls back to a plain pinned
allocation where the syscall is unavailable.
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


def pinned_empty(shape, dtype, interleave=False, numa=None):
    """A pinned host tensor.  numa="split": the leading axis is cut into one contiguous block per NUMA node (block k bound
    to node k with mbind(2) before first touch, then pinned in place with cudaHostRegister) - the placement the expansion
    threads of aie_step_host_compact are matched to: each is pinned to a node and writes the rows that live there.
    interleave=True (or numa="interleave"): 4 KB pages alternate between the nodes.  Default: plain pinned allocation."""
    import mmap

    import numpy as np
    import torch

    if numa == "interleave":
        interleave = True
    nodes = numa_nodes()
    if numa == "split" and len(nodes) > 1 and os.uname().machine == "x86_64":
        t0 = torch.empty(0, dtype=dtype)
        n_el = int(np.prod(shape))
        nbytes = max(n_el * t0.element_size(), 1)
        page = mmap.PAGESIZE
        size = (nbytes + page - 1) // page * page
        mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        buf = (ctypes.c_char * size).from_buffer(mm)
        base = ctypes.addressof(buf)
        libc = ctypes.CDLL(None, use_errno=True)
        per = (size // len(nodes)) // page * page
        ok = True
        for i, node in enumerate(nodes):
            lo = i * per
            hi = size if i == len(nodes) - 1 else (i + 1) * per
            mask = ctypes.c_ulong(1 << node)
            ok &= libc.syscall(_SYS_MBIND, ctypes.c_void_p(base + lo), ctypes.c_ulong(hi - lo), ctypes.c_int(_MPOL_BIND),
                               ctypes.byref(mask), ctypes.c_ulong(max(nodes) + 2), ctypes.c_uint(0)) == 0
        t = torch.frombuffer(buf, dtype=dtype, count=n_el).reshape(shape)
        t.zero_()                                             # first touch: pages land on their nodes
        rc = torch.cuda.cudart().cudaHostRegister(base, size, 0)
        if int(rc) == 0 and ok:
            _keep.append((mm, buf))
            return t
        # fall through to a plain pinned allocation when binding / registering is not possible here
    if not interleave:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    with interleaved():
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.zero_()   # first touch under the interleave policy (cudaHostAlloc normally touches the pages itself)
    return t

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


def pinned_empty(shape, dtype, interleave=True):
    import torch

    if not interleave:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    with interleaved():
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.zero_()   # first touch under the interleave policy (cudaHostAlloc normally touches the pages itself)
    return t

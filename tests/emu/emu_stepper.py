"""TEST INFRASTRUCTURE: drives tests/emu/libaie_emu.so (1-lane host emulation of the device source) through the
same C-ABI and the same Python BatchStepper as the CUDA product, with numpy arrays as the "device" buffers."""
import ctypes as C
import os
import subprocess

import numpy as np

from ai_economist_b200 import _abi
from ai_economist_b200.stepper import BatchStepper

_HERE = os.path.dirname(os.path.abspath(__file__))
_SANITIZE = bool(os.environ.get("AIE_EMU_SANITIZE"))   # ASan + UBSan build (the caller preloads libasan: tools/emu_sanitizers.sh)
_LIB = os.path.join(_HERE, "libaie_emu_asan.so" if _SANITIZE else "libaie_emu.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s", os.path.basename(_LIB)], stderr=subprocess.DEVNULL)
    return _LIB


_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        build()
        _lib = _abi.load_library(_LIB)
    return _lib


class EmuStepper(BatchStepper):
    def __init__(self, spec, n_envs, auto_reset=False, event_envs=0):
        super().__init__(spec, n_envs, emu_lib(), device_index=0, auto_reset=auto_reset, event_envs=event_envs)

    def _alloc(self, shape, dt):
        return np.zeros(shape, dtype=self._DTYPES[dt])

    def _ptr(self, buf):
        return buf.ctypes.data_as(C.c_void_p)

    def to_numpy(self, buf):
        return np.array(buf)

    def state_view(self, name, final=False):
        f = self.field(name)
        shape = [f.shape[i] for i in range(f.ndim)]
        n = int(np.prod(shape)) if shape else 1
        dt = {(1, 0, 0): np.uint8, (1, 0, 1): np.int8, (2, 0, 1): np.int16, (4, 0, 1): np.int32,
              (4, 0, 0): np.uint32, (8, 1, 1): np.float64}[(f.elem_bytes, f.is_float, f.is_signed)]
        state = self.buf["episode_final" if final else "state"]   # a live strided view, like the CUDA stepper's
        inner, s = [], np.dtype(dt).itemsize
        for d in reversed(shape):
            inner.insert(0, s)
            s *= d
        return np.ndarray(shape=[self.n_envs] + shape, dtype=dt, buffer=state, offset=f.offset,
                          strides=[state.strides[0]] + inner)


def emu_factory(spec, n_envs, auto_reset, event_envs=0):
    return EmuStepper(spec, n_envs, auto_reset=auto_reset, event_envs=event_envs)


class EmuCovidStepper:
    """COVID-19 scenario through the emulated device code (numpy buffers)."""

    def __new__(cls, params, n_envs, auto_reset=False, change_list=None):
        from ai_economist_b200.covid_stepper import CovidStepperBase

        class _Emu(CovidStepperBase):
            def _alloc(self, shape, dt):
                return np.zeros(shape, dtype=self._DT[dt])

            def _ptr(self, buf):
                return buf.ctypes.data_as(C.c_void_p)

            def to_numpy(self, buf):
                return np.array(buf)

        return _Emu(params, n_envs, emu_lib(), auto_reset=auto_reset, change_list=change_list)

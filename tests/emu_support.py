"""TEST TOOLING: bind SwiftlyCoreB200 to the host-emulated kernel library."""

import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def emu_core_class():
    if "cls" in _cache:
        return _cache["cls"]
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu  # pylint: disable=import-error,import-outside-toplevel
    import torch  # pylint: disable=import-outside-toplevel

    path = build_emu.build()
    from ska_sdp_distributed_fourier_transform_b200 import _lib, core

    lib = _lib.load(path)
    assert b"EMULATED" in lib.swiftly_b200_build_info()
    lib.swiftly_b200_debug_force_split.argtypes = [ctypes.c_void_p, ctypes.c_int]

    class EmuCore(core.SwiftlyCoreB200):
        """SwiftlyCoreB200 bound to the emulated library; "device" tensors are CPU tensors."""

        tensor_device = torch.device("cpu")

        def __init__(self, W, N, xM_size, yN_size, force_split=False, device=0):
            real_load = _lib.load
            _lib.load = lambda path=None: lib
            try:
                super().__init__(W, N, xM_size, yN_size, device=0)
            finally:
                _lib.load = real_load
            if force_split:
                lib.swiftly_b200_debug_force_split(self._plan, int(force_split))

        def _check_tensor(self, t):
            assert not t.is_cuda

    _cache["cls"] = EmuCore
    return EmuCore

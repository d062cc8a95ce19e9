"""TEST INFRASTRUCTURE ONLY -- run the package's host code against the CPU emulation of liblaplace_b200.

``emulated()`` is a context manager for tests: inside it the package's ctypes binding points at
tests/emu/_build/liblaplace_b200_emu.so (the UNMODIFIED kernel sources compiled against a warp-lockstep CUDA emulator,
see build_emu.py) and the package's "is this a CUDA tensor" gates accept CPU tensors, so the very same Python host
logic + kernel logic that runs on the B200 can be exercised by ``pytest -m "not gpu"``.  Outside the context manager
nothing is changed: the product has no knowledge of the emulator and keeps refusing CPU tensors.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import sys

import torch

from . import build_emu

PKG = "laplace_gnn_recommendation_b200"
_emu_lib = None


def load_emu() -> C.CDLL:
    global _emu_lib
    if _emu_lib is None:
        from laplace_gnn_recommendation_b200 import _lib
        lib = C.CDLL(build_emu.build())
        for name, (res, args) in _lib.PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _emu_lib = lib
    return _emu_lib


class _NullDeviceGuard:
    def __init__(self, *_a, **_k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


@contextlib.contextmanager
def emulated():
    import laplace_gnn_recommendation_b200 as lg   # noqa: F401  (makes sure every submodule is imported)
    from laplace_gnn_recommendation_b200 import _lib, loader

    lib = load_emu()
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    replacements = {
        _lib.require_cuda: lambda *tensors: None,
        _lib.stream: lambda: None,
        _lib.load: lambda: lib,
        _lib.on_device: lambda t: True,
    }
    # the modules did `from ._lib import check, ptr, stream`: rebind those names wherever they were imported
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == PKG or modname.startswith(PKG + ".")):
            continue
        for attr, val in list(vars(mod).items()):
            for orig, repl in replacements.items():
                if val is orig:
                    patch(mod, attr, repl)
    patch(_lib, "_lib", lib)
    patch(loader, "_device_for", lambda t: t.device)
    patch(torch.cuda, "device", _NullDeviceGuard)
    # initcheck: every torch.empty* allocation made while emulating is poisoned (NaN / a large bit pattern), so a kernel or
    # host path that reads memory nobody wrote yields NaNs or wild indices and fails its parity test
    real_empty, real_empty_like = torch.empty, torch.empty_like

    def poison(t):
        if t.numel():
            if t.dtype.is_floating_point:
                t.fill_(float("nan"))
            elif t.dtype in (torch.int32, torch.int64):
                t.fill_(0x7A7A7A7A)
            elif t.dtype == torch.uint8:
                t.fill_(0xAB)
        return t
    from laplace_gnn_recommendation_b200 import csr as _csr
    import time as _time
    patch(_csr, "_time_ms", lambda fn, reps, device: (lambda t0: (fn(), (_time.perf_counter() - t0) * 1e3)[1])(_time.perf_counter()))
    patch(torch, "empty", lambda *a, **k: poison(real_empty(*a, **k)))
    patch(torch, "empty_like", lambda *a, **k: poison(real_empty_like(*a, **k)))
    try:
        yield torch.device("cpu")
    finally:
        for obj, name, value in reversed(saved):
            setattr(obj, name, value)

"""TEST INFRASTRUCTURE: run the product's PYTHON host layer (partgs_b200/*.py: argument marshalling, autograd glue,
buffer management) on CPU tensors against the emulator build of the whole library (tests/cuda_emu).  The product code
is not changed: the three places where it touches the CUDA runtime through torch — the "is this a CUDA tensor" predicate
(`_lib.on_device`), the current-stream lookup and the device context manager — are patched for the duration of a test."""
import contextlib
import ctypes as C
import sys
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).parent / "cuda_emu"))
import build as emu_build  # noqa: E402

from partgs_b200 import _lib  # noqa: E402


class _NullDevice(contextlib.nullcontext):
    def __init__(self, *a, **k):
        super().__init__()


@pytest.fixture()
def emulated_host(monkeypatch):
    try:
        lib = C.CDLL(str(emu_build.build_full()))
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_lib, "_lib", lib)
    monkeypatch.setattr(_lib, "on_device", lambda t: isinstance(t, torch.Tensor))
    monkeypatch.setattr(_lib, "current_stream", lambda device: None)
    monkeypatch.setattr(torch.cuda, "device", _NullDevice)
    return lib

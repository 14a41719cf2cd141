"""Stand-in for libendosurf_b200.so in CPU tests of the HOST logic (autograd plumbing, gradient routing, data-parallel
protocol): every entry point returns 0 at once, so tensors the kernels would have written stay uninitialised.  The flat
gradient buffer of every library backward call is filled with a known pattern instead (``arange % 7``).  Test
infrastructure only - the product has no CPU path (tests/test_abi_cpu.py::test_no_cpu_fallback)."""
import ctypes

import torch


class FakeLib:
    def __getattr__(self, name):
        def fn(*a):
            if name == "es_train_stash_bytes":
                a[2]._obj.value = 64
            return 0
        return fn


def install(setattr_fn=None):
    """setattr_fn: pytest's monkeypatch.setattr (undone after the test) or None for a plain, permanent patch (spawned
    worker processes)."""
    from endosurf_b200 import _lib, renderer as rmod, training
    if setattr_fn is None:
        def setattr_fn(obj, name, value):
            setattr(obj, name, value)
    fake = FakeLib()
    setattr_fn(_lib, "load", lambda: fake)
    setattr_fn(rmod.EndoSurfRenderer, "_context", lambda self: ctypes.c_void_p(1))
    setattr_fn(rmod.EndoSurfRenderer, "_stream", lambda self: ctypes.c_void_p(0))
    orig_init = training._ParamTables.__init__

    def init(self, renderer, params):
        orig_init(self, renderer, params)
        self.flat.copy_(torch.arange(self.flat.numel(), dtype=torch.float32) % 7)
    setattr_fn(training._ParamTables, "__init__", init)


def pattern(n):
    return torch.arange(n, dtype=torch.float32) % 7

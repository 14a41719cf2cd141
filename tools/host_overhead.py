#!/usr/bin/env python
"""Host-side (Python / dispatcher) cost of one trainer iteration on this renderer, measured WITHOUT a GPU.

The library is replaced by a stand-in whose entry points return 0 immediately and the tensors live on the CPU with a
tiny ray batch, so what is timed is exactly the work the host does per iteration: Python in ``endosurf_b200``,
autograd bookkeeping, the dispatcher cost of every small torch op (a CPU op on a 64-element tensor costs about what a
CUDA launch costs the host).  On the GPU this host time is hidden only while the device has more work queued than the
host needs to issue; at the shipped 1024-ray batch it is not (profiles/r2_trainer_loop_shipped_cfg.json).
Profiling tool only: the numbers the stand-in produces are garbage.

    python tools/host_overhead.py [--profile]
"""
import argparse
import cProfile
import ctypes as C
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


class FakeLib:
    calls = 0

    def __getattr__(self, name):
        def fn(*a):
            FakeLib.calls += 1
            if name == "es_train_stash_bytes":
                a[2]._obj.value = 64
            return 0
        return fn


def install_fake():
    from endosurf_b200 import _lib, renderer
    fake = FakeLib()
    _lib.load = lambda: fake
    renderer.EndoSurfRenderer._context = lambda self: C.c_void_p(1)
    renderer.EndoSurfRenderer._stream = lambda self: C.c_void_p(0)
    renderer.EndoSurfRenderer.sync_check = lambda self: None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--rays", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    torch.set_num_threads(1)
    install_fake()
    from trainer_loop_bench import make_trainer, _Recorder
    work = tempfile.mkdtemp(prefix="es_host_")
    tr = make_trainer("fp16x3", args.rays, 8, 8, 4, 48, work)
    tr.writer = _Recorder()
    tr.renderer.train()

    def step(it):
        torch.manual_seed(it)
        np.random.seed(it)
        tr.train_step(global_step=it)
        tr.update_learning_rate(it)

    for it in range(1, 4):
        step(it)
    c0 = FakeLib.calls
    t0 = time.perf_counter()
    for it in range(4, 4 + args.steps):
        step(it)
    dt = (time.perf_counter() - t0) / args.steps
    print(f"host time per trainer iteration: {1e3 * dt:.2f} ms ({(FakeLib.calls - c0) / args.steps:.0f} library calls)")
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        for it in range(100, 100 + args.steps):
            step(it)
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()

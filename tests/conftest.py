import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_cfg():
    with open(os.path.join(GOLD, "base_pull_cfg.json")) as f:
        return json.load(f)


def load_ckpt(device="cpu"):
    g = np.load(os.path.join(GOLD, "ckpt_base_pull.npz"))
    ck = {}
    for k in g.files:
        net, name = k.split("/")
        ck.setdefault(net, {})[name] = torch.from_numpy(g[k]).to(device)
    return ck


def load_npz(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name)).items()}


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def assert_close(name, a, b, tol, kink_tol=None, q=0.995):
    """rel-max error check.  With kink_tol the quantity contains derivatives of the ReLU deformation net, which are
    discontinuous where a pre-activation crosses zero: two correct evaluations may disagree on a few samples, so the
    q-quantile must meet `tol` and the maximum only `kink_tol` (see tests/golden/make_golden.py::check)."""
    a = torch.as_tensor(a).double().cpu().flatten()
    b = torch.as_tensor(b).double().cpu().flatten()
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    assert torch.isfinite(a).all(), f"{name}: non-finite values"
    e = (a - b).abs() / b.abs().max().clamp_min(1e-12)
    mx = e.max().item()
    if kink_tol is None:
        assert mx <= tol, f"{name}: rel-max err {mx:.3e} > {tol:.1e}"
    else:
        qq = torch.quantile(e, q).item() if e.numel() > 1 else mx
        assert qq <= tol and mx <= kink_tol, f"{name}: p{100*q:.1f} {qq:.3e} (tol {tol:.1e}) max {mx:.3e} (tol {kink_tol:.1e})"
    return mx


@pytest.fixture(scope="session")
def cfg():
    return load_cfg()


@pytest.fixture(scope="session")
def ckpt():
    return load_ckpt()

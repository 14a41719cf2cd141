"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the Python mirror has the reference's member surface / parameter layout, and nothing falls back to the CPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_cfg


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "endosurf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(es_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from endosurf_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/endosurf_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTS), "ctypes prototypes and header disagree"


def test_struct_sizes_match_header():
    from endosurf_b200 import _lib
    assert ctypes.sizeof(_lib.EsNetConfig) == 10 * 4
    assert ctypes.sizeof(_lib.EsRenderParams) == 4 * 4 + 4 + 4 + 5 * 8  # 4 ints, float, pad, 5 pointers
    assert ctypes.sizeof(_lib.EsRenderOut) == 11 * 8


def test_renderer_surface_and_param_layout():
    from endosurf_b200 import EndoSurfRenderer
    cfg = load_cfg()
    r = EndoSurfRenderer(cfg["render"], cfg["net"], device="cpu")
    assert isinstance(r, torch.nn.Module)
    for member in ["render_rays", "forward", "get_train_params", "save_checkpoint", "load_checkpoint",
                   "get_cos_anneal_ratio", "n_samples", "n_importance"]:
        assert hasattr(r, member), member
    tp = r.get_train_params()
    assert list(tp.keys()) == ["deform_network", "sdf_network", "color_network", "deviation_network"]
    n = sum(p.numel() for v in tp.values() for p in v)
    assert n == 1654951  # SURVEY.md section 6 (probe of the reference)
    ck = r.save_checkpoint()
    assert list(ck["sdf_network"].keys())[:3] == ["net.0.bias", "net.0.weight_g", "net.0.weight_v"]
    assert ck["sdf_network"]["net.4.weight_v"].shape == (256, 295)
    assert ck["sdf_network"]["net.8.weight_v"].shape == (257, 256)
    assert ck["deform_network"]["net.3.weight_v"].shape == (204, 256)
    assert ck["deform_network"]["net.0.weight_v"].shape == (256, 52)
    assert ck["color_network"]["net.0.weight_v"].shape == (256, 349)
    assert ck["color_network"]["net.4.weight_v"].shape == (256, 605)
    assert list(ck["deviation_network"].keys()) == ["variance"]
    assert r.get_cos_anneal_ratio(25000) == 0.5


def test_checkpoint_roundtrip_with_reference_layout(ckpt):
    from endosurf_b200 import EndoSurfRenderer
    cfg = load_cfg()
    r = EndoSurfRenderer(cfg["render"], cfg["net"], device="cpu")
    r.load_checkpoint(ckpt)  # fixture written by the reference's own save_checkpoint()
    out = r.save_checkpoint()
    for net in ckpt:
        for k in ckpt[net]:
            assert torch.equal(out[net][k], ckpt[net][k])


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly instead of computing somewhere else."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from endosurf_b200 import EndoSurfRenderer
    cfg = load_cfg()
    r = EndoSurfRenderer(cfg["render"], cfg["net"], device="cpu")
    with torch.no_grad(), pytest.raises(RuntimeError):
        r.render_rays(torch.zeros(4, 9))


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = "import sys; import endosurf_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "endosurf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("test oracle", ""), f

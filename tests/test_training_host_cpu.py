"""Host-side logic of the training path that needs no GPU: the parameter tables handed to the C ABI follow the
reference's parameter order, and the plane-record chunk layout helper used by the GPU tests is self-consistent."""
import ctypes as C

import torch

from conftest import load_cfg
from plane_layout import to_chunks, from_chunks, geom_row_of


def test_param_list_matches_module_parameter_order():
    from endosurf_b200 import EndoSurfRenderer
    from endosurf_b200.training import param_list, net_ids
    cfg = load_cfg()
    r = EndoSurfRenderer(cfg["render"], cfg["net"], device="cpu")
    pl = param_list(r)
    mp = [p for n, p in r.model.named_parameters() if "variance" not in n]
    assert len(pl) == len(mp) == 81 and all(a is b for a, b in zip(pl, mp))
    names = [n for n, _ in r.model.named_parameters()][:3]
    assert names == ["deform_network.net.0.bias", "deform_network.net.0.weight_g", "deform_network.net.0.weight_v"]
    assert net_ids(r) == [0, 1, 2]


def test_train_param_struct_layout():
    from endosurf_b200 import _lib
    assert C.sizeof(_lib.EsTrainParams) == 5 * 3 * 8
    assert C.sizeof(_lib.EsRenderGrads) == 8 * 8
    assert C.sizeof(_lib.EsProfile) == 12 * 8 * 3


def test_chunk_layout_roundtrip():
    g = torch.Generator().manual_seed(0)
    m = torch.randn(3 * 128, 256, generator=g).half()
    rec = to_chunks(m)
    assert rec.shape == (3, 4, 16384)
    assert torch.equal(from_chunks(rec), m)
    # element (tile row r, column c) of a chunk sits at k-group c/8, row r, element c%8
    v = rec[2, 1].view(torch.float16)
    assert v[(5 * 128 + 77) * 8 + 3] == m[2 * 128 + 77, 64 + 5 * 8 + 3]
    # stream s of point (tile, Q, p) is tile row 32 Q + 8 s + p
    assert geom_row_of(torch.tensor([0, 9, 33, 63]), 2).tolist() == [16, 32 + 16 + 1, 128 + 16 + 1, 128 + 96 + 16 + 7]

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


def test_fast_param_traversal_and_gradient_tables():
    """The renderer's dict-level parameter traversal equals nn.Module.parameters() (with and without a deformation
    network, and after a parameter object has been replaced), and the cached es_train_params tables point at the
    per-parameter views of the flat gradient buffer."""
    import copy
    from endosurf_b200 import EndoSurfRenderer, training
    cfg = load_cfg()
    for use_deform in (True, False):
        nc = copy.deepcopy(cfg["net"])
        nc["use_deform"] = use_deform
        r = EndoSurfRenderer(cfg["render"], nc, device="cpu")
        fp, mp = r._fast_params(), list(r.model.parameters())
        assert len(fp) == len(mp) == (82 if use_deform else 55) and all(a is b for a, b in zip(fp, mp))
        v0 = r._params_version()
        lyr = r.model.sdf_network.net[3]
        lyr.bias = torch.nn.Parameter(lyr.bias.detach().clone())  # a replaced parameter object is seen
        assert all(a is b for a, b in zip(r._fast_params(), r.model.parameters()))
        assert r._params_version() != v0
        v1 = r._params_version()
        with torch.no_grad():
            lyr.weight_v.add_(1.0)  # and so is an in-place update (what the optimizer does)
        assert r._params_version() != v1
        pl = tuple(training.param_list(r))
        L = r._cfg_struct.n_layers
        for _ in range(2):  # second pass: the static tables come from the cache, the gradient buffer is new
            t = training._ParamTables(r, pl)
            g = t.grads
            assert t.flat.numel() == sum(p.numel() for p in pl) and float(t.flat.abs().sum()) == 0.0
            k = 0
            for net in training.net_ids(r):
                for l in range(L):
                    assert t.struct.grad_b[net][l] == g[k].data_ptr() and t.struct.grad_g[net][l] == g[k + 1].data_ptr()
                    assert t.struct.grad_v[net][l] == g[k + 2].data_ptr()
                    assert t.struct.v[net][l] == pl[k + 2].data_ptr() and t.struct.g[net][l] == pl[k + 1].data_ptr()
                    for j in range(3):
                        assert g[k + j].shape == pl[k + j].shape and g[k + j].is_contiguous()
                    k += 3
            assert k == len(pl)
            if not use_deform:
                assert not t.struct.v[0] and not t.struct.grad_v[0]
            for i, x in enumerate(g):
                x.fill_(i + 1.0)
            assert bool((t.flat != 0).all())  # the views tile the buffer

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


def test_param_hub_routes_flat_gradients(monkeypatch):
    """Autograd plumbing of the training path without a GPU: the library is replaced by a stand-in (every entry point
    returns 0) and the backward's flat gradient buffer by a known pattern.  Two library backward calls in one graph
    (render_rays + a point-field loss, as the reference's train_step has) must deliver pattern x 2 to every parameter
    through ONE flat addition + the hub node; a bound gradient sink must end up with the same numbers without the hub
    node running; the hub handle is reused until a parameter's requires_grad flag changes."""
    import collections
    import copy
    from torch.utils._python_dispatch import TorchDispatchMode
    import fake_lib
    from endosurf_b200 import EndoSurfRenderer, training, distributed as dp
    fake_lib.install(monkeypatch.setattr)

    cfg = load_cfg()
    rc = copy.deepcopy(cfg["render"])
    rc.update(n_samples=8, n_importance=8)
    r = EndoSurfRenderer(rc, cfg["net"], device="cpu")
    r.train()
    rays = torch.rand(16, 9)
    x, d, t = torch.rand(10, 3) - 0.5, torch.nn.functional.normalize(torch.randn(10, 3), dim=-1), torch.rand(10, 1)
    params = [p for v in r.get_train_params().values() for p in v]
    n_net = sum(p.numel() for p in params[:-1])
    pat = fake_lib.pattern(n_net)

    class Count(TorchDispatchMode):
        def __init__(self):
            super().__init__()
            self.c = collections.Counter()

        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            self.c[func.__name__] += 1
            return func(*args, **(kwargs or {}))

    def backward_twice():
        o = r.render_rays(rays, iter_step=1000)
        sdf = r.point_field(x, d, t)[0]
        with Count() as c:
            torch.autograd.backward([o["color_map"], sdf], [torch.ones_like(o["color_map"]), torch.ones_like(sdf)])
        return c.c

    # (1) plain autograd: p.grad of every network parameter = 2 x pattern, with one flat add instead of 81 x 2
    ops = backward_twice()
    assert torch.equal(torch.cat([p.grad.reshape(-1) for p in params[:-1]]), 2 * pat)
    assert all(p.grad.shape == p.shape for p in params)
    assert ops["add.Tensor"] == 1 and ops["as_strided.default"] == 81, dict(ops)
    # (2) flat bucket, unbound and bound: identical numbers; bound = 2 flat additions and no per-parameter work
    bucket = dp.FlatGradBucket(params)
    res = {}
    for bound in (False, True):
        r._grad_sink = bucket if bound else None
        bucket.zero()
        ops = backward_twice()
        res[bound] = bucket.flat[:n_net].clone()
        if bound:
            assert ops["as_strided.default"] == 0 and ops["add_.Tensor"] <= 4, dict(ops)
    r._grad_sink = None
    assert torch.equal(res[False], res[True]) and torch.equal(res[True], 2 * pat)
    # (3) the hub handle lives across calls; a requires_grad change or no-grad mode gives a new one
    h1 = training.param_hub(r)[1]
    assert training.param_hub(r)[1] is h1
    params[3].requires_grad_(False)
    assert training.param_hub(r)[1] is not h1
    params[3].requires_grad_(True)
    with torch.no_grad():
        assert not training.param_hub(r)[1].requires_grad
    assert training.param_hub(r)[1].requires_grad
    (g,) = torch.autograd.grad(r.point_field(x, d, t)[0].sum(), [params[2]])
    assert g.shape == params[2].shape

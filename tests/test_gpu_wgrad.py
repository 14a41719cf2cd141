"""GPU unit tests of the training kernels' building blocks: the split-K tcgen05 weight-gradient kernel on
caller-built plane records (MN-major operand descriptors, bias sums, split-K reduce) and the plane records the
forward training chains write with bulk shared->global copies."""
import ctypes as C
import copy

import pytest
import torch

from conftest import load_npz
from plane_layout import to_chunks, from_chunks, geom_row_of

pytestmark = pytest.mark.gpu


def _ctx_renderer(cfg, ckpt, use_deform=True):
    from endosurf_b200 import EndoSurfRenderer
    rc = copy.deepcopy(cfg["render"])
    nc = copy.deepcopy(cfg["net"])
    nc["use_deform"] = use_deform
    r = EndoSurfRenderer(rc, nc, device="cuda")
    r.load_checkpoint({k: v for k, v in ckpt.items() if use_deform or k != "deform_network"})
    return r


def _probe(r, zbar, a, bias_mode):
    from endosurf_b200 import _lib
    lib, ctx = _lib.load(), r._context()
    rows, n_in = a.shape
    zrec, arec = to_chunks(zbar).cuda(), to_chunks(a).cuda()
    out = torch.full((256, n_in), float("nan"), device="cuda")
    bias = torch.full((256,), float("nan"), device="cuda")
    rc = lib.es_wgrad_probe(ctx, C.c_void_p(zrec.data_ptr()), C.c_void_p(arec.data_ptr()), rows // 128, n_in // 64,
                            bias_mode, C.c_void_p(out.data_ptr()), C.c_void_p(bias.data_ptr()), r._stream())
    _lib.check(ctx, rc, "es_wgrad_probe")
    r.sync_check()
    return out.cpu(), bias.cpu()


@pytest.mark.parametrize("paired", [True, False])
@pytest.mark.parametrize("n_b,tiles,bias_mode", [(4, 7, 1), (1, 3, 2), (2, 1, 0), (3, 10, 2)])
def test_wgrad_kernel_matches_matmul(cfg, ckpt, n_b, tiles, bias_mode, paired):
    """out = zbar^T a over tiles*128 rows; fp16 inputs, fp32 accumulation: exact products, so the only error is the
    summation order."""
    from endosurf_b200 import _lib
    r = _ctx_renderer(cfg, ckpt)
    # paired: the two M halves run as a 2-CTA cluster sharing the B tile by multicast; unpaired: independent CTAs
    _lib.load().es_debug_set(r._context(), 5, int(paired))
    g = torch.Generator().manual_seed(7 + n_b)
    rows = tiles * 128
    zbar = (torch.randn(rows, 256, generator=g) * torch.rand(rows, 1, generator=g)).half()
    a = torch.randn(rows, 64 * n_b, generator=g).half()
    ref = zbar.double().t() @ a.double()
    out, bias = _probe(r, zbar, a, bias_mode)
    err = ((out.double() - ref).norm() / ref.norm()).item()
    if err > 1e-5:
        # tell a descriptor-convention problem from anything else: try the swapped (LBO, SBO) reading
        lib, ctx = _lib.load(), r._context()
        lib.es_debug_set(ctx, 1, 2048)
        lib.es_debug_set(ctx, 2, 128)
        out2, _ = _probe(r, zbar, a, bias_mode)
        err2 = ((out2.double() - ref).norm() / ref.norm()).item()
        pytest.fail(f"wgrad kernel rel err {err:.3e} with (LBO,SBO)=(128,2048); swapped strides give {err2:.3e}")
    if bias_mode == 1:
        bref = zbar.double().sum(0)
    elif bias_mode == 2:
        prim = (torch.arange(rows) % 32) < 8
        bref = zbar.double()[prim].sum(0)
    if bias_mode:
        berr = ((bias.double() - bref).norm() / bref.norm()).item()
        assert berr < 1e-5, f"bias sums rel err {berr:.3e}"


def test_forward_plane_records(cfg, ckpt):
    """The geometry training chain's record: the encoder chunk of the first deform layer (primal rows = enc6(x), enc6(t)
    in the kernel's column order, tangent rows = their x-derivatives) and the round trip of a hidden layer chunk
    through the stash to the reverse chain are what the layout says."""
    from endosurf_b200 import _lib
    from endosurf_b200.training import _stash, _ptr
    from oracle import endosurf_oracle as orc
    r = _ctx_renderer(cfg, ckpt)
    r._sync_weights()
    lib, ctx = _lib.load(), r._context()
    s = load_npz("stage_points.npz")
    n = 180  # not a multiple of 32: the last tile is padded
    x, d, t = (torch.from_numpy(s[k][:n]).cuda() for k in "xdt")
    stash = _stash(r, n, "cuda")
    stash.zero_()
    o = [torch.empty(n, k, device="cuda") for k in (3, 9, 1, 3, 3)]
    rc = lib.es_point_train_forward(ctx, _ptr(x), _ptr(t.reshape(-1)), 1, 1, _ptr(d), 1, 3, n, *[_ptr(v) for v in o],
                                    _ptr(stash), r._stream())
    _lib.check(ctx, rc, "es_point_train_forward")
    r.sync_check()
    tiles = (n + 31) // 32
    # record chunk 0 of every geometry tile = SRC_ENC_DEFORM chunk (es_api.cu assign_forward)
    cm = (C.c_int32 * 64)()
    lib.es_chunk_colmap(ctx, 0, 1, cm)
    cm = torch.tensor(list(cm))
    g_chunks, c_chunks = 66, 38  # chunks per tile of the geometry / colour records (es_api.cu assign_forward)
    assert stash.numel() == tiles * g_chunks * 16384 + ((n + 127) // 128) * c_chunks * 16384 + 256
    rec = stash[:tiles * g_chunks * 16384].reshape(tiles, g_chunks, 16384)
    enc = from_chunks(rec[:, :1].cpu()).float()  # [tiles*128, 64]
    ref = torch.cat([orc.freq_encode(x.cpu(), 6), orc.freq_encode(t.cpu().reshape(-1, 1), 6)], -1)  # [n, 52]
    pts = torch.arange(n)
    prim = enc[geom_row_of(pts, 0)]
    for col in range(64):
        if cm[col] >= 0:
            e = (prim[:, col] - ref[:, cm[col]]).abs().max().item()
            assert e < 2e-3 * max(1.0, ref[:, cm[col]].abs().max().item()), (col, e)
        else:
            assert prim[:, col].abs().max().item() == 0.0
    # tangent row j: derivative of the position features wrt x_j, zero for the time features
    xr = x.cpu().clone().requires_grad_(True)
    er = orc.freq_encode(xr, 6)
    for j in range(3):
        tan = enc[geom_row_of(pts, 1 + j)]
        for col in range(0, 64, 3):
            if 0 <= cm[col] < 39:
                (gcol,) = torch.autograd.grad(er[:, cm[col]].sum(), xr, retain_graph=True)
                e = (tan[:, col] - gcol[:, j]).abs().max().item()
                assert e < 2e-3 * max(1.0, gcol.abs().max().item()), (j, col, e)
            elif cm[col] >= 39:
                assert tan[:, col].abs().max().item() == 0.0

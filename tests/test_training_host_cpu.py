"""Host-side logic of the training path that needs no GPU: the geometry planes' tile row order, the vectorised
positional encoding (reference encoder.py:40-54), and the sync-free cumprod used by the differentiable compositing."""
import torch

from endosurf_b200 import training as T


def test_rows_points_roundtrip_and_layout():
    pg, k = 64, 5  # two 32-point tiles
    v = torch.arange(pg * 4 * k, dtype=torch.float32).reshape(pg, 4, k)
    rows = T.rows_from_points(v)
    assert rows.shape == (pg * 4, k)
    assert torch.equal(T.points_from_rows(rows), v)
    # stream s of point (tile, Q, p) is tile row 32 Q + 8 s + p (csrc/es_mlp.cu, tangent mode)
    for tile, Q, p, s in [(0, 0, 0, 0), (0, 1, 3, 2), (1, 3, 7, 3), (1, 2, 5, 1)]:
        point = tile * 32 + 8 * Q + p
        assert torch.equal(rows[tile * 128 + 32 * Q + 8 * s + p], v[point, s])
    for s in range(4):
        assert torch.equal(T.stream_rows(rows, s), v[:, s])


def _freq_enc_loop(x, n_freqs):
    out = [x]
    for kk in range(n_freqs):
        f = float(2.0 ** kk)
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


def test_freq_enc_matches_per_frequency_loop_bitwise():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(257, 3, generator=g) * 1.3
    t = torch.rand(257, 1, generator=g)
    for n_freqs in (4, 6, 10):
        assert torch.equal(T.freq_enc(x, n_freqs), _freq_enc_loop(x, n_freqs))
    assert torch.equal(T.freq_enc(t, 6), _freq_enc_loop(t, 6))
    # tangent rows = d enc / d x_j, checked against autograd of the loop form
    xr = x[:16].clone().requires_grad_(True)
    e = _freq_enc_loop(xr, 6)
    tan = T.freq_enc_tangent(xr.detach(), 6)
    for col in range(0, e.shape[1], 5):
        (gcol,) = torch.autograd.grad(e[:, col].sum(), xr, retain_graph=True)
        assert torch.allclose(tan[:, :, col], gcol, rtol=0, atol=1e-6)


def test_cumprod_pos_backward_matches_autograd():
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(32, 40, generator=g) * 0.999 + 1e-7).double()
    w = torch.randn(32, 40, generator=g).double()
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya = T._CumprodPos.apply(a)
    yb = torch.cumprod(b, -1)
    assert torch.equal(ya.detach(), yb.detach())
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-10, atol=1e-14)


def test_split16_carries_22_bits_above_the_subnormal_floor():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4096, generator=g) * 10
    hi, lo = T.split16(x)
    # 22 mantissa bits relative, down to the absolute floor of fp16 subnormals (half of 2^-24) for the lo half
    err = (hi.double() + lo.double() - x.double()).abs()
    assert (err <= 2.0 ** -21 * x.double().abs() + 2.0 ** -25).all()

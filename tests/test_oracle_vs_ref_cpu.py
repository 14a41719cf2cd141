"""The oracle's sampling arithmetic against the UNMODIFIED reference executed directly (byte-compiled under
oracle/_ref by oracle/build_ref.py), on seeded random inputs and on the edge cases the fixtures do not hold: rays that
miss the unit sphere or start inside it, degenerate (all-zero, one-hot) importance weights, two-sample rays, repeated
sample positions, flat and sign-changing SDF profiles.  Skipped where oracle/_ref has not been built."""
import copy
import importlib

import pytest
import torch

from conftest import load_cfg, load_ckpt
from oracle import endosurf_oracle as orc
from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="oracle/_ref (byte-compiled reference) not built")


@pytest.fixture(scope="module")
def ref():
    mod = ref_shims.load_reference()
    utils = importlib.import_module("src.renderer.utils")
    cfg, ckpt = load_cfg(), load_ckpt()
    rc = copy.deepcopy(cfg["render"])
    rc.update(n_samples=16, n_importance=16, perturb=False)
    torch.manual_seed(0)
    r = mod.EndoSurfRenderer(rc, cfg["net"], device="cpu")
    r.load_checkpoint(ckpt)
    return r, utils, orc.OracleNet(ckpt, cfg["net"])


def _rays(n, seed):
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, -1.5]) + 0.05 * torch.randn(n, 3, generator=g)
    d = torch.nn.functional.normalize(torch.cat([0.3 * torch.randn(n, 2, generator=g), torch.ones(n, 1)], -1), dim=-1)
    return o, d


def test_sphere_intersection_edge_cases(ref):
    _, utils, _ = ref
    o, d = _rays(64, 1)
    o[:8] = torch.tensor([3.0, 0.0, -1.5])            # misses the sphere
    o[8:16] = 0.1 * torch.randn(8, 3)                 # starts inside: near clamps to 0
    o[16:24] = torch.tensor([0.0, 0.0, 1.5])          # sphere behind the camera
    d[24:32] = 2.5 * d[24:32]                         # non-unit directions
    got = orc.sphere_intersection(o, d)
    want = utils.get_sphere_intersection(o, d)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert not bool(want[2][:8].any()) and bool((want[0][8:16] == 0).all())


@pytest.mark.parametrize("n_bins,n_new", [(2, 4), (16, 8), (64, 16), (112, 16)])
def test_sample_pdf_random_and_degenerate_weights(ref, n_bins, n_new):
    _, utils, _ = ref
    g = torch.Generator().manual_seed(n_bins)
    bins = torch.sort(torch.rand(12, n_bins, generator=g) * 2.0, dim=-1)[0]
    w = torch.rand(12, n_bins - 1, generator=g)
    w[0] = 0.0                                        # all-zero weights: uniform after the +1e-5
    w[1] = 0.0
    w[1, (n_bins - 1) // 2] = 1.0                     # one-hot: every other bin hits the denom < 1e-5 branch
    w[2] = 1e-7 * w[2]                                # weights far below the 1e-5 floor
    bins[3, 1:] = bins[3, :1]                         # every bin edge at the same position
    if n_bins > 3:
        bins[4, 2] = bins[4, 1]                       # one empty bin
    got = orc.sample_pdf_det(bins, w, n_new)
    want = utils.sample_pdf(bins, w, n_new, det=True)
    assert torch.equal(got, want)


@pytest.mark.parametrize("seed,n,inv_s", [(0, 16, 64.0), (1, 24, 128.0), (2, 32, 256.0), (3, 40, 512.0), (4, 2, 64.0)])
def test_up_sample_profiles(ref, seed, n, inv_s):
    r, _, _ = ref
    g = torch.Generator().manual_seed(100 + seed)
    R = 10
    o, d = _rays(R, seed)
    near, far, _ = orc.sphere_intersection(o, d)
    z = near + (far - near) * torch.sort(torch.rand(R, n, generator=g), dim=-1)[0]
    sdf = 0.4 * torch.randn(R, n, generator=g)
    sdf[0] = torch.linspace(0.5, -0.5, n)             # one clean crossing
    sdf[1] = 0.3                                      # no surface: flat positive
    sdf[2] = -0.2                                     # always inside
    sdf[3] = 0.2 * torch.cos(torch.linspace(0.0, 12.0, n))  # several crossings
    if n > 3:
        z[4, 2] = z[4, 1]                             # repeated sample position: the +1e-6 in cos_val matters
    o[5] = torch.tensor([3.0, 0.0, -1.5])             # samples outside the unit sphere: inside_sphere masks cos_val
    got = orc.up_sample(o, d, z, sdf, 8, inv_s)
    want = r.up_sample(o, d, z, sdf, 8, inv_s)
    assert got.shape == want.shape == (R, 8)
    assert torch.allclose(got, want, rtol=0.0, atol=2e-7), (got - want).abs().max().item()


def test_cat_z_vals_with_and_without_sdf(ref):
    r, _, net = ref
    g = torch.Generator().manual_seed(9)
    R, n, k = 6, 16, 4
    o, d = _rays(R, 9)
    time = torch.rand(R, generator=g)
    near, far, _ = orc.sphere_intersection(o, d)
    z = near + (far - near) * torch.sort(torch.rand(R, n, generator=g), dim=-1)[0]
    new_z = near + (far - near) * torch.sort(torch.rand(R, k, generator=g), dim=-1)[0]
    new_z[0, 0] = z[0, 3]                             # a new sample exactly on an old one (sort tie)
    sdf = torch.randn(R, n, generator=g)
    with torch.no_grad():
        z1, s1 = orc.cat_z_vals(net, o, d, time, z, new_z, sdf, last=False)
        z2, s2 = r.cat_z_vals(o, d, time, z, new_z, sdf, last=False)
        z3, _ = orc.cat_z_vals(net, o, d, time, z, new_z, sdf, last=True)
        z4, _ = r.cat_z_vals(o, d, time, z, new_z, sdf, last=True)
    assert torch.equal(z1, z2) and torch.equal(z3, z4) and torch.equal(z1, z3)
    assert torch.allclose(s1, s2, rtol=0.0, atol=2e-6), (s1 - s2).abs().max().item()
    assert bool((z1[:, 1:] >= z1[:, :-1]).all())


def test_freq_encoder_matches_reference(ref):
    r, _, _ = ref
    enc = r.model.sdf_network.enc_fn_pos if hasattr(r.model.sdf_network, "enc_fn_pos") else None
    if enc is None:
        pytest.skip("reference encoder attribute not found")
    x = torch.cat([torch.zeros(1, 3), torch.ones(1, 3), -torch.ones(1, 3), torch.rand(13, 3) * 2 - 1], 0)
    want = enc(x, bound=1.0)
    got = orc.freq_encode(x, 6)
    assert got.shape == want.shape and torch.allclose(got, want, rtol=0.0, atol=1e-6)


@pytest.mark.parametrize("iter_step", [0, 30000])
def test_render_rays_with_the_reference_rng_jitter(ref, iter_step):
    """perturb=True: the reference draws ONE torch.rand([R,1]) per call (endosurf.py:81); with the same generator state
    the oracle, fed that draw as t_rand, must reproduce the whole path (jittered coarse samples, 4 up-sampling steps,
    render_core) - and the RNG-free entries it is usually compared on must differ from the unjittered render."""
    r, _, net = ref
    cfg = load_cfg()
    rc = copy.deepcopy(cfg["render"])
    rc.update(n_samples=16, n_importance=16, perturb=True)
    rays = orc.synthetic_rays(12, frame=11, seed=4)
    r.train()
    torch.manual_seed(1234)
    want = r.render_rays(rays, iter_step=iter_step, perturb_overwrite=True)
    torch.manual_seed(1234)
    t_rand = torch.rand([rays.shape[0], 1]) - 0.5
    got = orc.render_rays(net, rc, rays, iter_step=iter_step, t_rand=t_rand)
    plain = orc.render_rays(net, rc, rays, iter_step=iter_step, perturb_overwrite=False)
    for k in ["color_map", "depth_map", "weights", "cdf", "gradients_o", "gradient_o_error", "weight_max", "s_val"]:
        a, b = got[k].detach(), want[k].detach()
        err = ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
        # maps: the oracle's pin level (tests/golden/ORACLE_PIN.txt, <= 3e-5); per-sample NeuS weights / cdf amplify
        # the fp32 rounding differences of the two evaluation orders (alpha = 1 - cdf ratio of nearly equal sigmoids)
        tol = 3e-5 if k in ("color_map", "depth_map", "gradient_o_error", "s_val") else 2e-4
        assert err < tol, (k, err)
    assert (got["depth_map"] - plain["depth_map"]).abs().max().item() > 1e-5

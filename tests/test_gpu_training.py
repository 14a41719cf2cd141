"""GPU parity of the differentiable (training) path: parameter gradients of the fused forward + reverse chains against
the oracle's torch.autograd gradients (the reference differentiates the same graph, endosurf.py:594-658)."""
import copy
import os

import pytest
import torch

from conftest import load_npz, assert_close, rel_err

pytestmark = pytest.mark.gpu


def _renderer(cfg, ckpt, ns, ni, use_deform=True, full_planes=False, precision_terms=3):
    from endosurf_b200 import EndoSurfRenderer, _lib
    rc = copy.deepcopy(cfg["render"])
    rc.update(n_samples=ns, n_importance=ni, perturb=False)
    nc = copy.deepcopy(cfg["net"])
    nc["use_deform"] = use_deform
    r = EndoSurfRenderer(rc, nc, device="cuda", precision_terms=precision_terms)
    r.load_checkpoint({k: v for k, v in ckpt.items() if use_deform or k != "deform_network"})
    r.train()
    _lib.check(r._context(), _lib.load().es_set_plane_mode(r._context(), int(full_planes)), "es_set_plane_mode")
    return r, rc, nc


def _oracle_grads(ckpt, nc, fn):
    from oracle import endosurf_oracle as orc
    ck = {n: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for n, sd in ckpt.items()
          if nc["use_deform"] or n != "deform_network"}
    net = orc.OracleNet(ck, nc)
    loss = fn(orc, net)
    loss.backward()
    return loss.detach(), {f"{n}.{k}": v.grad for n, sd in ck.items() for k, v in sd.items()}


def _my_grads(r, loss):
    r.zero_grad()
    loss.backward()
    out = {}
    for name, p in r.model.named_parameters():
        out[name] = None if p.grad is None else p.grad.detach().cpu()
    return out


def _compare(mine, ref, tol_norm=1e-2, tol_global=1e-3):
    """every parameter tensor: ||g - g_ref|| / ||g_ref|| <= tol_norm (Adam normalises per tensor), and the whole
    gradient vector within tol_global.  On tiny batches the per-tensor bound has to cover single ReLU-gate flips of
    the deformation net (conftest.assert_close); the batch-size tests below hold the timed path to 1e-3 / 1e-4."""
    worst = []
    for k in ref:
        if mine.get(k) is not None:
            assert torch.isfinite(mine[k]).all(), f"non-finite gradient for {k}"
    num = sum(((mine[k] - g) ** 2).sum().item() for k, g in ref.items() if g is not None and mine[k] is not None)
    den = sum((g ** 2).sum().item() for g in ref.values() if g is not None)
    assert (num / max(den, 1e-30)) ** 0.5 <= tol_global, f"global gradient error {(num / den) ** 0.5:.3e}"
    for k, g_ref in ref.items():
        g = mine[k]
        if g_ref is None:
            continue
        assert g is not None, f"no gradient for {k}"
        den = max(g_ref.norm().item(), 1e-12)
        e = (g - g_ref).norm().item() / den
        worst.append((e, k, den))
    worst.sort(reverse=True)
    bad = [(e, k, d) for e, k, d in worst if e > tol_norm]
    assert not bad, f"gradient mismatch (rel-norm err, param, ref norm): {bad[:8]}"
    return worst[:5]


@pytest.mark.parametrize("use_deform,full_planes", [(True, False), (False, False), (True, True)])
def test_point_field_gradients(cfg, ckpt, use_deform, full_planes):
    """Random adjoints on every output of the point pipeline (sdf, g_o, rgb) at explicit points."""
    s = load_npz("stage_points.npz")
    n = 96
    x, d, t = (torch.from_numpy(s[k][:n]) for k in "xdt")
    g = torch.Generator().manual_seed(3)
    a_sdf, a_go, a_rgb = torch.randn(n, 1, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    r, rc, nc = _renderer(cfg, ckpt, 32, 32, use_deform, full_planes)

    def oracle_loss(orc, net):
        raw = net.forward(torch.cat([x, d, t], -1))
        g_o = net.sdf_grad_observed(x.clone(), t)
        return (raw[:, :1] * a_sdf).sum() + (raw[:, 1:4] * a_rgb).sum() + (g_o * a_go).sum()

    ref_loss, ref = _oracle_grads(ckpt, nc, oracle_loss)
    sdf, g_c, jac, rgb = r.point_field(x.cuda(), d.cuda(), t.cuda())
    g_o = torch.einsum("nij,ni->nj", jac, g_c)
    loss = (sdf * a_sdf.cuda()).sum() + (rgb * a_rgb.cuda()).sum() + (g_o * a_go.cuda()).sum()
    assert rel_err(loss.detach(), ref_loss) < 1e-4
    mine = _my_grads(r, loss)
    r.sync_check()
    _compare(mine, ref)


@pytest.mark.parametrize("tag,use_deform,full_planes", [("r32_s32_i32_it25k", True, False),
                                                        ("r32_nodeform_s32_i32", False, False),
                                                        ("r32_s32_i32_it25k", True, True)])
def test_render_rays_training_gradients(cfg, ckpt, tag, use_deform, full_planes):
    """d loss / d every parameter for loss = 0.7 sum(color) + 0.3 sum(depth) + 0.1 eikonal on fixed z_vals."""
    g = load_npz(f"render_{tag}.npz")
    ns, ni = int(g["n_samples"]), int(g["n_importance"])
    r, rc, nc = _renderer(cfg, ckpt, ns, ni, use_deform, full_planes)
    rays = torch.from_numpy(g["rays"])
    z = torch.from_numpy(g["z_vals"])
    it = int(g["iter_step"])

    def lossf(o):
        return o["color_map"].sum() * 0.7 + o["depth_map"].sum() * 0.3 + o["gradient_o_error"] * 0.1

    ref_loss, ref = _oracle_grads(ckpt, nc, lambda orc, net: lossf(
        orc.render_rays(net, rc, rays, iter_step=it, perturb_overwrite=False, z_vals_override=z)))
    o = r.render_rays(rays.cuda(), iter_step=it, perturb_overwrite=False, z_vals_override=z.cuda())
    for k in ["color_map", "depth_map", "weights", "cdf"]:
        assert_close(k, o[k].detach(), g["core/" + k], 1e-4, kink_tol=2e-2)
    loss = lossf(o)
    assert rel_err(loss.detach(), ref_loss) < 1e-4
    mine = _my_grads(r, loss)
    r.sync_check()
    _compare(mine, ref)
    # the golden file holds a few of the reference's own gradients (different z: its own resampling), sanity only
    for key in ["model.deviation_network.variance"]:
        if "grad/" + key in g:
            assert rel_err(mine[key.replace("model.", "")], g["grad/" + key]) < 5e-2


def test_training_step_changes_loss(cfg, ckpt):
    """A few Adam steps through the fused forward/backward reduce a colour-fitting loss (end-to-end training smoke)."""
    from oracle import endosurf_oracle as orc
    r, rc, nc = _renderer(cfg, ckpt, 16, 16)
    rays = orc.synthetic_rays(64, frame=2, seed=4).cuda()
    target = torch.full((64, 3), 0.25, device="cuda")
    opt = torch.optim.Adam([p for v in r.get_train_params().values() for p in v], lr=5e-4)
    losses = []
    for it in range(6):
        opt.zero_grad()
        o = r(rays, iter_step=1000)
        loss = (o["color_map"] - target).abs().mean() + 0.1 * o["gradient_o_error"]
        loss.backward()
        opt.step()
        losses.append(loss.item())
    r.sync_check()
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]


def _trainer_loss(o, color_gt, depth_gt, mask):
    """The colour / depth / eikonal terms of the reference trainer with its mean normalisation
    (trainer_endosurf.py:131-152; weights of configs/endosurf/baseline/base_pull.yml), which makes the adjoints that
    enter the backward 1e-5 .. 1e-8."""
    import torch.nn.functional as F
    ce = (o["color_map"] - color_gt) * mask
    color_loss = F.l1_loss(ce, torch.zeros_like(ce), reduction="sum") / (mask.sum() + 1e-10)
    de = (o["depth_map"] - depth_gt) * mask
    depth_loss = F.l1_loss(de, torch.zeros_like(de), reduction="sum") / (mask.sum() + 1e-10)
    return color_loss * 1.0 + depth_loss * 1.0 + o["gradient_o_error"] * 0.1


def _rel(a, b):
    return (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)


@pytest.mark.parametrize("it", [0, 50000])
def test_timed_path_gradient_parity_512_rays(cfg, ckpt, it):
    """Gradient parity of exactly what bench.py times: default plane mode (fp16 hi planes, 1-term tcgen05 weight
    gradients, hi-only activation gates), 512 rays x (64+64) samples = 65,536 points (262,144 geometry rows), the
    trainer's mean-normalised losses, fixed z_vals, all 82 parameter tensors.

    The yardstick is the oracle evaluated in float64.  The reference's own fp32 autograd deviates from it by up to
    1.1e-3 on the first colour layers (heavy cancellation in those sums; tools/diag_grad_precision.py), so the bound is
        whole gradient vector   <= 1e-4   relative l2
        every parameter tensor  <= max(1e-3, 5 x the fp32 reference's own deviation for that tensor):
    the fp16 hi/lo products carry 22 mantissa bits against fp32's 24, so an ill-conditioned sum amplifies our rounding
    4x more than the reference's."""
    from oracle import endosurf_oracle as orc
    r, rc, nc = _renderer(cfg, ckpt, 64, 64)
    R = 512
    rays = orc.synthetic_rays(R, frame=7, seed=11)
    g = torch.Generator().manual_seed(12)
    color_gt = torch.rand(R, 3, generator=g)
    depth_gt = torch.rand(R, 1, generator=g) * 0.5 + 0.5
    mask = (torch.rand(R, 1, generator=g) < 0.8).float()
    with torch.no_grad():
        z = r._sample_z(rays.cuda(), it, False).cpu()

    def oracle_loss(dt):
        return lambda orc_, net: _trainer_loss(
            orc_.render_rays(net, rc, rays.to(dt), iter_step=it, perturb_overwrite=False, z_vals_override=z.to(dt)),
            color_gt.to(dt), depth_gt.to(dt), mask.to(dt))

    ref_loss, ref32 = _oracle_grads(ckpt, nc, oracle_loss(torch.float32))
    torch.set_default_dtype(torch.float64)
    try:
        _, ref64 = _oracle_grads({n: {k: v.double() for k, v in sd.items()} for n, sd in ckpt.items()}, nc,
                                 oracle_loss(torch.float64))
    finally:
        torch.set_default_dtype(torch.float32)
    o = r.render_rays(rays.cuda(), iter_step=it, perturb_overwrite=False, z_vals_override=z.cuda())
    loss = _trainer_loss(o, color_gt.cuda(), depth_gt.cuda(), mask.cuda())
    assert rel_err(loss.detach(), ref_loss) < 1e-4
    mine = _my_grads(r, loss)
    r.sync_check()
    num = sum(((mine[k].double() - v) ** 2).sum().item() for k, v in ref64.items())
    den = sum((v ** 2).sum().item() for v in ref64.values())
    glob = (num / den) ** 0.5
    rows = sorted(((_rel(mine[k], v), _rel(ref32[k], v), k) for k, v in ref64.items()), reverse=True)
    print(f"it={it}: global {glob:.3e}; worst (ours, fp32 reference, tensor): {rows[:5]}")
    assert all(torch.isfinite(v).all() for v in mine.values())
    assert glob <= 1e-4, f"global gradient error {glob:.3e}"
    bad = [(e, e32, k) for e, e32, k in rows if e > max(1e-3, 5.0 * e32)]
    assert not bad, f"per-tensor gradient error (ours, fp32 reference's own, tensor): {bad[:8]}"


def test_perturbed_sampling_with_injected_jitter(cfg, ckpt):
    """perturb=True: the per-ray jitter is drawn with torch.rand on the device like the reference (endosurf.py:81); with
    the same numbers injected into the oracle the sample positions and the rendered outputs agree."""
    from oracle import endosurf_oracle as orc
    r, rc, nc = _renderer(cfg, ckpt, 32, 32)
    R = 64
    rays = orc.synthetic_rays(R, frame=3, seed=5)
    torch.manual_seed(123)
    t_rand = (torch.rand([R, 1], device="cuda") - 0.5).cpu()
    torch.manual_seed(123)
    with torch.no_grad():
        z = r._sample_z(rays.cuda(), 50000, True).cpu()
    net = orc.OracleNet(ckpt, nc)
    with torch.no_grad():
        ref = orc.render_rays(net, rc, rays, iter_step=50000, perturb_overwrite=True, t_rand=t_rand)
    # hierarchical resampling is discontinuous in the coarse sdf: quantile gate as in test_gpu_parity
    assert_close("z_vals(perturb)", z, ref["z_vals"], 1e-4, kink_tol=5e-2, q=0.98)
    # coarse samples carry exactly the jitter: first / last sample positions per ray move by t_rand * 2/n_samples
    with torch.no_grad():
        z0 = r._sample_z(rays.cuda(), 50000, False).cpu()
    assert not torch.equal(z, z0)


def test_single_pass_fp16_forward_and_gradients(cfg, ckpt):
    """precision_terms=1 (one fp16 tensor-core pass instead of the 3-term split, the bf16/fp16 training config):
    forward within 5e-3 of the oracle, gradients within 5e-2 per tensor / 2e-2 overall."""
    g = load_npz("render_r32_s32_i32_it25k.npz")
    ns, ni = int(g["n_samples"]), int(g["n_importance"])
    r, rc, nc = _renderer(cfg, ckpt, ns, ni, True, False, precision_terms=1)
    rays = torch.from_numpy(g["rays"])
    z = torch.from_numpy(g["z_vals"])
    it = int(g["iter_step"])

    def lossf(o):
        return o["color_map"].sum() * 0.7 + o["depth_map"].sum() * 0.3 + o["gradient_o_error"] * 0.1

    ref_loss, ref = _oracle_grads(ckpt, nc, lambda orc, net: lossf(
        orc.render_rays(net, rc, rays, iter_step=it, perturb_overwrite=False, z_vals_override=z)))
    o = r.render_rays(rays.cuda(), iter_step=it, perturb_overwrite=False, z_vals_override=z.cuda())
    for k in ["color_map", "depth_map"]:
        assert_close(k, o[k].detach(), g["core/" + k], 5e-3, kink_tol=5e-2)
    mine = _my_grads(r, lossf(o))
    r.sync_check()
    _compare(mine, ref, tol_norm=5e-2, tol_global=2e-2)


def test_training_launch_budget(cfg, ckpt):
    """One training step of render_rays (forward + backward) stays within 60 library launches after the sampling, and
    none of them is a PyTorch / library GEMM kernel (the library counts its own launches)."""
    from oracle import endosurf_oracle as orc
    r, rc, nc = _renderer(cfg, ckpt, 16, 16)
    rays = orc.synthetic_rays(64, frame=2, seed=4).cuda()
    with torch.no_grad():
        z = r._sample_z(rays, 1000, False)
    n0 = r.launch_count()
    o = r.render_rays(rays, iter_step=1000, z_vals_override=z)
    (o["color_map"].sum() + o["gradient_o_error"]).backward()
    r.sync_check()
    assert r.launch_count() - n0 <= 60, r.launch_count() - n0


def test_data_parallel_shards_equal_the_union_batch(cfg, ckpt):
    """SURVEY 8e with the real renderer: two ray shards rendered separately (as two ranks would), their masked-mean
    numerators backpropagated with the GLOBAL denominators and the gradients summed, against one call on the union
    batch.  (The collective itself is covered on CPU with gloo, tests/test_dist_cpu.py.)"""
    from oracle import endosurf_oracle as orc
    from endosurf_b200 import distributed as dp
    r, rc, nc = _renderer(cfg, ckpt, 32, 32)
    R = 256
    rays = orc.synthetic_rays(R, frame=9, seed=21).cuda()
    g = torch.Generator().manual_seed(22)
    color_gt = torch.rand(R, 3, generator=g).cuda()
    depth_gt = (torch.rand(R, 1, generator=g) * 0.5 + 0.5).cuda()
    mask = (torch.rand(R, 1, generator=g) < 0.7).float().cuda()
    params = [p for v in r.get_train_params().values() for p in v]

    def grads_of(loss):
        r.zero_grad()
        loss.backward()
        return torch.cat([p.grad.reshape(-1) for p in params]).clone()

    def terms_of(sl):
        o = r.render_rays(rays[sl], iter_step=25000, perturb_overwrite=False)
        return dp.render_loss_terms(r, o, color_gt[sl], depth_gt[sl], mask[sl], mask[sl])

    t_all, eps = terms_of(slice(0, R))
    ref = grads_of(sum(w * num / (den + eps[k]) for k, (w, num, den) in t_all.items()))
    shards = [dp.shard_rays(R, 2, k) for k in range(2)]
    t_loc = [terms_of(sl)[0] for sl in shards]
    dens = {k: sum(t[k][2] for t in t_loc) + eps[k] for k in eps}
    for k in eps:  # the shards' denominators add up to the union's
        assert torch.allclose(dens[k], t_all[k][2] + eps[k], rtol=1e-6), k
    total = sum(grads_of(sum(t[k][0] * t[k][1] / dens[k] for k in eps)) for t in t_loc)
    r.sync_check()
    err = ((total - ref).norm() / ref.norm()).item()
    assert err < 2e-5, f"sum of shard gradients vs union-batch gradient: rel err {err:.3e}"


def test_gradient_sink_equals_autograd_accumulation(cfg, ckpt):
    """distributed.FlatGradBucket.bind: the backward adds its gradients into the flat bucket with one launch instead of
    handing 82 tensors to autograd - the bucket must end up bit-identical to the unbound path, across two accumulated
    backward calls (render_rays + a point-field loss, as the reference's train_step has)."""
    from oracle import endosurf_oracle as orc
    from endosurf_b200 import distributed as dp
    r, rc, nc = _renderer(cfg, ckpt, 16, 16)
    rays = orc.synthetic_rays(96, frame=5, seed=8).cuda()
    with torch.no_grad():
        z = r._sample_z(rays, 1000, False)
    params = [p for v in r.get_train_params().values() for p in v]
    bucket = dp.FlatGradBucket(params)
    x = torch.rand(160, 3, device="cuda") - 0.5
    d = torch.nn.functional.normalize(torch.randn(160, 3, device="cuda"), dim=-1)
    t = torch.rand(160, 1, device="cuda")

    def run(bound):
        r._grad_sink = bucket if bound else None
        bucket.zero()
        o = r.render_rays(rays, iter_step=1000, z_vals_override=z)
        sdf, g_c, jac, rgb = r.point_field(x, d, t)
        loss = o["color_map"].sum() + o["depth_map"].sum() + o["gradient_o_error"] + sdf.abs().mean() + rgb.mean()
        n0 = r.launch_count()
        loss.backward()
        r.sync_check()
        return bucket.flat.clone()

    a, b = run(False), run(True)
    r._grad_sink = None
    assert torch.equal(a, b), (a - b).abs().max().item()
    assert a.abs().sum().item() > 0


def test_bench_size_parity_against_the_reference_on_this_gpu():
    """BASELINE configs[1] at its full size, the exact configuration bench.py times: 4096 rays of a 512x512 frame,
    64 + 64 samples, 4 up-sampling steps, the bench's seeded random-init state and masked-mean loss - one training
    step (forward, loss, backward) of the UNMODIFIED reference (oracle/_ref, stock PyTorch fp32 on this GPU, TF32 off)
    against this library: rendered maps, loss and the gradients of all 83 parameter tensors."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("oracle/_ref (byte-compiled reference) has not been built: python oracle/build_ref.py")
    from endosurf_b200 import EndoSurfRenderer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    R = 4096
    rays = bench.make_rays(R, frame=7).cuda()
    cgt, dgt, msk = (x.cuda() for x in bench.make_targets(R, frame=7))
    mod = ref_shims.load_reference()
    torch.manual_seed(0)
    ref = mod.EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), copy.deepcopy(bench.NET_CFG), "cuda")
    bench.seeded_state(ref.model)
    ref.train()
    o_ref = ref(rays, iter_step=bench.ITER_STEP, perturb_overwrite=False)
    l_ref = bench.train_loss(o_ref, cgt, dgt, msk)
    l_ref.backward()
    g_ref = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    maps_ref = {k: o_ref[k].detach().clone() for k in ("color_map", "depth_map", "gradient_o_error", "weight_max")}
    state = copy.deepcopy(ref.save_checkpoint())
    del o_ref, l_ref, ref
    torch.cuda.empty_cache()
    # the reference against itself with another net_chunk: a LOWER bound of its fp32 noise (cuBLAS computes a row the same
    # way whatever the number of rows, so only the reductions over rows change); tools/diag_bench_parity.py compares both
    # implementations with the float64 oracle on the rays that disagree (profiles/r2_bench_size_parity.txt)
    rc2 = copy.deepcopy(bench.RENDER_CFG)
    rc2["net_chunk"] = 50000
    ref2 = mod.EndoSurfRenderer(rc2, copy.deepcopy(bench.NET_CFG), "cuda")
    ref2.load_checkpoint(copy.deepcopy(state))
    ref2.train()
    o2 = ref2(rays, iter_step=bench.ITER_STEP, perturb_overwrite=False)
    bench.train_loss(o2, cgt, dgt, msk).backward()
    self_flips = {k: ((o2[k].detach() - maps_ref[k]).abs() / maps_ref[k].abs().max() > 1e-4).double().mean().item()
                  for k in ("color_map", "depth_map")}
    sn = sum(((p.grad - g_ref[k]) ** 2).sum().item() for k, p in ref2.named_parameters())
    sd = sum((g ** 2).sum().item() for g in g_ref.values())
    self_glob = (sn / sd) ** 0.5
    print(f"[bench-size parity] reference vs reference (net_chunk 80000 vs 50000): entries beyond 1e-4: colour "
          f"{100 * self_flips['color_map']:.3f} % depth {100 * self_flips['depth_map']:.3f} %, whole-gradient rel "
          f"{self_glob:.2e}")
    del o2, ref2
    torch.cuda.empty_cache()

    r = EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), bench.NET_CFG, device="cuda")
    r.load_checkpoint(state)
    r.train()
    o = r(rays, iter_step=bench.ITER_STEP, perturb_overwrite=False)
    loss = bench.train_loss(o, cgt, dgt, msk)
    loss.backward()
    r.sync_check()
    # rendered maps: per-ray integrals after hierarchical re-sampling (discontinuous: quantile gate, flips reported)
    for k in ("color_map", "depth_map"):
        e = (o[k].detach() - maps_ref[k]).abs() / maps_ref[k].abs().max()
        flips = (e > 1e-4).double().mean().item()
        print(f"[bench-size parity] {k}: p99.9 {torch.quantile(e.flatten().float(), 0.999).item():.2e} max "
              f"{e.max().item():.2e}, entries beyond 1e-4: {100 * flips:.3f} %")
        # measured: 0.77 % / 0.81 % of the entries beyond 1e-4, max 1.5e-3 / 1.7e-3 (ill-conditioned rays, see above)
        assert flips <= 0.02 and e.max().item() <= 1e-2, (k, flips, e.max().item())
    ge = abs(o["gradient_o_error"].item() - maps_ref["gradient_o_error"].item()) / abs(maps_ref["gradient_o_error"].item())
    print(f"[bench-size parity] gradient_o_error rel {ge:.2e}")
    assert ge <= 1e-4
    worst, num, den = ("", 0.0), 0.0, 0.0
    for k, p in r.named_parameters():
        d = (p.grad - g_ref[k]).norm().item()
        n = g_ref[k].norm().item()
        num += d * d
        den += n * n
        if d / max(n, 1e-30) > worst[1]:
            worst = (k, d / max(n, 1e-30))
    glob = (num / den) ** 0.5
    print(f"[bench-size parity] gradients of {len(g_ref)} tensors: whole-gradient rel err {glob:.2e}, worst tensor "
          f"{worst[0]} {worst[1]:.2e}")
    # measured: whole gradient 5.6e-4, worst tensor (first colour layer, |g| ~ 1e-4 of the largest) 3.5e-3
    assert glob <= 1e-3 and worst[1] <= 1e-2, (glob, worst)

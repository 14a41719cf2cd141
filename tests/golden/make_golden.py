"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
  1. imports ``/root/reference/src/renderer/endosurf.py`` with the unrelated missing
     third-party modules (mcubes, kornia, lpips, open3d, imageio, trimesh, wandb) stubbed
     in ``sys.modules`` -- the reference source is executed, never copied;
  2. builds ``EndoSurfRenderer`` on ``configs/endosurf/baseline/base_pull.yml`` (seed 0),
     adds seeded noise to every parameter so the deformation net is non-trivial but a surface survives;
  3. runs the reference on seeded synthetic rays / points and stores inputs + outputs;
  4. checks ``oracle/endosurf_oracle.py`` against the reference on the same inputs and
     refuses to write fixtures if the restatement disagrees (this is the oracle's pin).
"""
import os
import sys
import copy
from unittest import mock

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    for name in ["mcubes", "kornia", "kornia.losses", "lpips", "open3d", "imageio", "imageio.v2", "trimesh",
                 "wandb", "cv2"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    try:
        from src.renderer.endosurf import EndoSurfRenderer  # noqa
    finally:
        os.chdir(cwd)
    return EndoSurfRenderer


def load_cfg():
    with open(os.path.join(REF, "configs/endosurf/baseline/base_pull.yml")) as f:
        return yaml.safe_load(f)


def t2n(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def flat_ckpt(ckpt):
    out = {}
    for net, sd in ckpt.items():
        for k, v in sd.items():
            out[f"{net}/{k}"] = v.detach().cpu().numpy()
    return out


def main():
    torch.set_default_dtype(torch.float32)
    torch.set_num_threads(8)
    from oracle import endosurf_oracle as orc

    Renderer = import_reference()
    cfg = load_cfg()
    render_cfg = copy.deepcopy(cfg["render"])
    net_cfg = copy.deepcopy(cfg["net"])

    import json
    with open(os.path.join(HERE, "base_pull_cfg.json"), "w") as f:  # cfg["render"], cfg["net"] of base_pull.yml
        json.dump({"render": cfg["render"], "net": cfg["net"]}, f, indent=1)

    torch.manual_seed(0)
    ref = Renderer(render_cfg, net_cfg, device="cpu")
    with torch.no_grad():
        # 0.02*randn on deform / colour / variance (non-trivial deformation), 0.004*randn on the SDF net: the
        # geometric init survives as a surface (ray weights sum to ~0.9-1.0, every ray_marching ray hits), which a
        # 0.02 perturbation of the SDF net destroys (weights ~1e-2, no zero crossing)
        for n, p in ref.named_parameters():
            p.add_((0.004 if "sdf_network" in n else 0.02) * torch.randn_like(p))
    ckpt = {k: {kk: vv.clone() for kk, vv in sd.items()} for k, sd in ref.save_checkpoint().items()}
    np.savez(os.path.join(HERE, "ckpt_base_pull.npz"), **flat_ckpt(ckpt))
    net = orc.OracleNet(ckpt, net_cfg)

    report = []

    def check(name, a, b, tol=2e-6, kink_tol=None):
        """rel-max error.  kink_tol: quantities that contain derivatives of the ReLU deformation net are
        discontinuous where a pre-activation crosses 0, so two fp32 evaluation orders (e.g. different net_chunk
        splits) can legitimately disagree on a handful of samples; for those the 99.5th percentile must meet
        `tol` and the maximum only `kink_tol`."""
        a, b = a.detach().double(), b.detach().double()
        scale = max(b.abs().max().item(), 1e-12)
        e = (a - b).abs().flatten() / scale
        err = e.max().item()
        report.append((name, err))
        if kink_tol is not None:
            q = torch.quantile(e, 0.995).item() if e.numel() > 1 else err
            if not (q <= tol and err <= kink_tol):
                raise SystemExit(f"oracle disagrees with reference on {name}: p99.5 {q:.3e} max {err:.3e}")
        elif not err <= tol:
            raise SystemExit(f"oracle disagrees with reference on {name}: rel-max err {err:.3e} > {tol}")

    # ---------------- stage goldens: per-point network quantities on 192 points
    g = torch.Generator().manual_seed(1)
    n_pts = 192
    x = (torch.rand(n_pts, 3, generator=g) * 2 - 1) * 0.8
    d = torch.randn(n_pts, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    t = torch.rand(n_pts, 1, generator=g)
    m = ref.model
    with torch.enable_grad():
        xg = x.clone().requires_grad_(True)
        r_xc = xg + m.deform_network(xg, t)
        r_h = m.sdf_network(r_xc)
        r_gc = m.get_sdf_grad_from_canonical_space(r_xc)
        r_J = m.get_deform_grad_from_observed_space(xg, t)
        r_go = m.get_sdf_grad_from_observed_space(x.clone(), t)
        r_raw = m.forward(torch.cat([x, d, t], -1))
    r_sdf_obs = m.get_sdf_from_observed_space(x, t)
    parts = net.forward_parts(torch.cat([x, d, t], -1))
    check("x_c", parts["x_c"], r_xc)
    check("sdf", parts["sdf"], r_h[:, :1])
    check("feat", parts["feat"], r_h[:, 1:])
    check("g_c", parts["g_c"], r_gc)
    check("jac", parts["jac"], r_J)
    check("g_o", parts["g_o"], r_go)
    check("rgb", parts["rgb"], r_raw[:, 1:4])
    check("sdf_obs", net.sdf_from_observed(x, t), r_sdf_obs)
    np.savez(os.path.join(HERE, "stage_points.npz"), x=x.numpy(), d=d.numpy(), t=t.numpy(),
             x_c=r_xc.detach().numpy(), sdf=r_h[:, :1].detach().numpy(), feat=r_h[:, 1:].detach().numpy(),
             g_c=r_gc.detach().numpy(), jac=r_J.detach().numpy(), g_o=r_go.detach().numpy(),
             rgb=r_raw[:, 1:4].detach().numpy(),
             enc6=m.sdf_network.enc_fn_pos(x, bound=1.0).numpy(),
             enc10=m.color_network.enc_fn_pos(x, bound=1.0).numpy())

    # ---------------- render goldens
    def run_case(tag, n_rays, ns, ni, iter_step, use_deform=True, frame=17):
        rc = copy.deepcopy(render_cfg)
        rc.update(n_samples=ns, n_importance=ni, perturb=False)
        nc = copy.deepcopy(net_cfg)
        nc["use_deform"] = use_deform
        torch.manual_seed(0)
        r = Renderer(rc, nc, device="cpu")
        sub = {k: v for k, v in ckpt.items() if use_deform or k != "deform_network"}
        r.load_checkpoint(sub)
        rays = orc.synthetic_rays(n_rays, frame=frame, seed=3)
        out = r.render_rays(rays, iter_step=iter_step, perturb_overwrite=False)
        onet = orc.OracleNet(sub, nc)
        o = orc.render_rays(onet, rc, rays, iter_step=iter_step, perturb_overwrite=False)
        for k in ["color_map", "depth_map", "gradient_o_error", "weight_max", "s_val"]:
            check(f"{tag}/{k}", o[k], out[k], tol=5e-5)
        # per-sample tensors: resampling (sort/searchsorted) is discontinuous, so they are pinned through
        # render_core on FIXED z_vals (the oracle's), ref endosurf.py:134-213
        z = o["z_vals"].detach()
        cr = cos = orc.cos_anneal_ratio(iter_step, rc["anneal_end"])
        core = r.render_core(rays[:, :3], rays[:, 3:6], rays[:, 8], z, 2.0 / ns, cos_anneal_ratio=cr)
        ocore = orc.render_core(onet, rays[:, :3], rays[:, 3:6], rays[:, 8], z, 2.0 / ns, cos_ratio=cr)
        for k in ["color_map", "depth_map", "gradient_o_error", "weights", "cdf", "gradients_o"]:
            check(f"{tag}/core/{k}", ocore[k], core[k], tol=2e-5, kink_tol=5e-3)
        trace = orc.hierarchical_z_vals(onet, rays, orc.coarse_z_vals(rays, ns), ni, rc["up_sample_steps"],
                                        return_trace=True)[1] if ni > 0 else []
        # training oracle: d(loss)/d(params) for a fixed scalar loss on the 8-key dict
        loss = (out["color_map"].sum() * 0.7 + out["depth_map"].sum() * 0.3 + out["gradient_o_error"] * 0.1)
        grads = torch.autograd.grad(loss, [p for p in r.parameters()], allow_unused=True)
        gnorm = {n: (g_.detach().norm().item() if g_ is not None else 0.0)
                 for (n, _), g_ in zip(r.named_parameters(), grads)}
        save = dict(rays=rays.numpy(), iter_step=np.int64(iter_step), n_samples=np.int64(ns),
                    n_importance=np.int64(ni), use_deform=np.bool_(use_deform), z_vals=z.numpy(),
                    sdf=o["sdf"].detach().numpy(), sampled_color=o["sampled_color"].detach().numpy())
        save.update({k: v for k, v in t2n({kk: out[kk] for kk in out}).items()})
        save.update({"core/" + k: v for k, v in t2n({kk: core[kk] for kk in core}).items()})
        for i, tr in enumerate(trace):
            save[f"up{i}_z"], save[f"up{i}_sdf"], save[f"up{i}_new_z"] = (a.numpy() for a in tr)
        for n, v in gnorm.items():
            save["gradnorm/" + n] = np.float64(v)
        # a few full gradients (small ones + one 256x256) for the training parity test
        named = dict(zip([n for n, _ in r.named_parameters()], grads))
        for n in ["model.deviation_network.variance", "model.sdf_network.net.8.bias", "model.sdf_network.net.2.weight_v",
                  "model.color_network.net.8.weight_v", "model.deform_network.net.8.weight_v",
                  "model.deform_network.net.0.weight_g"]:
            if n in named and named[n] is not None:
                save["grad/" + n] = named[n].detach().numpy()
        np.savez(os.path.join(HERE, f"render_{tag}.npz"), **save)

    run_case("r48_s64_i64_it0", 48, 64, 64, 0)
    run_case("r48_s64_i64_it50k", 48, 64, 64, 50000)
    run_case("r32_s32_i32_it25k", 32, 32, 32, 25000, frame=41)
    run_case("r32_s64_i0_it0", 32, 64, 0, 0, frame=5)
    run_case("r32_nodeform_s32_i32", 32, 32, 32, 10000, use_deform=False, frame=9)

    # ---------------- helper goldens (section 8f): errorondepth / ray_marching
    rc = copy.deepcopy(render_cfg)
    rays = orc.synthetic_rays(40, frame=23, seed=11)
    torch.manual_seed(5)
    d_i = ref.ray_marching(rays)
    check("ray_marching", torch.nan_to_num(orc.ray_marching(net, rays), posinf=1e9),
          torch.nan_to_num(d_i, posinf=1e9), tol=1e-4)
    d_gt = torch.where(torch.isfinite(d_i) & (d_i > 0), d_i, torch.full_like(d_i, 1.0)) + 0.01
    mask = (torch.arange(40) % 5 != 0).float()[:, None]
    se, ae, ins = ref.errorondepth(rays, d_gt, mask)
    ose, oae, oins = orc.errorondepth(net, rays, d_gt, mask)
    check("errorondepth/sdf", ose, se, tol=1e-4)  # sum of near-zero sdf values: fp32 noise is relatively large
    check("errorondepth/angle", oae, ae, tol=1e-4)
    np.savez(os.path.join(HERE, "helpers.npz"), rays=rays.numpy(), d_i=d_i.numpy(), d_gt=d_gt.numpy(),
             mask=mask.numpy(), sdf_err=se.detach().numpy(), angle_err=ae.detach().numpy(), inside=ins.numpy())

    with open(os.path.join(HERE, "ORACLE_PIN.txt"), "w") as f:
        f.write("oracle/endosurf_oracle.py vs reference (rel-max error; written by make_golden.py)\n")
        f.write(f"torch {torch.__version__}, cpu, fp32, reference commit 2b33413f\n")
        for n, e in report:
            f.write(f"{n:40s} {e:.3e}\n")
    for n, e in report:
        print(f"{n:40s} {e:.3e}")


if __name__ == "__main__":
    main()

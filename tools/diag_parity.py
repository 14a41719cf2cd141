"""Error distributions of the CUDA path vs golden / oracle (GPU box): python tools/diag_parity.py"""
import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_cfg, load_ckpt, load_npz
from endosurf_b200 import EndoSurfRenderer
from oracle import endosurf_oracle as orc

def q(name, a, b):
    a = torch.as_tensor(a).double().cpu().flatten(); b = torch.as_tensor(b).double().cpu().flatten()
    e = (a - b).abs() / b.abs().max()
    qs = torch.quantile(e, torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.double)) if e.numel() > 1 else e.repeat(4)
    print(f"  {name:18s} p50 {qs[0]:.2e} p90 {qs[1]:.2e} p99 {qs[2]:.2e} p99.9 {qs[3]:.2e} max {e.max():.2e}  n>1e-4: {(e>1e-4).sum().item()}/{e.numel()}")

cfg = load_cfg(); ckpt = load_ckpt()
def rend(ns, ni):
    rc = copy.deepcopy(cfg["render"]); rc.update(n_samples=ns, n_importance=ni, perturb=False)
    r = EndoSurfRenderer(rc, cfg["net"], device="cuda"); r.load_checkpoint(ckpt); r.eval(); return r, rc
for tag in ["r48_s64_i64_it0", "r48_s64_i64_it50k"]:
    g = load_npz(f"render_{tag}.npz"); r, rc = rend(64, 64)
    rays = torch.from_numpy(g["rays"]).cuda(); z = torch.from_numpy(g["z_vals"]).cuda()
    with torch.no_grad():
        o = r.render_rays(rays, iter_step=int(g["iter_step"]), z_vals_override=z, return_extras=True)
        e = r.render_rays(rays, iter_step=int(g["iter_step"]), perturb_overwrite=False, return_extras=True)
    print(tag, "fixed z")
    for k in ["color_map", "depth_map", "weights", "cdf", "gradients_o"]: q(k, o[k], g["core/" + k])
    q("sdf", o["sdf"], g["sdf"]); q("sampled_color", o["sampled_color"], g["sampled_color"])
    print(tag, "end to end")
    for k in ["color_map", "depth_map", "weight_max"]: q(k, e[k], g[k])
    q("z_vals", e["z_vals"], g["z_vals"])
r, rc = rend(64, 64)
rays = orc.synthetic_rays(256, frame=33, seed=21)
net = orc.OracleNet(ckpt, cfg["net"])
with torch.no_grad():
    ref = orc.render_rays(net, rc, rays, iter_step=50000, perturb_overwrite=False)
    o = r.render_rays(rays.cuda(), iter_step=50000, perturb_overwrite=False, return_extras=True)
    zc = orc.coarse_z_vals(rays, 64)
    z_ref, trace = orc.hierarchical_z_vals(net, rays, zc, 64, 4, return_trace=True)
print("oracle 256 rays end to end")
for k in ["color_map", "depth_map", "weight_max", "z_vals"]: q(k, o[k], ref[k])
# where do z_vals diverge first? coarse sdf then each up-sampling step with the ORACLE's inputs
pts = (rays[:, None, :3] + (rays[:, 3:6] / (rays[:, 5:6] + 1e-6))[:, None, :] * zc[..., None]).reshape(-1, 3)
tt = rays[:, None, 8:9].expand(256, 64, 1).reshape(-1, 1)
sd = r.sdf_from_observed_space(pts.cuda(), tt.cuda()).reshape(256, 64)
q("coarse sdf", sd, trace[0][1])
for i, (z, s, nz) in enumerate(trace):
    out = r.up_sample(rays[:, :3].cuda(), rays[:, 3:6].cuda(), z.cuda(), s.cuda(), 16, 64 * 2 ** i)
    q(f"up_sample[{i}] same in", out, nz)
bad = ((o["color_map"].cpu() - ref["color_map"]).abs().max(-1)[0]).argmax().item()
print("worst ray", bad, "dz max", (o["z_vals"].cpu()[bad] - ref["z_vals"][bad]).abs().max().item(),
      "color", o["color_map"][bad].cpu(), ref["color_map"][bad])

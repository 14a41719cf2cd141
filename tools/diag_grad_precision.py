"""Which approximation of the default training mode costs gradient accuracy?  512 rays x 128 samples against the
oracle's autograd under: default planes, every lo plane (3-term weight gradients, exact gates), larger adjoint scale."""
import copy, sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_cfg, load_ckpt
from test_gpu_training import _renderer, _oracle_grads, _my_grads, _trainer_loss
from oracle import endosurf_oracle as orc
from endosurf_b200 import _lib

cfg, ckpt = load_cfg(), load_ckpt()
R, it = int(os.environ.get("R", 512)), 50000
rays = orc.synthetic_rays(R, frame=7, seed=11)
g = torch.Generator().manual_seed(12)
color_gt = torch.rand(R, 3, generator=g); depth_gt = torch.rand(R, 1, generator=g) * 0.5 + 0.5
mask = (torch.rand(R, 1, generator=g) < 0.8).float()
r, rc, nc = _renderer(cfg, ckpt, 64, 64)
with torch.no_grad():
    z = r._sample_z(rays.cuda(), it, False).cpu()
ref_loss, ref = _oracle_grads(ckpt, nc, lambda o_, net: _trainer_loss(
    o_.render_rays(net, rc, rays, iter_step=it, perturb_overwrite=False, z_vals_override=z), color_gt, depth_gt, mask))
# a second oracle evaluation in float64 tells how much of the difference is the fp32 oracle's own rounding
ck64 = {n: {k: v.double() for k, v in sd.items()} for n, sd in ckpt.items()}
torch.set_default_dtype(torch.float64)
try:
    _, ref64 = _oracle_grads(ck64, nc, lambda o_, net: _trainer_loss(
        o_.render_rays(net, rc, rays.double(), iter_step=it, perturb_overwrite=False, z_vals_override=z.double()),
        color_gt.double(), depth_gt.double(), mask.double()))
except Exception as e:
    print("float64 oracle failed:", e); ref64 = None
torch.set_default_dtype(torch.float32)

def report(tag, mine, ref):
    num = sum(((mine[k].double() - v.double()) ** 2).sum().item() for k, v in ref.items())
    den = sum((v.double() ** 2).sum().item() for v in ref.values())
    rows = sorted(((mine[k].double() - v.double()).norm().item() / max(v.double().norm().item(), 1e-30), k) for k, v in ref.items())[::-1]
    print(f"{tag}: global {(num/den)**0.5:.3e}; worst " + ", ".join(f"{k.replace('_network.net','')}={e:.1e}" for e, k in rows[:6]), flush=True)

if ref64 is not None:
    report("oracle fp32 vs oracle fp64", {k: v for k, v in ref.items()}, ref64)
for tag, full, tgt in [("default", 0, 4), ("full planes", 1, 4), ("default, target 2^10", 0, 10), ("full, target 2^10", 1, 10)]:
    lib, ctx = _lib.load(), r._context()
    lib.es_set_plane_mode(ctx, full); lib.es_debug_set(ctx, 3, tgt)
    o = r.render_rays(rays.cuda(), iter_step=it, perturb_overwrite=False, z_vals_override=z.cuda())
    mine = _my_grads(r, _trainer_loss(o, color_gt.cuda(), depth_gt.cuda(), mask.cuda()))
    r.sync_check()
    report(tag + " vs fp32 oracle", mine, ref)
    if ref64 is not None:
        report(tag + " vs fp64 oracle", mine, {k: v.float() for k, v in ref64.items()})

"""GPU box: who is right on the rays where this library and the reference (fp32, same GPU) disagree at bench size?
Both are compared with the oracle evaluated in float64 on the host (the algorithm's exact answer) on the 48 worst and 48
random rays of a 4096-ray batch (forward, perturb off, bench state)."""
import copy, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import ref_shims, endosurf_oracle as orc
from endosurf_b200 import EndoSurfRenderer
torch.backends.cuda.matmul.allow_tf32 = False
R = 4096
rays = bench.make_rays(R, frame=7).cuda()
mod = ref_shims.load_reference()
torch.manual_seed(0)
ref = mod.EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), copy.deepcopy(bench.NET_CFG), "cuda")
bench.seeded_state(ref.model); ref.eval()
o_ref = ref(rays, iter_step=bench.ITER_STEP, perturb_overwrite=False)
o_ref = {k: o_ref[k].detach() for k in ("color_map", "depth_map")}
state = copy.deepcopy(ref.save_checkpoint())
r = EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), bench.NET_CFG, device="cuda")
r.load_checkpoint(state); r.eval()
with torch.no_grad():
    o = r.render_rays(rays, iter_step=bench.ITER_STEP, perturb_overwrite=False)
r.sync_check()
dev = ((o["depth_map"] - o_ref["depth_map"]).abs() / o_ref["depth_map"].abs().max()).flatten()
worst = torch.argsort(dev, descending=True)[:48].cpu()
rnd = torch.randperm(R, generator=torch.Generator().manual_seed(3))[:48]
ck64 = {n: {k: v.detach().double().cpu() for k, v in sd.items()} for n, sd in state.items()}
net64 = orc.OracleNet(ck64, bench.NET_CFG)
rc = copy.deepcopy(bench.RENDER_CFG)
for name, idx in (("48 worst rays (ours vs reference)", worst), ("48 random rays", rnd)):
    with torch.no_grad():
        t = orc.render_rays(net64, rc, rays[idx].double().cpu(), iter_step=bench.ITER_STEP, perturb_overwrite=False)
    for k in ("color_map", "depth_map"):
        sc = o_ref[k].abs().max().item()
        eo = ((o[k][idx].double().cpu() - t[k]).abs() / sc)
        er = ((o_ref[k][idx].double().cpu() - t[k]).abs() / sc)
        eb = ((o[k][idx] - o_ref[k][idx]).abs() / sc).double().cpu()
        print(f"{name:36s} {k:10s} |ours-fp64| median {eo.median():.2e} max {eo.max():.2e}   |ref32-fp64| median "
              f"{er.median():.2e} max {er.max():.2e}   |ours-ref32| median {eb.median():.2e} max {eb.max():.2e}")

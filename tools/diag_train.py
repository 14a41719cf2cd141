"""Gradient error table of the training path vs the oracle (GPU box): python tools/diag_train.py"""
import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_cfg, load_ckpt, load_npz
from endosurf_b200 import EndoSurfRenderer
from oracle import endosurf_oracle as orc
cfg, ckpt = load_cfg(), load_ckpt()
s = load_npz("stage_points.npz"); n = 64
x, d, t = (torch.from_numpy(s[k][:n]) for k in "xdt")
g = torch.Generator().manual_seed(3)
a_sdf, a_go, a_rgb = torch.randn(n, 1, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "sdf": a_go *= 0; a_rgb *= 0
if which == "rgb": a_sdf *= 0; a_go *= 0
if which == "go": a_sdf *= 0; a_rgb *= 0
ck = {nn: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for nn, sd in ckpt.items()}
net = orc.OracleNet(ck, cfg["net"])
raw = net.forward(torch.cat([x, d, t], -1)); g_o = net.sdf_grad_observed(x.clone(), t)
ref_loss = (raw[:, :1] * a_sdf).sum() + (raw[:, 1:4] * a_rgb).sum() + (g_o * a_go).sum(); ref_loss.backward()
rc = copy.deepcopy(cfg["render"]); r = EndoSurfRenderer(rc, cfg["net"], device="cuda"); r.load_checkpoint(ckpt); r.train()
sdf, g_c, jac, rgb = r.point_field(x.cuda(), d.cuda(), t.cuda())
go = torch.einsum("nij,ni->nj", jac, g_c)
loss = (sdf * a_sdf.cuda()).sum() + (rgb * a_rgb.cuda()).sum() + (go * a_go.cuda()).sum()
print("loss", loss.item(), ref_loss.item())
loss.backward(); r.sync_check()
for name, p in r.model.named_parameters():
    nn, rest = name.split(".", 1)
    gr = ck[nn][rest].grad
    if p.grad is None or gr is None:
        print(name, "no grad"); continue
    gm = p.grad.cpu()
    e = (gm - gr).norm().item() / max(gr.norm().item(), 1e-12)
    flag = "  <<<<" if e > 2e-3 else ""
    print(f"{name:42s} ref {gr.norm().item():10.3e} mine {gm.norm().item():10.3e} rel err {e:9.2e}{flag}")

"""GPU box: which part of a training step can be captured in a CUDA graph (forward / point field / compositing / weight-norm fold /
full colour loss), each variant in a fresh process.  Found torch.cumprod's backward (host sync) -> training._CumprodPos."""
import copy, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) < 2:
    for w in ["pf_sdf", "pf_rgb", "composite_only", "effw", "color"]:
        for mode in ["thread_local", "relaxed"]:
            subprocess.call([sys.executable, __file__, w, mode])
    sys.exit(0)
import torch
sys.path.insert(0, ROOT)
import bench
from endosurf_b200 import EndoSurfRenderer, training
which, mode = sys.argv[1], sys.argv[2]
R = 256
torch.manual_seed(0)
r = EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), bench.NET_CFG, device="cuda"); bench.seeded_state(r.model); r.train()
rays = bench.make_rays(R, 3).cuda(); cgt, dgt = (x.cuda() for x in bench.make_targets(R, 3))
n = 32768
x = (torch.rand(n, 3, device="cuda") - 0.5); d = torch.randn(n, 3, device="cuda"); t = torch.rand(n, 1, device="cuda")
leaf = [torch.rand(64, 32, device="cuda", requires_grad=True), torch.randn(64, 32, 3, device="cuda", requires_grad=True),
        torch.rand(64, 32, 3, device="cuda", requires_grad=True)]
def run():
    if which.startswith("pf"):
        sdf, g_c, jac, rgb = r.point_field(x, d, t)
        loss = sdf.sum() if which == "pf_sdf" else rgb.sum()
    elif which == "composite_only":
        z = torch.linspace(0, 1, 32, device="cuda").expand(64, 32).contiguous()
        o = training.composite(leaf[0] - 0.5, leaf[1], leaf[2], torch.randn(64, 3, device="cuda"), torch.randn(64, 32, 3, device="cuda"),
                               z, 0.03, torch.tensor(20.0, device="cuda"), 1.0)
        loss = o["color_map"].sum() + o["gradient_o_error"]
    elif which == "effw":
        loss = sum(w.sum() for w in training.effective_weights(r.model))
    else:
        o = r(rays, iter_step=50000); loss = o["color_map"].sum()
    loss.backward()
    return loss
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        r.zero_grad(set_to_none=True); run()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
r.zero_grad(set_to_none=True)
for l in leaf: l.grad = None
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode=mode):
        l = run()
    g.replay(); torch.cuda.synchronize()
    print(which, mode, "capture OK", float(l.detach()), flush=True)
except Exception as e:
    print(which, mode, "capture FAILED:", str(e).split("\n")[0][:200], flush=True)

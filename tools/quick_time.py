"""First timing of the forward path (GPU box): python tools/quick_time.py [rays]"""
import copy, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_cfg, load_ckpt
from endosurf_b200 import EndoSurfRenderer
from oracle import endosurf_oracle as orc
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = load_cfg(); rc = copy.deepcopy(cfg["render"]); rc.update(n_samples=64, n_importance=64, perturb=True)
r = EndoSurfRenderer(rc, cfg["net"], device="cuda"); r.load_checkpoint(load_ckpt()); r.eval()
rays = orc.synthetic_rays(R, frame=3, seed=1).cuda()
with torch.no_grad():
    for _ in range(3): r.render_rays(rays, iter_step=50000)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): r.render_rays(rays, iter_step=50000)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"render_rays fwd: {ms:.3f} ms/step, {R / ms * 1e3:.0f} rays/s, alg {R*1.1376e9/ms/1e9:.1f} TFLOP/s")
    x = torch.rand(R * 128, 3, device="cuda") - 0.5; t = torch.rand(R * 128, device="cuda"); d = torch.randn(R*128, 3, device="cuda")
    for name, fn in [("sdf_query", lambda: r.sdf_from_observed_space(x, t)), ("point_forward", lambda: r.point_forward(x, d, t))]:
        fn(); torch.cuda.synchronize(); e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} on {R*128} pts: {e0.elapsed_time(e1)/5:.3f} ms")
r.sync_check()

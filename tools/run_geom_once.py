"""Launch each fused chain a few times (for ncu captures): python tools/run_geom_once.py [points]"""
import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from endosurf_b200 import EndoSurfRenderer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 32 * 8
torch.manual_seed(0)
r = EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), bench.NET_CFG, device="cuda"); bench.seeded_state(r.model); r.eval()
x = (torch.rand(n, 3, device="cuda") - 0.5) * 1.2; t = torch.rand(n, device="cuda"); d = torch.randn(n, 3, device="cuda")
with torch.no_grad():
    for _ in range(3):
        r.point_forward(x, d, t)
        r.sdf_from_observed_space(x, t)
torch.cuda.synchronize(); r.sync_check(); print("ok")

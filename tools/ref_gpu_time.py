"""Stock-PyTorch reference path ON THE GPU (oracle port on device='cuda' = what the unmodified reference runs as on a
B200, SURVEY 8d 'B1'): python tools/ref_gpu_time.py [rays] [mode]    -- context for the >=10x target, not a product path"""
import copy, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import endosurf_oracle as orc
from endosurf_b200 import EndoSurfNet
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
torch.manual_seed(0)
model = EndoSurfNet(bench.NET_CFG); bench.seeded_state(model)
train = mode == "train"
ck = {k: {kk: vv.detach().cuda().requires_grad_(train) for kk, vv in sd.items()} for k, sd in model.save_checkpoint().items()}
net = orc.OracleNet(ck, bench.NET_CFG); rc = copy.deepcopy(bench.RENDER_CFG)
params = [p for sd in ck.values() for p in sd.values()]
opt = torch.optim.Adam(params, lr=5e-4) if train else None
rays = bench.make_rays(R, 3).cuda(); cgt, dgt = (x.cuda() for x in bench.make_targets(R, 3))
def step():
    if train:
        opt.zero_grad(); o = orc.render_rays(net, rc, rays, iter_step=50000); bench.train_loss(o, cgt, dgt).backward(); opt.step()
    else:
        with torch.no_grad(): orc.render_rays(net, rc, rays, iter_step=50000)
for _ in range(2): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 3
for _ in range(n): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print(f"reference algorithm, stock PyTorch {torch.__version__} fp32 on {torch.cuda.get_device_name()}: {mode} {R} rays: "
      f"{dt*1e3:.1f} ms/step = {R/dt:.0f} rays/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")

"""torch.profiler breakdown of one training step (GPU box): python tools/profile_train_step.py [rays]"""
import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from endosurf_b200 import EndoSurfRenderer
from endosurf_b200 import training
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
r = EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), bench.NET_CFG, device="cuda"); bench.seeded_state(r.model); r.train()
params = [p for v in r.get_train_params().values() for p in v]
opt = torch.optim.Adam(params, lr=5e-4)
rays = bench.make_rays(R, 3).cuda(); cgt, dgt = (x.cuda() for x in bench.make_targets(R, 3))
def step():
    opt.zero_grad(set_to_none=True)
    o = r(rays, iter_step=50000); loss = bench.train_loss(o, cgt, dgt); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
print("mm out_dtype fast path:", training._MM_OUT_DTYPE_OK)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=90))
if len(sys.argv) > 2:  # per input shape
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
        step(); torch.cuda.synchronize()
    print(prof2.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=50, max_name_column_width=40))

"""GPU box: accuracy of the weight-gradient GEMM variants on a real training batch.
Reference = exact fp32 products of the same planes; compares 3-term, 2-term and 1-term fp16 tensor-core GEMMs."""
import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from endosurf_b200 import EndoSurfRenderer, training
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
r = EndoSurfRenderer(copy.deepcopy(bench.RENDER_CFG), bench.NET_CFG, device="cuda"); bench.seeded_state(r.model); r.train()
rays = bench.make_rays(R, 3).cuda(); cgt, dgt = (x.cuda() for x in bench.make_targets(R, 3))
def grads(terms, fast):
    training.WGRAD_TERMS = terms
    training.FAST_MIN_ROWS = 32768 if fast else 1 << 62
    r.zero_grad()
    torch.manual_seed(1)
    o = r(rays, iter_step=50000); loss = bench.train_loss(o, cgt, dgt); loss.backward()
    return {n: p.grad.detach().clone() for n, p in r.model.named_parameters() if p.grad is not None}
ref = grads(3, False)
for terms in (3, 2, 1):
    g = grads(terms, True)
    num = sum(((g[k] - ref[k]) ** 2).sum() for k in ref).sqrt().item(); den = sum((ref[k] ** 2).sum() for k in ref).sqrt().item()
    worst = sorted(((((g[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)).item(), k) for k in ref), reverse=True)[:4]
    print(f"terms={terms}: global rel err {num/den:.3e}; worst tensors {[(f'{e:.2e}', k) for e, k in worst]}")

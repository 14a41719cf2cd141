"""GPU box: gradients of the reference trainer's total loss after ONE train_step, reference renderer vs this renderer
(same checkpoint, frame, rays, jitter): per-parameter relative deviation, grouped by network."""
import importlib, os, sys, pathlib, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_trainer as T
tmp = pathlib.Path("/tmp/diag_trainer")
ref, _ = T._trainer(tmp / "ref", reference_renderer=True)
ours, te = T._trainer(tmp / "ours")
ours.renderer.load_checkpoint(ref.renderer.save_checkpoint())
grads = {}
for name, tr in (("ref", ref), ("ours", ours)):
    tr.writer = T._Recorder(); tr.renderer.train()
    torch.manual_seed(7); np.random.seed(7)
    tr.train_step(global_step=int(os.environ.get("STEP", "3000")))
    grads[name] = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in tr.renderer.named_parameters()}
    print(name, {k: round(v, 6) for k, v in tr.writer.scalars.items() if "loss" in k})
tot_d = tot_r = 0.0
for k, g in grads["ref"].items():
    o = grads["ours"].get(k)
    if g is None or o is None:
        print(f"{k:50s} ref {'None' if g is None else 'ok'} ours {'None' if o is None else 'ok'}"); continue
    d = (g - o).norm().item(); r = g.norm().item()
    tot_d += d * d; tot_r += r * r
    flag = " <<<" if d > 1e-2 * max(r, 1e-12) else ""
    print(f"{k:50s} |ref| {r:.3e} |ours| {o.norm().item():.3e} rel {d / max(r, 1e-30):.2e}{flag}")
print("whole gradient rel", (tot_d / tot_r) ** 0.5)

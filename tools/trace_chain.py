"""Pipeline trace of one CTA of a fused chain (GPU box): python tools/trace_chain.py [sdfq|geom|color]
Rebuilds the library with -DES_TRACE (the trace hooks are compiled out of the product build) and restores it after."""
import copy, ctypes as C, os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("ES_TRACE_CHILD") != "1":
    env = dict(os.environ, ES_NVCC_FLAGS="-DES_TRACE", ES_TRACE_CHILD="1")
    subprocess.check_call([sys.executable, "-m", "endosurf_b200.build", "--force"], cwd=ROOT, env=env, stdout=subprocess.DEVNULL)
    try:
        subprocess.check_call([sys.executable] + sys.argv, cwd=ROOT, env=env)
    finally:
        subprocess.check_call([sys.executable, "-m", "endosurf_b200.build", "--force"], cwd=ROOT, stdout=subprocess.DEVNULL)
    sys.exit(0)
import torch
from conftest import load_cfg, load_ckpt
from endosurf_b200 import EndoSurfRenderer, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "geom"
cfg = load_cfg(); r = EndoSurfRenderer(cfg["render"], cfg["net"], device="cuda"); r.load_checkpoint(load_ckpt()); r.eval()
n = 148 * 128 * 4
x = torch.rand(n, 3, device="cuda") - 0.5; t = torch.rand(n, device="cuda"); d = torch.randn(n, 3, device="cuda")
lib, ctx = _lib.load(), r._context()
fn = {"sdfq": lambda: r.sdf_from_observed_space(x, t), "geom": lambda: r.point_forward(x, None, t),
      "color": lambda: r.point_forward(x, d, t)}[which]  # "color": both chains run, the colour chain's trace is the one kept
fn(); torch.cuda.synchronize()
lib.es_debug_trace(ctx, None, 0)
fn(); torch.cuda.synchronize()
buf = np.zeros(2 + 2 * 8000, dtype=np.int64)
lib.es_debug_trace(ctx, buf.ctypes.data_as(C.c_void_p), 8000)
n0, n1 = int(min(buf[0], 4000)), int(min(buf[1], 4000))
ev = np.concatenate([buf[2:2 + 2 * n0].reshape(-1, 2), buf[2 + 8000:2 + 8000 + 2 * n1].reshape(-1, 2)])
cnt = n0 + n1
ev = ev[np.argsort(ev[:, 0])]; t0 = ev[0, 0]
names = {1: "MMA layer start", 2: "MMA chunk ready", 3: "MMA layer issued", 4: "EPI acc ready", 5: "EPI chunk written",
         6: "EPI  slot free", 7: "EPI  values ready", 8: "EPI  stores issued", 9: "EPI  fence done"}
print("events", cnt)
last = {}
for clk, code in ev[:400]:
    k = code // 1000; rest = code % 1000
    what = f"L{rest}" if k in (1, 3, 4) else f"L{rest // 16} c{rest % 16}"
    print(f"{clk - t0:9d}  {names[k]:18s} {what}")

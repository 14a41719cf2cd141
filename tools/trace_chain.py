"""Pipeline trace of one CTA of a fused chain (GPU box): python tools/trace_chain.py [sdfq|geom|color]
Rebuilds the library with -DES_TRACE (the trace hooks are compiled out of the product build) and restores it after."""
import copy, ctypes as C, os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
TRACE_LIB = os.path.join(ROOT, "endosurf_b200", "libendosurf_b200_trace.so")
if os.environ.get("ES_TRACE_CHILD") != "1":
    # the instrumented variant is a separate library (prebuilt in the build container, or built here once)
    env = dict(os.environ, ES_NVCC_FLAGS="-DES_TRACE", ES_TRACE_CHILD="1", ES_LIB_PATH=TRACE_LIB)
    if not os.path.exists(TRACE_LIB):
        subprocess.check_call([sys.executable, "-m", "endosurf_b200.build", "--tag=trace"], cwd=ROOT, env=env,
                              stdout=subprocess.DEVNULL)
    subprocess.check_call([sys.executable] + sys.argv, cwd=ROOT, env=env)
    sys.exit(0)
import torch
from conftest import load_cfg, load_ckpt
from endosurf_b200 import EndoSurfRenderer, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "geom"
cfg = load_cfg(); r = EndoSurfRenderer(cfg["render"], cfg["net"], device="cuda"); r.load_checkpoint(load_ckpt()); r.eval()
n = 148 * 128 * 4
x = torch.rand(n, 3, device="cuda") - 0.5; t = torch.rand(n, device="cuda"); d = torch.randn(n, 3, device="cuda")
lib, ctx = _lib.load(), r._context()
def train_fn(kind):
    # one training forward + backward over the points; only launches of `kind` (es_api.cu K_*) write the trace
    lib.es_debug_set(ctx, 6, kind)
    r.train()
    def f():
        sdf, g_c, jac, rgb = r.point_field(x, d, t.reshape(-1, 1))
        (sdf.sum() + g_c.sum() + jac.sum() + rgb.sum()).backward()
    return f
fn = {"sdfq": lambda: r.sdf_from_observed_space(x, t), "geom": lambda: r.point_forward(x, None, t),
      "color": lambda: r.point_forward(x, d, t),
      "geom_train": 0, "color_train": 1, "rev_deform": 3, "rev_sdf": 4, "rev_color": 5}[which]
if isinstance(fn, int):
    fn = train_fn(fn)  # "color": both chains run, the colour chain's trace is the one kept
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
         6: "EPI  slot free", 7: "EPI  values ready", 8: "EPI  tail done", 9: "EPI  tail stored"}
print("events", cnt)
last = {}
for clk, code in ev[:int(os.environ.get('ES_TRACE_EVENTS', '400'))]:
    k = code // 1000; rest = code % 1000
    what = f"L{rest}" if k in (1, 3, 4, 8, 9) else f"L{rest // 16} c{rest % 16}"
    print(f"{clk - t0:9d}  {names[k]:18s} {what}")

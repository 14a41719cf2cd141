"""Run the tcgen05 layout probe standalone (GPU box): python tools/probe.py"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endosurf_b200 import _lib
lib = _lib.load()
cfg = _lib.EsNetConfig(1, 9, 4, 256, 6, 6, 6, 10, 4, 3)
ctx = C.c_void_p()
assert lib.es_create(C.byref(ctx), C.byref(cfg)) == 0
g = torch.Generator().manual_seed(0)
a = torch.randn(128, 64, generator=g).to(torch.float16).cuda()
b = torch.randn(256, 64, generator=g).to(torch.float16).cuda()
ref = a.float() @ b.float().t()
for name, v in {"default": (0, 0, 0, 0)}.items():
    d = torch.zeros(128, 256, device="cuda")
    rc = lib.es_umma_probe(ctx, a.data_ptr(), b.data_ptr(), d.data_ptr(), *v, None)
    torch.cuda.synchronize()
    e = ((d - ref).abs().max() / ref.abs().max()).item()
    print(name, "rc", rc, "relerr", e, "check", lib.es_sync_check(ctx, None), flush=True)

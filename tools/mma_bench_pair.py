"""tcgen05 cta_group::2 MMA rate (GPU box): ES_MMAB_PAIR=1 python tools/mma_bench_pair.py"""
import ctypes as C, os, sys, numpy as np
os.environ["ES_MMAB_PAIR"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endosurf_b200 import _lib
lib = _lib.load()
cfg = _lib.EsNetConfig(1, 9, 4, 256, 6, 6, 6, 10, 4, 3)
ctx = C.c_void_p(); assert lib.es_create(C.byref(ctx), C.byref(cfg)) == 0
def run(name, n, a, b, grid=148, iters=256):
    arr = (C.c_int32 * 15)(n, iters, 4, *a, *b)
    out = (C.c_int64 * grid)()
    rc = lib.es_mma_bench(ctx, arr, grid, out)
    cyc = np.array(list(out), dtype=np.float64)[0::2] / (iters * 4)
    print(f"{name:64s} N={n:3d} grid={grid:3d} rc={rc} cycles/MMA(M=256): median {np.median(cyc):7.1f} min {cyc.min():7.1f} max {cyc.max():7.1f}", flush=True)
for grid in (2, 148):
    # A: [kgroup][128 rows][16B] per CTA; B per CTA: [kgroup][128 rows][16B] (N/2 = 128 rows), K=16 step = 2 k-groups
    run("A no-swizzle / B half units no-swizzle (chain kernel, pair mode)", 256, (0, 2048, 128, 4096, 4, 16384), (0, 2048, 128, 4096, 4, 8192), grid)
    run("same, N=128 (64 rows of B per CTA)", 128, (0, 2048, 128, 4096, 4, 16384), (0, 1024, 128, 2048, 4, 4096), grid)

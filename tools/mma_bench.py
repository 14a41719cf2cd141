"""tcgen05 MMA issue-rate vs shared-memory operand layout (GPU box): python tools/mma_bench.py"""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from endosurf_b200 import _lib
lib = _lib.load()
cfg = _lib.EsNetConfig(1, 9, 4, 256, 6, 6, 6, 10, 4, 3)
ctx = C.c_void_p(); assert lib.es_create(C.byref(ctx), C.byref(cfg)) == 0
def run(name, n, a, b, grid=148, iters=256, ksteps=4):
    # a/b = (layout, lbo, sbo, kadv, tiles, tile_bytes)
    arr = (C.c_int32 * 15)(n, iters, ksteps, *a, *b)
    out = (C.c_int64 * grid)()
    rc = lib.es_mma_bench(ctx, arr, grid, out)
    cyc = np.array(list(out), dtype=np.float64) / (iters * ksteps)
    print(f"{name:58s} N={n:3d} grid={grid:3d} rc={rc} cycles/MMA: median {np.median(cyc):7.1f} min {cyc.min():7.1f} max {cyc.max():7.1f}", flush=True)
NOSW_A = (0, 2048, 128, 4096, 4, 16384)      # current A: [kgroup][128 rows][16B]
NOSW_B = (0, 4096, 128, 8192, 4, 16384)      # current B unit: [kgroup][256 rows][16B], K=32 per 16 KiB unit (2 ksteps)
SW128_A = (2, 16, 1024, 32, 4, 16384)        # 128 rows x 128 B swizzled, +32 B per K step
SW128_B = (2, 16, 1024, 32, 4, 32768)        # 256 rows x 128 B
SW64_B = (4, 16, 512, 32, 4, 16384)          # 256 rows x 64 B (K=32 units): only 2 ksteps per tile
for grid in (1, 148):
    run("A no-swizzle / B no-swizzle (current), K=64 tiles", 256, NOSW_A, (0, 4096, 128, 8192, 2, 32768), grid)
    run("A no-swizzle / B no-swizzle, 2 ksteps per B unit (as kernel)", 256, (0, 2048, 128, 4096, 8, 8192), NOSW_B, grid, ksteps=2)
    run("A SW128 / B SW128", 256, SW128_A, SW128_B, grid)
    run("A SW128 / B SW64 (2 ksteps)", 256, (2, 16, 1024, 32, 8, 8192), SW64_B, grid, ksteps=2)
    run("A SW128 / B no-swizzle", 256, SW128_A, (0, 4096, 128, 8192, 2, 32768), grid)
    run("A no-swizzle / B SW128", 256, NOSW_A, SW128_B, grid)
    run("A no-swizzle / B no-swizzle  N=128", 128, NOSW_A, (0, 4096, 128, 8192, 2, 32768), grid)
    run("A SW128 / B SW128  N=128", 128, SW128_A, SW128_B, grid)
    run("A SW128 / B SW128  N=64", 64, SW128_A, SW128_B, grid)

"""tcgen05 MMA rate under background shared-memory traffic (GPU box): python tools/mma_noise.py
Each line: ES_MMAB_NOISE = (noise warps, 128-byte stores per burst, idle cycles between bursts, weight-stream TMA on/off).
The fused MLP kernel's epilogue issues 16 warps x 16 stores per 64-wide K chunk (1536 tensor cycles) and streams 64 KiB of
weights per chunk."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    sys.path.insert(0, ROOT)
    from endosurf_b200 import _lib
    lib = _lib.load()
    cfg = _lib.EsNetConfig(1, 9, 4, 256, 6, 6, 6, 10, 4, 3)
    ctx = C.c_void_p(); assert lib.es_create(C.byref(ctx), C.byref(cfg)) == 0
    grid, iters, ksteps = 148, 256, 2
    arr = (C.c_int32 * 15)(256, iters, ksteps, 0, 2048, 128, 4096, 8, 8192, 0, 4096, 128, 8192, 4, 16384)
    out = (C.c_int64 * grid)()
    rc = lib.es_mma_bench(ctx, arr, grid, out)
    cyc = np.array(list(out), dtype=np.float64) / (iters * ksteps)
    print(f"noise={os.environ.get('ES_MMAB_NOISE', 'none'):16s} commit_every={os.environ.get('ES_MMAB_COMMIT', '0'):3s} ldtm={os.environ.get('ES_MMAB_LDTM', '0'):3s} rc={rc} cycles/MMA median {np.median(cyc):7.1f} min {cyc.min():7.1f} max {cyc.max():7.1f}", flush=True)
else:
    for nz, cm, ld in [("0,0,0,0", 0, 0), ("16,16,600,1", 0, 0), ("16,64,0,1", 0, 0),
                       ("0,0,0,0", 4, 0), ("0,0,0,0", 2, 0), ("0,0,0,0", 1, 0), ("16,16,600,1", 2, 0),
                       ("16,0,600,0", 0, 1), ("16,0,100,0", 0, 4), ("16,0,0,0", 0, 16), ("16,16,600,1", 2, 2)]:
        subprocess.call([sys.executable, __file__, "child"],
                        env=dict(os.environ, ES_MMAB_NOISE=nz, ES_MMAB_COMMIT=str(cm), ES_MMAB_LDTM=str(ld)))

"""Build libendosurf_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI).

    python -m endosurf_b200.build            # incremental
    python -m endosurf_b200.build --force
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libendosurf_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# per-file extra flags: the per-ray kernels round every op like separate PyTorch elementwise kernels do
SOURCES = {
    "es_mlp.cu": [],
    "es_probe.cu": [],
    "es_mmabench.cu": [],
    "es_pack.cu": [],
    "es_wgrad.cu": [],
    "es_rays.cu": ["-fmad=false"],
    "es_api.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the endosurf_b200 CUDA library cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, tag: str = "") -> str:
    """tag: build a variant (ES_NVCC_FLAGS, e.g. -DES_TRACE) into libendosurf_b200_<tag>.so with its own objects, next
    to the product library (profiling tools load it through ES_LIB_PATH)."""
    nvcc = _nvcc()
    lib = LIB if not tag else LIB.replace(".so", f"_{tag}.so")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "endosurf_b200.h"))
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", (f".{tag}" if tag else "") + ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, *os.environ.get("ES_NVCC_FLAGS", "").split(), "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{log}")
            if verbose:
                print(log)
    if force or _stale(lib, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", lib, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return lib


if __name__ == "__main__":
    tags = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tag=tags[0] if tags else ""))

"""Recipe that carries the UNMODIFIED reference to the GPU box (test / baseline infrastructure, never the product).

The reference (Ruyi-Zha/endosurf) is pure Python; its "build" is byte-compilation.  This script compiles the package
``src/`` of ``/root/reference`` where it lies into ONE zip of sourceless ``.pyc`` files under ``oracle/_ref/`` (git-ignored, but
not gpurun-ignored: it travels to the GPU box like our own built ``.so``).  No reference source is copied into the
repository.  ``oracle/ref_shims.py`` makes the result importable (``src.renderer.endosurf`` etc.) and supplies
stand-ins for the third-party packages this image lacks.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box: uses the prebuilt files)
"""
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ENDOSURF_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ZIP = os.path.join(OUT, "reference_pyc.zip")


def build(verbose=True) -> bool:
    src_root = os.path.join(REF, "src")
    if not os.path.isdir(src_root):
        if verbose:
            print(f"[oracle/_ref] {REF} not present: keeping whatever is already built")
        return os.path.isfile(ZIP)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    n = 0
    # one binary artefact: a zip of sourceless .pyc files (zipimport loads them; a single file also travels to the GPU
    # box whatever the snapshot tool thinks of *.pyc)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ZIP, "w", zipfile.ZIP_DEFLATED) as zf:
        for dirpath, _, files in os.walk(src_root):
            rel = os.path.relpath(dirpath, REF)
            zf.writestr(rel + "/", "")  # explicit directory entry: zipimport then treats it as a (namespace) package
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                cfile = os.path.join(tmp, f"{n}.pyc")
                py_compile.compile(os.path.join(dirpath, f), cfile=cfile, doraise=True,
                                   dfile=os.path.join("reference", rel, f))
                zf.write(cfile, os.path.join(rel, f + "c"))
                n += 1
    # the one config the benchmark / trainer test uses, as data (YAML -> JSON), so that nothing reads /root/reference
    # at run time
    try:
        import json
        import yaml
        cfg_path = os.path.join(REF, "configs", "endosurf", "baseline", "base_pull.yml")
        with open(cfg_path) as fh:
            cfg = yaml.safe_load(fh)
        with open(os.path.join(OUT, "base_pull.json"), "w") as fh:
            json.dump(cfg, fh, indent=1)
    except Exception as e:  # the config is optional (tests carry their own copy under tests/golden)
        if verbose:
            print(f"[oracle/_ref] config not exported: {e}")
    with open(os.path.join(OUT, "BUILT_FROM.txt"), "w") as fh:
        fh.write(f"byte-compiled from {REF}/src by oracle/build_ref.py with python {sys.version.split()[0]}; "
                 f"{n} modules, no sources\n")
    if verbose:
        print(f"[oracle/_ref] {n} modules byte-compiled into {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)

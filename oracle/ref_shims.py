"""Import shims for the byte-compiled reference under ``oracle/_ref`` (test / baseline infrastructure only).

``load_reference()`` returns the reference's own ``src.renderer.endosurf`` module.  The third-party packages the
reference imports at module scope but this image lacks (``mcubes, kornia, lpips, open3d, imageio, trimesh``; none is
touched by ``render_rays``) are replaced by inert stand-ins in ``sys.modules`` first."""
import importlib
import os
import sys
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_ZIP = os.path.join(REF_DIR, "reference_pyc.zip")
MISSING = ["mcubes", "kornia", "lpips", "open3d", "imageio", "imageio.v2", "trimesh", "wandb"]


def available() -> bool:
    return os.path.isfile(REF_ZIP)


def install_shims():
    for name in MISSING:
        try:
            importlib.import_module(name)
        except Exception:
            m = mock.MagicMock(name=name)
            m.__spec__ = None
            sys.modules[name] = m
            if "." in name:
                setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
    if REF_ZIP not in sys.path:
        sys.path.insert(0, REF_ZIP)  # zipimport: src/renderer/endosurf.pyc -> src.renderer.endosurf


def load_reference():
    """-> the reference's src.renderer.endosurf module (EndoSurfRenderer, EndoSurfNet, ...)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists")
    install_shims()
    return importlib.import_module("src.renderer.endosurf")

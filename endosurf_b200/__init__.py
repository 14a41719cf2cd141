"""endosurf_b200: B200-native (sm_100a) implementation of EndoSurf's per-ray volume-rendering hot path.

Public surface mirrors the reference (``src/renderer/endosurf.py``): :class:`EndoSurfRenderer`.
"""
from .renderer import EndoSurfRenderer, EndoSurfNet  # noqa: F401

__all__ = ["EndoSurfRenderer", "EndoSurfNet"]

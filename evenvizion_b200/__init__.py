"""evenvizion_b200: B200-native (sm_100a) implementation of EvenVizion's frame-to-frame
geometry hot path (descriptor matching -> RANSAC homography -> None-H / superposition /
coordinate remap), drop-in behind the reference's Python API.

    from evenvizion_b200.processing import KeyPoints, compute_homography, superposition_dict, ...
    from evenvizion_b200 import GeometryEngine            # batched device API

Importing the package does not need a GPU; creating a GeometryEngine (or calling any
processing function) does, and fails loudly otherwise -- there is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import EvzError, LIB_PATH          # noqa: F401
from .engine import GeometryEngine, FrameStore, PairResults   # noqa: F401

_default_engine = None


def default_engine(device=None):
    """Process-wide engine on the current CUDA device (created on first use)."""
    global _default_engine
    if _default_engine is None:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("evenvizion_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _default_engine = GeometryEngine(torch.cuda.current_device() if device is None else device)
    return _default_engine

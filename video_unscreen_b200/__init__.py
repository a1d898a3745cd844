"""B200-native (sm_100a) implementation of video_unscreen's per-pixel,
frame-parallel matte hot path, behind the reference's own Python interface.

    video_unscreen_b200.unscreen   drop-in mirror of the reference's package
    video_unscreen_b200.ops        device-level operators (CUDA tensors)
    video_unscreen_b200.clip       batched, device-resident clip pipelines
    include/vu_b200.h              the C ABI all of the above call
"""
import sys

__version__ = "0.1.0"


def install(name="unscreen"):
    """alias the mirror package as ``unscreen`` (and its sub-modules) so that
    ``from unscreen.colorfiltering import ColorFilteringAgent`` etc. resolve to
    the B200 implementation.  Call before importing the reference's tools."""
    import importlib
    base = __name__ + ".unscreen"
    for sub in ("", ".utils", ".colorfiltering", ".trimap", ".bgmodel"):
        sys.modules[name + sub] = importlib.import_module(base + sub)
    return sys.modules[name]

"""B200-native (sm_100a) implementation of video_unscreen's per-pixel,
frame-parallel matte hot path, behind the reference's own Python interface.

    video_unscreen_b200.unscreen   drop-in mirror of the reference's package
    video_unscreen_b200.ops        device-level operators (CUDA tensors)
    video_unscreen_b200.clip       batched, device-resident clip pipelines
    include/vu_b200.h              the C ABI all of the above call
"""
import importlib
import sys

__version__ = "0.2.0"

# reference module -> mirror module whose ``__all__`` is patched over it.  Only hot-path names are replaced: everything
# else of the reference package (parallel_read_img, save_img, save_video, get_center, return_date,
# unscreen.binseg / stm / vmatting / iseg / harmonization ...) stays the reference's own.
_OVERLAY = {
    "utils.fgfuncs": "utils.fgfuncs",
    "utils.maskprocess": "utils.maskprocess",
    "utils.imgprocess": "utils.imgprocess",
    "utils.visualize": "utils.visualize",
    "utils.region_fill": "utils.region_fill",
    "utils": "utils.temporal",            # no reference counterpart: the scripts' inline lines + the temporal median
    "colorfiltering.agent": "colorfiltering.agent",
    "trimap.agent": "trimap.agent",
    "bgmodel.agent": "bgmodel.agent",
}
_AGENT_CLASSES = {"colorfiltering.agent": "ColorFilteringAgent", "trimap.agent": "TrimapAgent", "bgmodel.agent": "BackgroundAgent"}


def _reference_importable(name):
    """True when a package called ``name`` that is NOT this mirror can be imported (the reference checkout on sys.path)."""
    if name in sys.modules:
        return not getattr(sys.modules[name], "__name__", "").startswith(__name__)
    try:
        spec = importlib.util.find_spec(name)
    except (ImportError, ValueError):
        return False
    return spec is not None


def install(name="unscreen", overlay=None, io=False):
    """Put the B200 hot path behind the import name ``unscreen``.  Call before importing the reference's tools.

    * The reference package is importable (its checkout is on ``sys.path``) -> OVERLAY: the reference is imported as it
      is and the hot-path names of this mirror are patched over the reference's in ``unscreen.utils`` (and the
      sub-modules that define them, so the reference's internal callers pick them up too), ``unscreen.colorfiltering``
      ``unscreen.trimap`` and ``unscreen.bgmodel``.  Everything else -- file I/O, the CNN agents -- stays the
      reference's, so ``tools/unscreen/green.py``, ``bg.py``, ``bg_offline.py`` and ``tools/replace/replace.py``
      import and run unchanged.
    * No reference package around -> ALIAS: ``sys.modules['unscreen'...]`` point at the mirror (hot-path names only).

    ``io=True`` also replaces ``parallel_read_img`` / ``save_img`` (utils/fileio.py) with the nvJPEG-based ones: JPEG
    decoders are not bit-identical, so that one is opt-in.  ``overlay`` forces one mode (True / False).  Returns
    ``sys.modules[name]``; ``installed_names(name)`` lists what was patched."""
    import importlib.util  # noqa: F401  (find_spec)
    base = __name__ + ".unscreen"
    if overlay is None:
        overlay = _reference_importable(name)
    if not overlay:
        for sub in ("", ".utils", ".colorfiltering", ".trimap", ".bgmodel"):
            sys.modules[name + sub] = importlib.import_module(base + sub)
        _PATCHED[name] = ["*"]
        return sys.modules[name]

    ref = importlib.import_module(name)
    if getattr(ref, "__name__", "").startswith(__name__):
        raise ImportError(f"install(overlay=True): '{name}' resolves to the B200 mirror itself, not to the reference package")
    patched = []
    overlay_map = dict(_OVERLAY, **({"utils.fileio": "utils.fileio"} if io else {}))
    for ref_sub, mirror_sub in overlay_map.items():
        mirror = importlib.import_module(f"{base}.{mirror_sub}")
        target = importlib.import_module(f"{name}.{ref_sub}")
        names = getattr(mirror, "__all__", None) or [_AGENT_CLASSES[ref_sub]]
        # the defining sub-module, and the package that star-exports it (unscreen.utils / unscreen.trimap ...)
        parent = importlib.import_module(f"{name}.{ref_sub.split('.')[0]}")
        for n in names:
            obj = getattr(mirror, n)
            setattr(target, n, obj)
            setattr(parent, n, obj)
            patched.append(f"{name}.{ref_sub}.{n}")
    _PATCHED[name] = patched
    return ref


_PATCHED = {}


def installed_names(name="unscreen"):
    """what the last ``install(name)`` replaced (['*'] for the alias mode)"""
    return list(_PATCHED.get(name, []))

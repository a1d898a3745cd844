"""reference: unscreen/utils/visualize.py (only the blend is on the hot path)."""
from ... import _lib, ops
from ..._io import back, to_dev

__all__ = ["fuse_fgbg"]


def fuse_fgbg(fg, bg, mask):
    """reference visualize.py:7-24: u8(a*fg + (1-a)*bg), a = mask/255 in float64."""
    f, as_np = to_dev(fg)
    b, _ = to_dev(bg)
    m, _ = to_dev(mask)
    return back(ops.blend(_lib.BLEND_FUSE, f, m, b), as_np)

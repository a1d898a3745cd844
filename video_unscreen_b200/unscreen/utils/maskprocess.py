"""mask processing (reference: unscreen/utils/maskprocess.py)."""
import torch

from ... import _lib, ops
from ..._io import back, to_dev

__all__ = ["dilate_mask", "erode_mask", "exist_foreground", "get_outer_boundary"]


def _morph(mask, kernelsize, iters, op):
    t, as_np = to_dev(mask)
    if t.ndim == 3 and t.shape[-1] == 3:
        # the reference also feeds 3-channel masks (bg_offline.py:116): per-channel morphology
        planes = t.permute(2, 0, 1).contiguous()
        out = ops.morph(planes, kernelsize, iters, op).permute(1, 2, 0).contiguous()
    else:
        out = ops.morph(t, kernelsize, iters, op)
    return back(out, as_np)


def dilate_mask(mask, kernelsize=5, iters=10):
    """reference maskprocess.py:7-19."""
    return _morph(mask, kernelsize, iters, _lib.DILATE)


def erode_mask(mask, kernelsize=5, iters=10):
    """reference maskprocess.py:22-34."""
    return _morph(mask, kernelsize, iters, _lib.ERODE)


def exist_foreground(mask, fg_exist_thr):
    """reference maskprocess.py:56-60: count(mask >= 128) > thr*h*w (strict)."""
    t, _ = to_dev(mask)
    h, w = t.shape
    n = int(ops.count_cmp(t, _lib.CMP_GE, 128).item())
    return bool(n > fg_exist_thr * h * w)


def get_outer_boundary(mask, kernelsize=7, iters=10):
    """reference maskprocess.py:63-74 (uint8 subtraction wraps; the clip is a no-op)."""
    t, as_np = to_dev(mask)
    d = ops.dilate(t, kernelsize, iters)
    return back(ops.sub_wrap(d, t), as_np)

"""regionfill on the device (reference: unscreen/utils/region_fill.py:7-17; callers bgmodel/agent.py:150, bg.py:79).

Same signature and return type as the reference (float64 [H,W]); the sparse direct solve is replaced by conjugate gradients
on the pixel grid (csrc/vu_regionfill.cu), so parity is a tolerance: within ~1e-6 of the reference's float result, i.e. the
truncated uint8 images agree except where the exact value sits on an integer (a flat boundary: the reference's own
round-off decides there)."""
import numpy as np
import torch

from ... import ops
from ..._io import device

__all__ = ["regionfill"]


def _mask_dev(mask):
    if isinstance(mask, torch.Tensor):
        return (mask != 0).to(device(), torch.uint8)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(mask) != 0).astype(np.uint8)).to(device())


def regionfill(I, mask, factor=1.0):
    """I: [H,W] image plane (any real dtype), mask: [H,W], != 0 where to fill -> float64 [H,W] (numpy in, numpy out; CUDA
    tensors in, CUDA tensor out).  ``I`` may also be [C,H,W] CUDA planes sharing the mask (one solve for the B, G, R planes
    of bg.py:79)."""
    as_np = not isinstance(I, torch.Tensor)
    m = _mask_dev(mask)
    if as_np:
        if not np.asarray(mask).any():
            return np.asarray(I).copy()              # region_fill.py:8-9 returns a copy of the input, in its own dtype
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(I), dtype=np.float64)).to(device())
    else:
        t = I.to(device())
    out = ops.regionfill(t[None] if t.ndim == 2 else t, m, float(factor))
    out = out[0] if t.ndim == 2 else out
    return out.cpu().numpy() if as_np else out

"""TrimapAgent on the B200 (reference: unscreen/trimap/agent.py)."""
import numpy as np
import torch

from ... import ops
from ..._io import back, to_dev
from ..utils.fgfuncs import _inrange_dev
from ..utils.imgprocess import get_target_size


class TrimapAgent():
    """same constructor and methods as the reference (trimap/agent.py:25-33)."""

    def __init__(self, input_long_side=960, kernelsize=3, iters=5, color_winsize=(10, 100, 180)):
        self.kernelsize = kernelsize
        self.iters = iters
        self.input_long_side = input_long_side
        self.color_winsize = color_winsize

    def _trimap_dev(self, mask):
        ori_h, ori_w = mask.shape
        ih, iw = get_target_size(ori_h, ori_w, self.input_long_side)
        m = ops.resize_nearest_mask(mask, ih, iw)
        if self.kernelsize == 3 and self.iters <= ops.CROSS_MAX_PASSES:
            tri = ops.trimap_core(m, self.iters)
        else:
            tri = ops.trimap_classify(ops.dilate(m, self.kernelsize, self.iters), ops.erode(m, self.kernelsize, self.iters))
        # trimap/agent.py:59 passes INTER_NEAREST in the dst slot: the up-scale is bilinear
        tri = ops.resize_linear_mask(tri, ori_h, ori_w)
        return ops.trimap_snap(tri)

    def generate_trimap(self, mask):
        """trimap/agent.py:35-61."""
        m, as_np = to_dev(mask)
        return back(self._trimap_dev(m), as_np)

    def generate_trimap_withbg(self, mask, img, bgimg):
        """trimap/agent.py:63-101; ``bgimg`` is (h,w,3) or (3,)."""
        m, as_np = to_dev(mask)
        f, _ = to_dev(img)
        if bgimg.ndim == 1:
            bg = bgimg.cpu().numpy() if isinstance(bgimg, torch.Tensor) else np.asarray(bgimg)
        else:
            bg, _ = to_dev(bgimg)
        if m.ndim == 2 and f.ndim == 3 and m.numel() % 4 == 0:
            # the batched path: the ratio test of :92-96 is decided on the device (no host round trip between the kernels); an
            # empty mask comes back as it went in (all zeros)
            from ... import clip
            return back(clip.trimap_clip(m[None], self, f[None], bg, chunk=1, streams=1)[0], as_np)
        bgmask = _inrange_dev(f, bg, self.color_winsize)
        fuzzy_n, pos_n = (int(v) for v in ops.count_and(m, bgmask)[0].tolist())
        if pos_n == 0:
            return mask
        if float(fuzzy_n) / pos_n > 0.1:
            return back(self._trimap_dev(m), as_np)
        fuzzy = ops.mask_and01(m, bgmask)
        tri = self._trimap_dev(ops.mask_clear_where(m, fuzzy))
        return back(ops.mask_set128_where(tri, fuzzy), as_np)

    def forward(self, *args, **kwargs):
        """trimap/agent.py:103-128."""
        if len(args) > 2:
            return self.generate_trimap_withbg(*args, **kwargs)
        return self.generate_trimap(*args, **kwargs)

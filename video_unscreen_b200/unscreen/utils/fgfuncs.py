"""foreground / background extraction and compositing
(reference: unscreen/utils/fgfuncs.py)."""
import numpy as np
import torch

from ... import _lib, ops
from ..._io import back, to_dev
from .imgprocess import get_target_size

__all__ = ["is_pixel_inrange", "get_fg_naive", "get_fg", "get_bg", "get_fg_with_colorremove", "composite_fgbg"]


def bgr2hsv_pixel(bgr):
    """cv2 BGR2HSV of one uint8 pixel (integer fixed point); host-side
    because it only turns a 3-byte colour into range bounds."""
    b, g, r = (int(v) for v in bgr)
    v = max(b, g, r)
    d = v - min(b, g, r)
    sdiv = int(np.rint((255 << 12) / v)) if v else 0
    hdiv = int(np.rint((180 << 12) / (6.0 * d))) if d else 0
    s = (d * sdiv + 2048) >> 12
    h = (g - b) if v == r else ((b - r + 2 * d) if v == g else (r - g + 4 * d))
    h = (h * hdiv + 2048) >> 12
    if h < 0:
        h += 180
    return np.array([h, s, v], np.int64)


def _inrange_dev(img_t, bg_t_or_color, winsize):
    half = np.array(winsize) // 2
    if isinstance(bg_t_or_color, np.ndarray):  # (3,) colour
        hsv = bgr2hsv_pixel(bg_t_or_color)
        lo = np.clip(hsv - half, 10, 255)
        hi = np.clip(hsv + half, 10, 255)
        return ops.inrange_color(img_t, lo, hi)
    return ops.inrange_image(img_t, bg_t_or_color, half)


def is_pixel_inrange(img, bgimg, winsize=(20, 20, 120), long_side_input=-1):
    """reference fgfuncs.py:9-65.  ``bgimg`` is (h,w,3) or (3,); returns a bool mask."""
    assert bgimg.ndim == 3 or bgimg.ndim == 1
    t, as_np = to_dev(img)
    h, w = t.shape[:2]
    if bgimg.ndim == 1:
        bg = bgimg.cpu().numpy() if isinstance(bgimg, torch.Tensor) else np.asarray(bgimg)
    else:
        bg, _ = to_dev(bgimg)
    if long_side_input > 0:
        ih, iw = get_target_size(h, w, long_side_input)
        t = ops.resize_linear_image(t, ih, iw)
        if bgimg.ndim == 3:
            bg = ops.resize_linear_image(bg, ih, iw)
    m = _inrange_dev(t, bg, winsize)
    if long_side_input > 0:
        # fgfuncs.py:51,63: the INTER_NEAREST flag lands in the dst slot => bilinear up-scale, then > 0
        if bgimg.ndim == 1:
            m = ops.binarise(m, 0)  # cv2.inRange gives 0/255, the torch path 0/1
        m = ops.resize_linear_mask(m, h, w)
        m = ops.binarise(m, 0)
    out = back(m, as_np)
    return (out > 0) if as_np else (m > 0)


def get_fg_naive(img, alpha):
    """reference fgfuncs.py:68-81: u8(f64(img) * alpha/255)."""
    t, as_np = to_dev(img)
    a, _ = to_dev(alpha)
    return back(ops.blend(_lib.BLEND_NAIVE, t, a), as_np)


def get_fg(img, alpha, bg, patch=None):
    """reference fgfuncs.py:84-110.  ``patch`` ('lt128' | 'eq0') fuses the
    callers' ``bg[alpha<128] = img[alpha<128]`` (green.py:125) or
    ``bg[alpha==0] = img[alpha==0]`` (bg.py:99, bg_offline.py:171) into the
    same pass; the input ``bg`` is never mutated."""
    t, as_np = to_dev(img)
    a, _ = to_dev(alpha)
    b, _ = to_dev(bg)
    mode = {None: _lib.PATCH_NONE, "lt128": _lib.PATCH_ALPHA_LT128, "eq0": _lib.PATCH_ALPHA_EQ0}[patch]
    return back(ops.get_fg(t, a, b, mode), as_np)


def get_bg(alpha, bg):
    """reference fgfuncs.py:113-137."""
    a, as_np = to_dev(alpha)
    b, _ = to_dev(bg)
    return back(ops.get_bg(a, b), as_np)


def get_fg_with_colorremove(img, alpha, bg, winsize=(10, 100, 120), long_side_input=960):
    """reference fgfuncs.py:140-169 (without mutating ``alpha``)."""
    t, as_np = to_dev(img)
    a, _ = to_dev(alpha)
    b, _ = to_dev(bg)
    m = is_pixel_inrange(t, b, winsize, long_side_input).to(torch.uint8)
    a = ops.mask_clear_where(a, m)
    return back(ops.get_fg(t, a, b), as_np)


def composite_fgbg(fg, alpha, bg, extend=False):
    """reference fgfuncs.py:172-214."""
    f, as_np = to_dev(fg)
    a, _ = to_dev(alpha)
    b, _ = to_dev(bg)
    fg_h, fg_w = f.shape[:2]
    bg_h, bg_w = b.shape[:2]
    if float(fg_h) / fg_w > float(bg_h) / bg_w:
        new_bg_h = fg_h
        new_bg_w = int(float(bg_w) * new_bg_h / bg_h)
    else:
        new_bg_w = fg_w
        new_bg_h = int(float(bg_h) * new_bg_w / bg_w)
    b = ops.resize_linear_image(b, new_bg_h, new_bg_w)
    left = max(new_bg_w // 2 - fg_w // 2, 0)
    top = max(new_bg_h // 2 - fg_h // 2, 0)
    roi = b[top:top + fg_h, left:left + fg_w].contiguous()
    comp = ops.blend(_lib.BLEND_COMPOSITE, f, a, roi)
    if extend:
        out = b.clone()
        out[top:top + fg_h, left:left + fg_w] = comp
        comp = out
    return back(comp, as_np)

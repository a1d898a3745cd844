"""Temporal background estimators and the pipeline scripts' inline per-pixel
steps (no reference signature: these are the inline lines of
tools/unscreen/bg.py, bg_offline.py and tools/replace/replace.py, plus the
new-spec exact temporal median)."""
from ... import _lib, ops
from ..._io import back, to_dev

__all__ = ["temporal_median", "masked_temporal_mean", "fuse_bg", "bgdiff_gate", "binarise_dilate", "replace_blend"]


def temporal_median(frames):
    """exact per-pixel temporal median of frames[N,H,W,3] (or any [N,...] uint8)
    == np.median(frames, 0).astype(np.uint8).  SURVEY.md section 8 a23."""
    t, as_np = to_dev(frames)
    return back(ops.temporal_median(t), as_np)


def masked_temporal_mean(frames, masks, ksize=3, iters=2, min_count=10):
    """tools/unscreen/bg_offline.py:106-125: masks[N,H,W] are dilated
    (dilate_mask(mask, 3, 2)), frames averaged where un-masked.  Returns
    (bg_always[H,W,3], mask_always[H,W]); the TELEA inpaint of :127-129 is out of scope."""
    f, as_np = to_dev(frames)
    m, _ = to_dev(masks)
    bg, always = ops.masked_temporal_mean_raw(f, m, ksize, iters, min_count)
    return back(bg, as_np), back(always, as_np)


def fuse_bg(bgimg, bgimg_always, beta):
    """tools/unscreen/bg_offline.py:150-151."""
    b, as_np = to_dev(bgimg)
    a, _ = to_dev(bgimg_always)
    return back(ops.fuse_bg(b, a, beta), as_np)


def bgdiff_gate(frame, bgimg, mask, thr=25):
    """tools/unscreen/bg.py:85-92 == bg_offline.py:154-160."""
    f, as_np = to_dev(frame)
    b, _ = to_dev(bgimg)
    m, _ = to_dev(mask)
    if ops.bgdiff_gate_supported(f, b, m):
        return back(ops.bgdiff_gate(f, b, m, thr), as_np)
    g = ops.bgdiff_gray(f, b, thr)    # widths that are not a multiple of 4: the unfused kernels
    g = ops.dilate(g, 4, 2)
    return back(ops.gate(m, g), as_np)


def binarise_dilate(alpha):
    """tools/unscreen/bg.py:74-77."""
    a, as_np = to_dev(alpha)
    return back(ops.dilate(ops.binarise(a, 128), 3, 2), as_np)


def replace_blend(fg, mask, bg):
    """tools/replace/replace.py:74-76: u8(fg*m + bg*(1-m)), m = mask/255 in
    float64; mask is HWC (as decoded from JPEG) or HW."""
    f, as_np = to_dev(fg)
    m, _ = to_dev(mask)
    b, _ = to_dev(bg)
    return back(ops.blend(_lib.BLEND_REPLACE, f, m, b), as_np)

"""Batched, device-resident clip pipelines: the per-frame stages of the
reference's pipeline scripts run over [N,H,W,...] stacks with no host round
trip (per-frame branches are taken on the device), which is how the per-frame
hot path gets anywhere near the HBM roofline (a 1080p stage moves ~10 MB: less
than one kernel launch).  Results are identical, frame by frame, to the
per-frame agents in ``video_unscreen_b200.unscreen`` (tests/test_gpu_clip.py).

Frames are processed in chunks so that the low-resolution intermediates stay
L2-resident between the kernels that produce and consume them.
"""
import numpy as np
import torch

from . import _lib, ops
from .unscreen.utils.fgfuncs import bgr2hsv_pixel
from .unscreen.utils.imgprocess import get_target_size


def _chunks(n, chunk):
    for s in range(0, n, chunk):
        yield s, min(n, s + chunk)


def cf_predict_clip(frames, segmasks, agent, chunk=32, out=None):
    """ColorFilteringAgent.forward(frame, mask, iters=0) for every frame of
    frames[N,H,W,3] / segmasks[N,H,W] with the agent's current mixtures
    (reference colorfiltering/agent.py:285-354, predict-only branch :319-321).
    Returns alpha[N,H,W]; the constant background image is ``agent.bg_color_bgr()``."""
    n, h, w, _ = frames.shape
    th, tw = get_target_size(h, w, agent.input_long_side)
    luts = agent.tables_dev()
    lut3d = agent.lut3d_dev()
    alpha = out if out is not None else torch.empty((n, h, w), dtype=torch.uint8, device=frames.device)
    fg_min, bg_min = max(agent.fg_ncomp) * 5, max(agent.bg_ncomp) * 5
    for s, e in _chunks(n, chunk):
        fr, sm = frames[s:e], segmasks[s:e]
        flags = ops.cf_degenerate_flags(ops.count_cmp(sm, _lib.CMP_GT, 128), ops.count_cmp(sm, _lib.CMP_LT, 128), fg_min, bg_min)
        hsv_lo = ops.resize_linear_image(ops.bgr2hsv(fr), th, tw)
        mask_lo = ops.resize_linear_mask(sm, th, tw)
        a = ops.cf_alpha_lut3d(hsv_lo, lut3d) if lut3d is not None else ops.cf_alpha(hsv_lo, luts)
        a = ops.cf_postprocess(a, mask_lo, 0.8)
        a = ops.resize_linear_mask(a, h, w)
        alpha[s:e] = ops.select_frames(sm, a, flags)     # degenerate masks are returned as they came (agent.py:303-307)
    return alpha


def _trimap_plain(masks, agent):
    n, h, w = masks.shape
    ih, iw = get_target_size(h, w, agent.input_long_side)
    m = ops.resize_nearest_mask(masks, ih, iw)
    if agent.kernelsize == 3 and agent.iters <= ops.CROSS_MAX_PASSES:
        tri = ops.trimap_core(m, agent.iters)
    else:
        tri = ops.trimap_classify(ops.dilate(m, agent.kernelsize, agent.iters), ops.erode(m, agent.kernelsize, agent.iters))
    return ops.trimap_snap(ops.resize_linear_mask(tri, h, w))


def trimap_clip(masks, agent, frames=None, bg=None, chunk=32, out=None):
    """TrimapAgent.forward for every frame: mask-only (trimap/agent.py:35-61) or,
    with ``frames`` and ``bg`` ((3,) colour or [H,W,3] / [N,H,W,3] image), the
    background-gated variant (:63-101) with its per-frame ratio test decided on
    the device."""
    n, h, w = masks.shape
    tri = out if out is not None else torch.empty((n, h, w), dtype=torch.uint8, device=masks.device)
    half = np.array(agent.color_winsize) // 2
    for s, e in _chunks(n, chunk):
        m = masks[s:e]
        if frames is None:
            tri[s:e] = _trimap_plain(m, agent)
            continue
        fr = frames[s:e]
        if isinstance(bg, np.ndarray) and bg.ndim == 1:
            hsv = bgr2hsv_pixel(bg)
            bgmask = ops.inrange_color(fr, np.clip(hsv - half, 10, 255), np.clip(hsv + half, 10, 255))
        else:
            b = bg[s:e] if bg.ndim == 4 else bg
            bgmask = ops.inrange_image(fr, b, half)
        flags = ops.ratio_flags(ops.count_and(m, bgmask), 0.1)       # 0: ensemble, 1: trust mask, 2: empty mask
        fuzzy = ops.mask_and01(m, bgmask)
        src = ops.select_frames(m, ops.mask_clear_where(m, fuzzy), flags)
        t = ops.set128_unflagged(_trimap_plain(src, agent), fuzzy, flags)
        # an empty mask is returned as is (all zeros); the plain branch of an all-zero mask is all zeros too
        tri[s:e] = t
    return tri


def green_clip(frames, segmasks, cf_agent, trimap_agent, chunk=16):
    """the green-screen loop of tools/unscreen/green.py:70-138 without its CNN
    stages (alpha := colour-filter alpha): cf predict -> trimap with bg colour ->
    bgimg[alpha<128] = frame[...] -> get_fg.  Returns alpha, trimap, fg, bg."""
    n, h, w, _ = frames.shape
    dev = frames.device
    alpha = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    tri = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    fg = torch.empty_like(frames)
    bgo = torch.empty_like(frames)
    bg_color = cf_agent.bg_color_bgr()
    bg_tile = torch.from_numpy(np.tile(bg_color, (1, 4, 1))).to(dev)     # constant background: a 4-pixel periodic image
    for s, e in _chunks(n, chunk):
        a = cf_predict_clip(frames[s:e], segmasks[s:e], cf_agent, chunk=chunk)
        alpha[s:e] = a
        tri[s:e] = trimap_clip(a, trimap_agent, frames[s:e], bg_color, chunk=chunk)
        f, b = ops.get_fg(frames[s:e], a, bg_tile, _lib.PATCH_ALPHA_LT128, want_bg=True)
        fg[s:e] = f
        bgo[s:e] = b
    return alpha, tri, fg, bgo


def replace_clip(fg, alpha, bg):
    """tools/replace/replace.py:74-76 for a whole clip; ``bg`` is [H,W,3] (shared) or [N,H,W,3]."""
    return ops.blend(_lib.BLEND_REPLACE, fg, alpha, bg)


def bgstep_clip(frames, masks, trimap_agent, thr=25, chunk=16):
    """bg_step: exact temporal-median background, then per frame the difference
    gate (bg_offline.py:154-160), mask-only trimap (:166) and get_fg with the
    alpha==0 patch (:171-172), CNN stage skipped (alpha := gated mask).
    Returns background, alpha, trimap, fg."""
    n, h, w, _ = frames.shape
    dev = frames.device
    bg = ops.temporal_median(frames)
    alpha = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    tri = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    fg = torch.empty_like(frames)
    for s, e in _chunks(n, chunk):
        g = ops.dilate(ops.bgdiff_gray(frames[s:e], bg, thr), 4, 2)
        a = ops.gate(masks[s:e], g)
        alpha[s:e] = a
        tri[s:e] = trimap_clip(a, trimap_agent, chunk=chunk)
        fg[s:e] = ops.get_fg(frames[s:e], a, bg, _lib.PATCH_ALPHA_EQ0)
    return bg, alpha, tri, fg

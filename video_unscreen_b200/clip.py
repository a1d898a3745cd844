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


class Graphed:
    """A clip pipeline captured as ONE CUDA graph: ``Graphed(fn)`` runs ``fn()`` once eagerly (warm-up: kernel
    attributes, allocator), captures a second run, and every ``replay()`` relaunches the whole launch sequence
    (about a hundred kernels for a 300-frame clip) with a single driver call.  The captured kernels read and write
    the very tensors ``fn`` closed over: refill those in place between replays.  ``result`` is what ``fn`` returned
    during capture (tensors owned by the graph's memory pool, overwritten by every replay)."""

    def __init__(self, fn):
        fn()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def replay(self):
        self.graph.replay()
        return self.result


def _chunks(n, chunk):
    for s in range(0, n, chunk):
        yield s, min(n, s + chunk)


_SIDE_STREAMS = {}


def _overlap_chunks(n, chunk, body, streams=2):
    """``body(s, e)`` for every chunk [s, e) of n frames.  The chunks of these pipelines are independent, so with
    ``streams`` > 1 consecutive chunks go to alternating side streams: the partial last waves and launch gaps of one
    chunk's kernels are filled by the next chunk's.  Inputs must be ready on the current stream; everything the bodies
    wrote is ready on it at return.  Sequential while a CUDA graph is being captured."""
    spans = list(_chunks(n, chunk))
    if streams <= 1 or len(spans) <= 1 or torch.cuda.is_current_stream_capturing():
        for s, e in spans:
            body(s, e)
        return
    cur = torch.cuda.current_stream()
    key = (cur.device, streams)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=cur.device) for _ in range(streams)]
    pool = _SIDE_STREAMS[key][:min(streams, len(spans))]
    ready = torch.cuda.Event()
    ready.record(cur)
    for st in pool:
        st.wait_event(ready)
    for i, (s, e) in enumerate(spans):
        with torch.cuda.stream(pool[i % len(pool)]):
            body(s, e)
    for st in pool:
        done = torch.cuda.Event()
        done.record(st)
        cur.wait_event(done)


def cf_predict_clip(frames, segmasks, agent, chunk=64, out=None, streams=2):
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

    def body(s, e):
        fr, sm = frames[s:e], segmasks[s:e]
        if ops.cf_lowres_supported(h, w, th, tw):
            # one pass over the frames (the early-out counts of the masks ride along where the pass sees every mask
            # byte), then threshold + d2e2e2d2 marching in registers, then the up-scale
            if ops.cf_lowres_counts_supported(fr, sm, th, tw):
                a_lo, stats, mc = ops.cf_lowres(fr, sm, th, tw, lut3d, want_mask_counts=True)
                flags = ops.degenerate_flags_from_counts(mc, fg_min, bg_min)
            else:
                flags = ops.cf_degenerate_flags(sm, fg_min, bg_min)
                a_lo, stats = ops.cf_lowres(fr, sm, th, tw, lut3d)
            a_lo = ops.cross_chain(a_lo, [(_lib.DILATE, 2), (_lib.ERODE, 2), (_lib.ERODE, 2), (_lib.DILATE, 2)], stats, 0.8)
        else:
            flags = ops.cf_degenerate_flags(sm, fg_min, bg_min)
            hsv_lo = ops.resize_linear_image(ops.bgr2hsv(fr), th, tw)
            a_lo = ops.cf_postprocess(ops.cf_alpha_lut3d(hsv_lo, lut3d), ops.resize_linear_mask(sm, th, tw), 0.8)
        # degenerate masks are returned as they came (agent.py:303-307): alt_src / alt_flags
        ops.resize_up(a_lo, h, w, alt_src=sm, alt_flags=flags, out=alpha[s:e])
    _overlap_chunks(n, chunk, body, streams)
    return alpha


def _trimap_tail(masks, agent, fuzzy=None, flags=None, out=None):
    """nearest down (+ ensemble clearing) -> dilate/erode/classify -> bilinear up + snap (+ fuzzy override); written into
    ``out`` when given"""
    n, h, w = masks.shape
    ih, iw = get_target_size(h, w, agent.input_long_side)
    if agent.kernelsize == 3 and ops.trimap_bits_supported(masks, ih, iw, agent.iters, fuzzy):
        # exact 2x / 4x working resolution (1080p, 4K): the whole tail in bit logic, two launches
        return ops.trimap_bits(masks, ih, iw, agent.iters, fuzzy, flags, out=out)
    m = ops.trimap_src_lo(masks, ih, iw, fuzzy, flags)
    if agent.kernelsize == 3 and agent.iters <= ops.CROSS_MAX_PASSES:
        tri = ops.trimap_core(m, agent.iters)
    else:
        tri = ops.trimap_classify(ops.dilate(m, agent.kernelsize, agent.iters), ops.erode(m, agent.kernelsize, agent.iters))
    if (ih, iw) == (h, w) or (ih == 2 * h and iw == 2 * w):   # copy / cv2's INTER_AREA corner cases: unfused tail
        t = ops.trimap_snap(ops.resize_linear_mask(tri, h, w))
        t = ops.set128_unflagged(t, fuzzy, flags) if fuzzy is not None else t
        if out is not None:
            out.copy_(t)
            return out
        return t
    return ops.resize_up(tri, h, w, mode=1, fuzzy=fuzzy, flags=flags, out=out)


def trimap_clip(masks, agent, frames=None, bg=None, chunk=64, out=None, streams=2):
    """TrimapAgent.forward for every frame: mask-only (trimap/agent.py:35-61) or,
    with ``frames`` and ``bg`` ((3,) colour or [H,W,3] / [N,H,W,3] image), the
    background-gated variant (:63-101) with its per-frame ratio test decided on
    the device."""
    n, h, w = masks.shape
    tri = out if out is not None else torch.empty((n, h, w), dtype=torch.uint8, device=masks.device)
    half = np.array(agent.color_winsize) // 2

    def body(s, e):
        m = masks[s:e]
        if frames is None:
            _trimap_tail(m, agent, out=tri[s:e])
            return
        fr = frames[s:e]
        if isinstance(bg, np.ndarray) and bg.ndim == 1:
            hsv = bgr2hsv_pixel(bg)
            fuzzy, counts = ops.fuzzy_count(fr, m, np.clip(hsv - half, 10, 255), np.clip(hsv + half, 10, 255))
        else:
            b = bg[s:e] if bg.ndim == 4 else bg
            bgmask = ops.inrange_image(fr, b, half)
            counts = ops.count_and(m, bgmask)
            fuzzy = ops.mask_and01(m, bgmask)
        flags = ops.ratio_flags(counts, 0.1)       # 0: ensemble, 1: trust mask, 2: empty mask
        # an empty mask is returned as is (all zeros); the plain branch of an all-zero mask is all zeros too
        _trimap_tail(m, agent, fuzzy, flags, out=tri[s:e])
    _overlap_chunks(n, chunk, body, streams)
    return tri


def cf_trimap_clip(frames, segmasks, cf_agent, trimap_agent, bg_color=None, chunk=50, out_alpha=None, out_trimap=None, streams=2):
    """BASELINE config 1: ColorFilteringAgent.forward(iters=0) then TrimapAgent.forward(alpha, frame, bg colour) for every
    frame, chunk by chunk.  Chunks bound the temporaries (about 12 bytes per pixel and frame); make them as large as memory
    allows: 300 x 1080p takes 4.1 ms in chunks of 30, 3.6 ms in chunks of 100, 3.4 ms in one piece (launch gaps and the
    partial last wave of every kernel); with ``streams`` = 2 consecutive chunks overlap on two streams and fill each
    other's gaps: 3.05 ms in chunks of 50.  Returns alpha, trimap."""
    n, h, w, _ = frames.shape
    dev = frames.device
    alpha = out_alpha if out_alpha is not None else torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    tri = out_trimap if out_trimap is not None else torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    if bg_color is None:
        bg_color = cf_agent.bg_color_bgr()
    def body(s, e):
        a = cf_predict_clip(frames[s:e], segmasks[s:e], cf_agent, chunk=chunk, out=alpha[s:e])
        trimap_clip(a, trimap_agent, frames[s:e], bg_color, chunk=chunk, out=tri[s:e])
    cf_agent.tables_dev(), cf_agent.lut3d_dev()      # built (once) on the current stream, not on a side stream
    _overlap_chunks(n, chunk, body, streams)
    return alpha, tri


def color_correct_clip(frames, alpha, bg_color, target_long_side=960, mean_exp=0.95, chunk=64, out=None, streams=2):
    """color_correct (imgprocess.py:263-300) over a clip, chunk by chunk (``out`` must not alias ``alpha``)."""
    from .unscreen.utils.imgprocess import get_target_size
    n, h, w, _ = frames.shape
    th, tw = get_target_size(h, w, target_long_side)
    if out is None:
        out = torch.empty_like(alpha)
    _overlap_chunks(n, chunk, lambda s, e: ops.color_correct(frames[s:e], alpha[s:e], bg_color, th, tw, mean_exp, out=out[s:e]), streams)
    return out


def green_clip(frames, segmasks, cf_agent, trimap_agent, chunk=24, bg_color=None, bg_tile=None, color_correct=False, streams=2):
    """the green-screen loop of tools/unscreen/green.py:70-138 without its CNN
    stages (alpha := colour-filter alpha): cf predict -> trimap with bg colour ->
    [color_correct, green.py:120, when asked for] -> bgimg[alpha<128] = frame[...] -> get_fg.
    Returns alpha, trimap, fg, bg.
    ``bg_color`` ((3,) BGR, host) / ``bg_tile`` ([1,4,3] device) default to the agent's background colour; pass them
    in when the call is captured into a CUDA graph (fetching them synchronises)."""
    n, h, w, _ = frames.shape
    dev = frames.device
    alpha = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    tri = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    fg = torch.empty_like(frames)
    bgo = torch.empty_like(frames)
    if bg_color is None:
        bg_color = cf_agent.bg_color_bgr()
    if bg_tile is None:
        bg_tile = torch.from_numpy(np.tile(bg_color, (1, 4, 1))).to(dev)     # constant background: a 4-pixel periodic image
    def body(s, e):
        # every stage writes straight into its slice of the clip-sized results
        a = cf_predict_clip(frames[s:e], segmasks[s:e], cf_agent, chunk=chunk, out=alpha[s:e])
        trimap_clip(a, trimap_agent, frames[s:e], bg_color, chunk=chunk, out=tri[s:e])
        if color_correct:
            a = color_correct_clip(frames[s:e], a.clone(), bg_color, chunk=chunk, out=alpha[s:e])
        ops.get_fg(frames[s:e], a, bg_tile, _lib.PATCH_ALPHA_LT128, want_bg=True, out=fg[s:e], bg_out=bgo[s:e])
    cf_agent.tables_dev(), cf_agent.lut3d_dev()      # built (once) on the current stream, not on a side stream
    _overlap_chunks(n, chunk, body, streams)
    return alpha, tri, fg, bgo


def replace_clip(fg, alpha, bg, dx=None, dy=None, scale=None):
    """tools/replace/replace.py:69-76 for a whole clip; ``bg`` is [H,W,3] (shared) or [N,H,W,3].  With ``dx``/``dy``
    and/or ``scale`` the foreground and its mask first go through shift_fg / rescale_fg (:69-72) like in the script;
    without them this is the blend of :74-76 alone (BASELINE config 4).  Whole-clip launches: these kernels fill the
    machine on their own (chunks on two streams were slower, 2.43 against 2.27 ms per 120 frames)."""
    ach = 3 if alpha.ndim == fg.ndim else 1
    if dx is not None or dy is not None:
        fg = ops.shift(fg, dx or 0, dy or 0, 3)
        alpha = ops.shift(alpha, dx or 0, dy or 0, ach)
    if scale is not None:
        fg = ops.rescale_cubic(fg, scale, 3)
        alpha = ops.rescale_cubic(alpha, scale, ach)
    return ops.blend(_lib.BLEND_REPLACE, fg, alpha, bg)


def bgstep_clip(frames, masks, trimap_agent, thr=25, chunk=24, streams=2):
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
    def body(s, e):
        # every stage writes straight into its slice of the clip-sized results
        if ops.bgdiff_gate_supported(frames[s:e], bg, masks[s:e]):
            a = ops.bgdiff_gate(frames[s:e], bg, masks[s:e], thr, out=alpha[s:e])
        else:
            a = ops.gate(masks[s:e], ops.dilate(ops.bgdiff_gray(frames[s:e], bg, thr), 4, 2))
            alpha[s:e] = a
        trimap_clip(a, trimap_agent, chunk=chunk, out=tri[s:e])
        ops.get_fg(frames[s:e], a, bg, _lib.PATCH_ALPHA_EQ0, out=fg[s:e])
    _overlap_chunks(n, chunk, body, streams)
    return bg, alpha, tri, fg


BGSTEP_HALO = 24   # full-resolution rows: 6 working-resolution rows at 1/4 scale (r=5 diamond + one bilinear tap); covers dilate(4,2)


def bgstep_clip_tile(frames, masks, trimap_agent, rank, world, thr=25, chunk=24, scale=4):
    """bgstep_clip on this rank's ROW TILE of the clip (BASELINE config 5: spatial-tile sharding):
    the temporal median of the tile's rows (every pixel is independent), then the per-frame stages on the tile plus a
    halo of BGSTEP_HALO rows read from the local frames, cropped back.  ``frames`` / ``masks`` are the WHOLE frames here
    (the caller may hold only rows [r0 - halo, r1 + halo) and pass those with the matching offsets instead).  Tile
    boundaries are multiples of ``scale`` (frame size / working size), so the tile's down-scales sample the pixels the
    whole frame's do: the results equal the corresponding rows of bgstep_clip on the whole clip, bit for bit.
    Returns (r0, r1), background, alpha, trimap, fg for rows [r0, r1)."""
    from . import shard
    n, h, w, _ = frames.shape
    r0, r1, ht, hb = shard.my_row_tile(h, rank, world, halo=BGSTEP_HALO, align=scale)
    a0, a1 = r0 - ht, r1 + hb
    ftile = frames[:, a0:a1].contiguous()
    mtile = masks[:, a0:a1].contiguous()
    bg_t, alpha_t, tri_t, fg_t = bgstep_clip(ftile, mtile, trimap_agent, thr=thr, chunk=chunk)
    lo, hi = ht, ht + (r1 - r0)
    return (r0, r1), bg_t[lo:hi], alpha_t[:, lo:hi], tri_t[:, lo:hi], fg_t[:, lo:hi]

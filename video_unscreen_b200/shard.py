"""Partitioning of the hot path across the GPUs of one box.

No data-path collective is needed: per-frame stages shard by contiguous frame
range (mirroring the reference's ``--range a-b``, tools/unscreen/green.py:147),
temporal reducers shard by row tile because every pixel is independent.  The
functions here are pure host-side arithmetic plus an optional gather of result
tiles through ``torch.distributed`` (off the hot path, once per clip).
"""


def frame_ranges(n_frames, world, align=1):
    """contiguous [start, stop) per rank; interior boundaries are multiples of
    ``align`` (use the colour-filter refit cadence, 30, so every shard starts on
    a refit frame -- green.py:88)."""
    if world < 1 or n_frames < 0 or align < 1:
        raise ValueError("bad arguments")
    units = -(-n_frames // align)  # ceil
    out, start = [], 0
    for r in range(world):
        u = units // world + (1 if r < units % world else 0)
        stop = min(n_frames, start + u * align)
        out.append((start, stop))
        start = stop
    return out


def bgstep_halo(scale, trimap_iters=5):
    """(top, bottom) halo rows a row tile needs so that the bg_step per-frame stages (difference gate -> mask-only trimap
    -> get_fg) on the tile equal the same rows of the whole frame, for a working resolution of 1/``scale``:

    * trimap row y of a tile starting at r0 (a multiple of ``scale``) takes the bilinear taps r0/scale - 1 and r0/scale
      of the working-resolution trimap; a working-resolution row j looks ``trimap_iters`` rows up and down (3x3 cross);
      working row j is full-resolution row j*scale (nearest): the gated alpha is read from row r0 - scale*(iters+1) on;
    * that alpha row comes out of dilate_mask(gray, 4, 2), which reaches 4 rows up and 2 rows down (anchor (2,2) of the
      4x4 ellipse, twice).

    top = scale*(iters+1) + 4; bottom = scale*iters + 2 (+1: rows, not offsets), both rounded up to multiples of
    ``scale`` so that the tile plus halo still starts and ends on the working-resolution grid.  scale 4, iters 5 (4K):
    (28, 24); scale 2 (1080p): (16, 14)."""
    up = lambda v: -(-v // scale) * scale
    return up(scale * (trimap_iters + 1) + 4), up(scale * trimap_iters + 3)


def row_tiles(height, world, halo=0, align=1):
    """[(row_start, row_stop, halo_top, halo_bottom)] per rank.  ``halo`` rows
    (one number, or (top, bottom): see ``bgstep_halo``) of neighbouring tiles
    are needed when a tile is later pushed through the stencil stages without
    gathering.  Interior boundaries are multiples of ``align`` (the down-scale factor of
    the working resolution, 4 at 4K, so that a tile's nearest / area
    down-scale samples the same pixels as the whole frame's); halos are
    clipped to the image."""
    top, bottom = (halo, halo) if isinstance(halo, int) else halo
    if world < 1 or height < 0 or top < 0 or bottom < 0 or align < 1:
        raise ValueError("bad arguments")
    units = -(-height // align)
    out, start = [], 0
    for r in range(world):
        u = units // world + (1 if r < units % world else 0)
        stop = min(height, start + u * align)
        out.append((start, stop, min(top, start), min(bottom, height - stop)))
        start = stop
    return out


def row_tiles_weighted(row_cost, world, halo=0, align=1):
    """``row_tiles`` with content-aware heights: contiguous tiles whose interior boundaries are multiples of ``align`` and
    whose MAXIMUM cost is as small as such a partition allows, for a per-row cost estimate ``row_cost`` (length = image
    height; e.g. for the bg_step stages a constant per row plus a term per segmentation-mask pixel of the row, summed over
    the clip: where the person stands, get_fg and the trimap's unknown band cost more, and the job takes the slowest rank).
    A tile's cost counts its halo rows too (they are recomputed).  Every rank gets at least one ``align`` unit when the
    image has that many.  Same return format as ``row_tiles``; equal costs give the same heights up to one unit.

    Host-side arithmetic only: the masks (1 byte per pixel) can be reduced to row sums before the frames (3 bytes per
    pixel) are placed, so the tiles can be chosen before a rank loads its rows."""
    top, bottom = (halo, halo) if isinstance(halo, int) else halo
    cost = [float(c) for c in row_cost]
    height = len(cost)
    if world < 1 or top < 0 or bottom < 0 or align < 1 or any(c < 0 for c in cost):
        raise ValueError("bad arguments")
    units = -(-height // align)
    if units <= world:                       # nothing to balance: one unit per rank, the rest empty
        return row_tiles(height, world, halo, align)
    pre = [0.0]
    for c in cost:
        pre.append(pre[-1] + c)

    def tile_cost(u0, u1):                   # units [u0, u1) plus the halo rows around them
        a, b = u0 * align, min(height, u1 * align)
        return pre[min(height, b + bottom)] - pre[max(0, a - top)]

    def cuts_for(limit):
        """greedy: every tile as long as its cost stays within ``limit`` while leaving one unit for every later rank; None if
        ``world`` tiles cannot cover the image that way"""
        cuts, u = [0], 0
        for r in range(world):
            left = world - r - 1
            end = u + 1
            if tile_cost(u, end) > limit:
                return None
            while end < units - left and tile_cost(u, end + 1) <= limit:
                end += 1
            if left == 0 and end < units:
                return None
            u = end
            cuts.append(u)
        return cuts

    lo = max(tile_cost(u, u + 1) for u in range(units))
    hi = tile_cost(0, units)
    best = cuts_for(hi)
    for _ in range(60):                      # bisection on the admissible maximum
        mid = 0.5 * (lo + hi)
        got = cuts_for(mid)
        if got is None:
            lo = mid
        else:
            hi, best = mid, got
        if hi - lo <= 1e-9 * max(hi, 1.0):
            break
    out = []
    for r in range(world):
        start, stop = best[r] * align, min(height, best[r + 1] * align)
        out.append((start, stop, min(top, start), min(bottom, height - stop)))
    return out


def my_frame_range(n_frames, rank, world, align=1):
    return frame_ranges(n_frames, world, align)[rank]


def my_row_tile(height, rank, world, halo=0, align=1):
    return row_tiles(height, world, halo, align)[rank]


def reduce_rows_sharded(frames, fn, rank, world, gather=False):
    """temporal reduction of frames[N,H,W,C] over this rank's row tile with
    ``fn(frames_tile) -> [rows,W,C]``; optionally assemble the full result on
    every rank with one all_gather (not on the hot path)."""
    h = frames.shape[1]
    r0, r1, _, _ = my_row_tile(h, rank, world)
    tile = fn(frames[:, r0:r1])
    if not gather or world == 1:
        return tile, (r0, r1)
    import torch
    import torch.distributed as dist
    t = tile if isinstance(tile, torch.Tensor) else torch.from_numpy(tile)
    sizes = [b - a for a, b, _, _ in row_tiles(h, world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    full = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
    return (full if isinstance(tile, torch.Tensor) else full.numpy()), (0, h)

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

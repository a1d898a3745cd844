"""N>1 host logic on CPU: partition arithmetic, and a world_size-2 gloo run in
which every rank reduces its own row tile (the oracle stands in for the kernel:
this test is about the sharding, not the arithmetic) and the tiles are gathered."""
import os
import socket

import numpy as np
import pytest

from video_unscreen_b200 import shard


def test_frame_ranges_cover_and_align():
    for n, w, a in [(300, 8, 30), (300, 4, 30), (2000, 8, 30), (7, 8, 1), (0, 2, 1), (301, 2, 30)]:
        rs = shard.frame_ranges(n, w, a)
        assert len(rs) == w and rs[0][0] == 0 and rs[-1][1] == n
        for (a0, a1), (b0, b1) in zip(rs, rs[1:]):
            assert a1 == b0 and a0 <= a1
        for s, e in rs[:-1]:
            assert e % a == 0 or e == n


def test_row_tiles_cover_with_halo():
    for h, w in [(1080, 8), (2160, 8), (1080, 2), (5, 8), (33, 4)]:
        ts = shard.row_tiles(h, w, halo=24)
        assert ts[0][0] == 0 and ts[-1][1] == h
        for (a0, a1, at, ab), (b0, b1, bt, bb) in zip(ts, ts[1:]):
            assert a1 == b0
        for r0, r1, ht, hb in ts:
            assert 0 <= r0 - ht and r1 + hb <= h


def test_row_tiles_weighted_balance_the_cost():
    """content-aware tile heights: same cover / alignment / halo rules as row_tiles, and never a larger maximum cost than
    the equal-height partition (the job takes the slowest rank)."""
    rng = np.random.default_rng(4)

    def cost_of(tiles, cost):
        h = len(cost)
        return [float(cost[max(0, a - ht): min(h, b + hb)].sum()) for a, b, ht, hb in tiles]

    cases = [(2160, 8, (28, 24), 4), (2160, 4, (28, 24), 4), (1080, 8, (16, 14), 2), (1080, 3, 0, 1), (96, 5, (8, 8), 4), (37, 4, 2, 1)]
    for h, world, halo, align in cases:
        y = np.arange(h)
        person = 1.0 + 6.0 * np.exp(-((y - 0.5 * h) / (0.16 * h)) ** 2)          # the person stands in the middle rows
        for cost in (person, np.ones(h), rng.random(h) + 0.01, np.where(y < h // 3, 10.0, 0.0)):
            ts = shard.row_tiles_weighted(cost, world, halo, align)
            assert len(ts) == world and ts[0][0] == 0 and ts[-1][1] == h
            for (a0, a1, _, _), (b0, b1, _, _) in zip(ts, ts[1:]):
                assert a1 == b0 and a0 < a1
            for r0, r1, ht, hb in ts:
                assert 0 <= r0 - ht and r1 + hb <= h and (r0 % align == 0) and (r1 % align == 0 or r1 == h)
            top, bottom = (halo, halo) if isinstance(halo, int) else halo
            for r0, r1, ht, hb in ts:
                assert ht == min(top, r0) and hb == min(bottom, h - r1)
            even = shard.row_tiles(h, world, halo, align)
            assert max(cost_of(ts, cost)) <= max(cost_of(even, cost)) * (1 + 1e-9)
    # the middle tiles of a person-shaped cost are the short ones
    ts = shard.row_tiles_weighted(1.0 + 6.0 * np.exp(-((np.arange(2160) - 1080) / 350.0) ** 2), 8, (28, 24), 4)
    heights = [b - a for a, b, _, _ in ts]
    assert min(heights) == min(heights[3:5]) and max(heights) in (heights[0], heights[-1])
    # fewer units than ranks: the equal partition's answer (empty tiles at the end)
    assert shard.row_tiles_weighted(np.ones(5), 8, 0, 1) == shard.row_tiles(5, 8, 0, 1)
    with pytest.raises(ValueError):
        shard.row_tiles_weighted([1.0, -1.0], 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle import refport as R
    from video_unscreen_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames, _, _ = synth.bgstep_clip(9, 21, 16, seed=2)  # 21 rows: uneven tiles
    full, span = shard.reduce_rows_sharded(frames, R.temporal_median, rank, world, gather=True)
    tile, (r0, r1) = shard.reduce_rows_sharded(frames, R.temporal_median, rank, world, gather=False)
    want = R.temporal_median(frames)
    ok = np.array_equal(full, want) and np.array_equal(tile, want[r0:r1]) and span == (0, 21)
    fr = shard.my_frame_range(300, rank, world, 30)
    q.put((rank, bool(ok), fr))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == (0, 150) and res[1][2] == (150, 300)


def _halo_case(h=128, w=256, seed=0):
    """frames whose whole-frame results depend on rows up to 28 above a tile boundary: sparse difference pixels and sparse
    mask pixels around row h/2 (a diff pixel at r0-26 plus a mask pixel at r0-24 changes the trimap at r0)."""
    rng = np.random.default_rng(seed)
    bg = np.full((h, w, 3), 90, np.uint8)
    frame = bg.copy()
    mask = np.zeros((h, w), np.uint8)
    r0 = h // 2
    for x in range(8, w - 8, 8):
        d = int(rng.integers(20, 32))
        frame[r0 - d - 2, x] = 250          # a difference two rows above ...
        mask[r0 - d, x] = 255               # ... opens the gate for a mask pixel at r0 - d (dilate(4,2) reaches 4 up / 2 down)
        d = int(rng.integers(18, 28))
        frame[r0 + d + 1, x + 4] = 250
        mask[r0 + d, x + 4] = 255
    return frame, bg, mask


def _tile_pipeline_oracle(frame, bg, mask, a0, a1, long_side):
    from oracle import refport as R
    alpha = R.bgdiff_gate(frame[a0:a1], bg[a0:a1], mask[a0:a1], 25)
    return alpha, R.generate_trimap(alpha, long_side)


def test_bgstep_halo_covers_the_stencils():
    """the (top, bottom) halo of shard.bgstep_halo makes a row tile's gate + trimap equal the whole frame's (oracle
    arithmetic, CPU); the old symmetric 24-row halo does not (the regression this guards against)."""
    from oracle import refport as R
    assert shard.bgstep_halo(4, 5) == (28, 24) and shard.bgstep_halo(2, 5) == (16, 14)
    h, w, long_side, scale = 128, 256, 64, 4
    bad = 0
    for seed in range(4):
        frame, bg, mask = _halo_case(h, w, seed)
        alpha_w = R.bgdiff_gate(frame, bg, mask, 25)
        tri_w = R.generate_trimap(alpha_w, long_side)
        assert (tri_w == 128).any()
        for halo in (shard.bgstep_halo(scale, 5), 24):
            for r0, r1, ht, hb in shard.row_tiles(h, 2, halo=halo, align=scale):
                a, t = _tile_pipeline_oracle(frame, bg, mask, r0 - ht, r1 + hb, long_side)
                same = np.array_equal(a[ht:ht + r1 - r0], alpha_w[r0:r1]) and np.array_equal(t[ht:ht + r1 - r0], tri_w[r0:r1])
                if halo == 24:
                    bad += not same
                else:
                    assert same, (seed, r0, r1)
    assert bad > 0, "the test data no longer exercises rows 25..28 above the boundary"

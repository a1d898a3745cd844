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

"""Batched clip pipelines == the per-frame agents, frame by frame (which are
themselves checked against the reference goldens in test_gpu_parity.py), and
against the oracle on the same seeded clips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import refport as R  # noqa: E402
from video_unscreen_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def env(golden):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_oracle_golden import cf_tables
    from video_unscreen_b200 import clip, ops
    from video_unscreen_b200.unscreen.colorfiltering import ColorFilteringAgent
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    c = golden("colorfilter")

    class E:
        pass
    e = E()
    e.clip, e.ops, e.CF, e.TA = clip, ops, ColorFilteringAgent, TrimapAgent
    e.tables = {tag: cf_tables(c, tag) for tag in ("x2", "x4")}
    return e


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag,h,w,L", [("x2", 270, 480, 240), ("x4", 360, 640, 160)])
def test_green_clip_matches_agents_and_oracle(env, tag, h, w, L):
    n = 7
    frames, segs = synth.green_clip(n, h, w, seed=5)
    segs[3] = 0          # degenerate: no foreground -> early-out
    segs[5] = 255        # degenerate: no background
    lb, lf, bgh = env.tables[tag]
    cf = env.CF(input_long_side=L)
    cf.set_tables(lb, lf, bgh)
    ta = env.TA(input_long_side=L)
    alpha, tri, fg, bgo = env.clip.green_clip(dev(frames), dev(segs), cf, ta, chunk=3)
    alpha, tri, fg, bgo = (t.cpu().numpy() for t in (alpha, tri, fg, bgo))
    bg_color = cf.bg_color_bgr()
    for i in range(n):
        a_i, bgimg_i, _ = cf.forward(frames[i], segs[i], 0)
        assert np.array_equal(alpha[i], a_i), i
        assert np.array_equal(tri[i], ta.forward(a_i, frames[i], bg_color)), i
        # oracle, end to end
        a_o, _, _ = R.cf_forward_predict(frames[i], segs[i], lb, lf, bgh, L)
        assert np.array_equal(alpha[i], a_o), i
        assert np.array_equal(tri[i], R.generate_trimap_withbg(a_o, frames[i], bg_color, L)), i
        bg_full = np.broadcast_to(bg_color, frames[i].shape)
        patched = R.patch_bg(bg_full, frames[i], a_o, "lt128")
        assert np.array_equal(bgo[i], patched), i
        assert np.array_equal(fg[i], R.get_fg(frames[i], a_o, patched)), i


@pytest.mark.parametrize("tag,h,w,L", [("x2", 108, 192, 96), ("x4", 216, 384, 96)])
def test_green_clip_with_color_correct(env, tag, h, w, L):
    """green.py:99-126 with the colour-correction stage (:120) between the trimap and get_fg; the default working
    resolution of color_correct (960) is above these frames, so it runs at full resolution here."""
    n = 4
    frames, segs = synth.green_clip(n, h, w, seed=8)
    lb, lf, bgh = env.tables[tag]
    cf = env.CF(input_long_side=L)
    cf.set_tables(lb, lf, bgh)
    ta = env.TA(input_long_side=L)
    alpha, tri, fg, bgo = (t.cpu().numpy() for t in env.clip.green_clip(dev(frames), dev(segs), cf, ta, chunk=3, color_correct=True))
    bg_color = cf.bg_color_bgr()
    for i in range(n):
        a_o, _, _ = R.cf_forward_predict(frames[i], segs[i], lb, lf, bgh, L)
        assert np.array_equal(tri[i], R.generate_trimap_withbg(a_o, frames[i], bg_color, L)), i
        a_c = R.color_correct(frames[i], a_o, bg_color)
        assert np.array_equal(alpha[i], a_c), i
        patched = R.patch_bg(np.broadcast_to(bg_color, frames[i].shape), frames[i], a_c, "lt128")
        assert np.array_equal(bgo[i], patched), i
        assert np.array_equal(fg[i], R.get_fg(frames[i], a_c, patched)), i


def test_trimap_clip_variants(env, golden):
    t = golden("trimap")
    L = int(t["x2_L"])
    ta = env.TA(input_long_side=L)
    masks = np.stack([t["x2_soft"], t["x2_leak"], t["x2_ring"], np.zeros_like(t["x2_mask"]), t["x2_mask"]])
    frames = np.stack([t["x2_frame"]] * 5)
    bgcol = np.array([60, 200, 40], np.uint8)
    got = env.clip.trimap_clip(dev(masks), ta, dev(frames), bgcol, chunk=2).cpu().numpy()
    want = [t["x2_withcolor"], t["x2_leak_withcolor"], t["x2_ring_withcolor"], np.zeros_like(t["x2_mask"]), None]
    for i in range(5):
        ref = want[i] if want[i] is not None else R.generate_trimap_withbg(masks[i], frames[i], bgcol, L)
        assert np.array_equal(got[i], ref), i
    got = env.clip.trimap_clip(dev(masks), ta, dev(frames), dev(t["x2_bgimg"]), chunk=5).cpu().numpy()
    assert np.array_equal(got[0], t["x2_withimage"]) and np.array_equal(got[2], t["x2_ring_withimage"])
    got = env.clip.trimap_clip(dev(masks), ta).cpu().numpy()
    assert np.array_equal(got[0], t["x2_plain_soft"]) and np.array_equal(got[4], t["x2_plain"])


def test_bgstep_and_replace_clip(env):
    n, h, w = 12, 96, 160
    frames, masks, _ = synth.bgstep_clip(n, h, w, seed=4)
    ta = env.TA(input_long_side=80)
    bg, alpha, tri, fg = (x.cpu().numpy() for x in env.clip.bgstep_clip(dev(frames), dev(masks), ta, thr=25, chunk=5))
    bg_o = R.temporal_median(frames)
    assert np.array_equal(bg, bg_o)
    for i in range(n):
        a_o = R.bgdiff_gate(frames[i], bg_o, masks[i], 25)
        assert np.array_equal(alpha[i], a_o), i
        assert np.array_equal(tri[i], R.generate_trimap(a_o, 80)), i
        assert np.array_equal(fg[i], R.get_fg(frames[i], a_o, R.patch_bg(bg_o, frames[i], a_o, "eq0"))), i
    rng = np.random.default_rng(0)
    newbg = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    out = env.clip.replace_clip(dev(fg), dev(alpha), dev(newbg)).cpu().numpy()
    for i in range(n):
        assert np.array_equal(out[i], R.replace_blend(fg[i], alpha[i], newbg)), i


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_equals_whole(env, world):
    """SURVEY section 8e on one GPU, shards run one after the other: frame-range shards of the per-frame stages and
    row-tile shards of the temporal median (video_unscreen_b200.shard) give exactly the whole-clip results."""
    from video_unscreen_b200 import shard
    n, h, w = 24, 96, 160
    frames, masks, _ = synth.bgstep_clip(n, h, w, seed=6)
    f_d, m_d = dev(frames), dev(masks)
    ta = env.TA(input_long_side=80)
    whole_bg, whole_alpha, whole_tri, whole_fg = env.clip.bgstep_clip(f_d, m_d, ta, thr=25, chunk=5)
    # temporal median by row tiles (no halo: every pixel is independent)
    tiles = [env.ops.temporal_median(f_d[:, r0:r1].contiguous()) for r0, r1, _, _ in shard.row_tiles(h, world)]
    assert torch.equal(torch.cat(tiles, 0), whole_bg)
    # per-frame stages by frame range, background shared
    alphas, tris, fgs = [], [], []
    for s, e in shard.frame_ranges(n, world, align=1):
        if e == s:
            continue
        a = env.ops.bgdiff_gate(f_d[s:e], whole_bg, m_d[s:e], 25)
        alphas.append(a)
        tris.append(env.clip.trimap_clip(a, ta, chunk=4))
        fgs.append(env.ops.get_fg(f_d[s:e], a, whole_bg, 2))
    assert torch.equal(torch.cat(alphas, 0), whole_alpha)
    assert torch.equal(torch.cat(tris, 0), whole_tri)
    assert torch.equal(torch.cat(fgs, 0), whole_fg)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_row_tile_sharding_with_halo(env, world):
    """BASELINE config 5's layout on one GPU, tiles run one after the other: every rank's row tile (+ 28 / 24-row halo) through
    median, difference gate, trimap and get_fg equals the same rows of the whole-clip pipeline."""
    n, h, w = 10, 384, 640                    # working resolution 96 x 160: scale 4, like 4K -> 540 x 960
    frames, masks, _ = synth.bgstep_clip(n, h, w, seed=8)
    f_d, m_d = dev(frames), dev(masks)
    ta = env.TA(input_long_side=160)
    bg, alpha, tri, fg = env.clip.bgstep_clip(f_d, m_d, ta, thr=25, chunk=4)
    covered = 0
    for rank in range(world):
        (r0, r1), bg_t, a_t, t_t, f_t = env.clip.bgstep_clip_tile(f_d, m_d, ta, rank, world, thr=25, chunk=4)
        assert r0 % 4 == 0 and (r1 % 4 == 0 or r1 == h)
        assert torch.equal(bg_t, bg[r0:r1])
        assert torch.equal(a_t, alpha[:, r0:r1])
        assert torch.equal(t_t, tri[:, r0:r1])
        assert torch.equal(f_t, fg[:, r0:r1])
        covered += r1 - r0
    assert covered == h


def test_row_tile_halo_regression(env):
    """the top halo must cover the 4-row upward reach of dilate(4,2) on top of the trimap's 24 rows; data with
    difference / mask pixels 20..32 rows off the tile boundary (tests/test_sharding_gloo.py::_halo_case).  Also: a rank
    that holds only its rows (+ halo) of the clip gets the same result, and a portrait clip is refused."""
    from test_sharding_gloo import _halo_case
    h, w = 128, 256
    cases = [_halo_case(h, w, seed) for seed in range(4)]
    frames = np.stack([c[0] for c in cases])
    masks = np.stack([c[2] for c in cases])
    # the median of these 4 frames is the flat background (differences are sparse and at different places)
    f_d, m_d = dev(frames), dev(masks)
    ta = env.TA(input_long_side=64)
    bg, alpha, tri, fg = env.clip.bgstep_clip(f_d, m_d, ta, thr=25, chunk=3)
    assert np.array_equal(bg.cpu().numpy(), R.temporal_median(frames))
    for i in range(len(cases)):
        a_o = R.bgdiff_gate(frames[i], bg.cpu().numpy(), masks[i], 25)
        assert np.array_equal(alpha[i].cpu().numpy(), a_o) and np.array_equal(tri[i].cpu().numpy(), R.generate_trimap(a_o, 64))
    for world in (2, 4):
        for rank in range(world):
            (r0, r1), bg_t, a_t, t_t, f_t = env.clip.bgstep_clip_tile(f_d, m_d, ta, rank, world, thr=25, chunk=3)
            assert torch.equal(a_t, alpha[:, r0:r1]) and torch.equal(t_t, tri[:, r0:r1]) and torch.equal(f_t, fg[:, r0:r1])
            assert torch.equal(bg_t, bg[r0:r1])
            # the rank holds rows [a0, a1) only
            _, _, ht, hb, _, _ = env.clip.bgstep_tile_geometry(h, w, ta, rank, world)
            a0, a1 = r0 - ht, r1 + hb
            (q0, q1), _, a_l, t_l, f_l = env.clip.bgstep_clip_tile(f_d[:, a0:a1].contiguous(), m_d[:, a0:a1].contiguous(), ta, rank, world,
                                                                 thr=25, chunk=3, rows=(a0, a1, h))
            assert (q0, q1) == (r0, r1) and torch.equal(a_l, a_t) and torch.equal(t_l, t_t) and torch.equal(f_l, f_t)
    with pytest.raises(ValueError):
        env.clip.bgstep_tile_geometry(250, 130, ta, 0, 2)       # 250x130 -> 64x33: no exact scale


@pytest.mark.parametrize("tag,h,w,L", [("x2", 270, 480, 240), ("x4", 384, 640, 160), ("x2", 64, 96, 48)])
def test_fused_green_chunk_equals_staged(env, tag, h, w, L):
    """the two-pass chunk (vu_cf_alpha_up_fuzzy + vu_trimap_bits_packed, clip._fused_green_chunk) against the
    stage-by-stage kernels: alpha, trimap, fg, bg, for ordinary frames, both early-outs, an all-zero matte and
    soft mattes (random working-resolution alpha through the second pass alone)."""
    n = 9
    frames, segs = synth.green_clip(n, h, w, seed=21)
    segs[2] = 0
    segs[4] = 255
    frames[6] = np.array(synth.GREEN_BG, np.uint8)          # nothing but background: alpha all zero -> empty-mask branch
    lb, lf, bgh = env.tables[tag]
    cf = env.CF(input_long_side=L)
    cf.set_tables(lb, lf, bgh)
    ta = env.TA(input_long_side=L)
    f_d, s_d = dev(frames), dev(segs)
    assert env.clip._fused_green_supported(f_d, s_d, cf, ta)
    a1, t1 = env.clip.cf_trimap_clip(f_d, s_d, cf, ta, chunk=4, fused=True)
    a0, t0 = env.clip.cf_trimap_clip(f_d, s_d, cf, ta, chunk=4, fused=False)
    assert torch.equal(a1, a0) and torch.equal(t1, t0)
    g1 = env.clip.green_clip(f_d, s_d, cf, ta, chunk=4, fused=True)
    g0 = env.clip.green_clip(f_d, s_d, cf, ta, chunk=4, fused=False)
    for x, y, name in zip(g1, g0, ("alpha", "trimap", "fg", "bg")):
        assert torch.equal(x, y), name
    # the second pass alone on soft, noisy working-resolution mattes (every interpolation phase and weight, borders)
    from video_unscreen_b200.unscreen.utils.fgfuncs import bgr2hsv_pixel
    th, tw = h // (2 if tag == "x2" else 4), w // (2 if tag == "x2" else 4)
    rng = np.random.default_rng(5)
    a_lo = rng.integers(0, 256, (n, th, tw), dtype=np.uint8)
    a_lo[:, : th // 3] = 0
    a_lo[:, th // 2:, : tw // 2] = 255
    col = cf.bg_color_bgr()
    hsv = bgr2hsv_pixel(col)
    half = np.array(ta.color_winsize) // 2
    lo, hi = np.clip(hsv - half, 10, 255), np.clip(hsv + half, 10, 255)
    deg = dev(np.array([0, 0, 1, 0, 0, 0, 0, 1, 0], np.uint8))
    alpha, fzb, mb, counts, fg, bg = env.ops.cf_alpha_up_fuzzy(dev(a_lo), h, w, f_d, lo, hi, alt_src=s_d, alt_flags=deg, bg_bgr=col)
    want_a = env.ops.resize_up(dev(a_lo), h, w, alt_src=s_d, alt_flags=deg)
    assert torch.equal(alpha, want_a)
    fz, cnt = env.ops.fuzzy_count(f_d, want_a, lo, hi)
    assert torch.equal(counts, cnt)
    bits = np.unpackbits(fzb.cpu().numpy(), axis=-1, bitorder="little").reshape(n, h, w)
    assert np.array_equal(bits, fz.cpu().numpy())
    sc = h // th
    mbits = np.unpackbits(mb.cpu().numpy(), axis=-1, bitorder="little").reshape(n, th, tw)
    assert np.array_equal(mbits, (want_a.cpu().numpy()[:, ::sc, ::sc] >= 128).astype(np.uint8))
    tile = torch.from_numpy(np.tile(col, (1, 4, 1))).cuda()
    fg0, bg0 = env.ops.get_fg(f_d, want_a, tile, 1, want_bg=True)
    assert torch.equal(fg, fg0) and torch.equal(bg, bg0)
    flags = env.ops.ratio_flags(cnt, 0.1)
    tri = env.ops.trimap_bits_packed(mb, fzb, flags, h, w, th, tw, 5)
    assert torch.equal(tri, env.ops.trimap_bits(want_a, th, tw, 5, fz, flags))
    assert torch.equal(env.ops.trimap_bits_packed(mb, None, None, h, w, th, tw, 3), env.ops.trimap_bits(want_a, th, tw, 3))


def test_green_clip_with_object_removal(env):
    """green.py:99-126 with remove_invalid_objects (:106-109) between colour filtering and the trimap, all on the device:
    every frame against the oracle chain"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import OBJ_CFGS
    n, h, w, L = 5, 108, 192, 96
    frames, segs = synth.green_clip(n, h, w, seed=31)
    rng = np.random.default_rng(2)
    for i in range(n):          # clutter for the removal to work on: small person-coloured patches away from the person
        for _ in range(6):
            y, x, s = rng.integers(0, h - 14), rng.integers(0, w - 14), rng.integers(3, 14)
            frames[i, y:y + s, x:x + s] = np.array(synth.PERSON, np.uint8)
    lb, lf, bgh = env.tables["x2"]
    cf = env.CF(input_long_side=L)
    cf.set_tables(lb, lf, bgh)
    ta = env.TA(input_long_side=L)
    cfg = OBJ_CFGS[0]
    alpha, tri, fg, bgo = (t.cpu().numpy() for t in env.clip.green_clip(dev(frames), dev(segs), cf, ta, chunk=2, remove_objects=cfg))
    col = cf.bg_color_bgr()
    removed = 0
    for i in range(n):
        a_cf, _, _ = R.cf_forward_predict(frames[i], segs[i], lb, lf, bgh, L)
        a_o = R.remove_invalid_objects(cfg, a_cf, segs[i])
        removed += int((a_o != a_cf).sum())
        assert np.array_equal(alpha[i], a_o), i
        assert np.array_equal(tri[i], R.generate_trimap_withbg(a_o, frames[i], col, L)), i
        patched = R.patch_bg(np.broadcast_to(col, frames[i].shape), frames[i], a_o, "lt128")
        assert np.array_equal(bgo[i], patched) and np.array_equal(fg[i], R.get_fg(frames[i], a_o, patched)), i
    assert removed > 0


@pytest.mark.parametrize("h,w,L,thr", [(96, 160, 80, 25), (384, 640, 160, 10), (100, 240, 60, 254), (64, 48, 32, 25), (150, 464, 116, 25)])
def test_bgstep_frames_fused_equals_staged(env, h, w, L, thr):
    """vu_bgstep_frames (TMA tiles with halo: gate + get_fg + trimap bits in one pass) against the separate kernels and the
    oracle: tile seams at 224 columns / 48 rows, image borders, thresholds, per-frame and shared backgrounds, working
    resolutions that are 2x, 4x and no exact fraction of the frame"""
    n = 5
    frames, masks, _ = synth.bgstep_clip(n, h, w, seed=h + w)
    rng = np.random.default_rng(7)
    masks[1] = rng.integers(0, 256, (h, w), dtype=np.uint8)          # grey-valued mask, differences everywhere
    frames[2, :3] = 255 - frames[2, :3]                               # differences on the image's top rows
    frames[3, :, -2:] = 0
    f_d, m_d = dev(frames), dev(masks)
    ta = env.TA(input_long_side=L)
    got = env.clip.bgstep_clip(f_d, m_d, ta, thr=thr, chunk=2, fused=True)
    want = env.clip.bgstep_clip(f_d, m_d, ta, thr=thr, chunk=2, fused=False)
    for g, wv, name in zip(got, want, ("bg", "alpha", "trimap", "fg")):
        assert torch.equal(g, wv), name
    bg_o = R.temporal_median(frames)
    for i in range(n):
        a_o = R.bgdiff_gate(frames[i], bg_o, masks[i], thr)
        assert np.array_equal(got[1][i].cpu().numpy(), a_o), i
        assert np.array_equal(got[3][i].cpu().numpy(), R.get_fg(frames[i], a_o, R.patch_bg(bg_o, frames[i], a_o, "eq0"))), i
    # a background per frame
    bgs = dev(np.stack([np.roll(bg_o, k, axis=1) for k in range(n)]))
    a1, f1, _ = env.ops.bgstep_frames(f_d, bgs, m_d, thr)
    a0 = env.ops.bgdiff_gate(f_d, bgs, m_d, thr)
    assert torch.equal(a1, a0) and torch.equal(f1, env.ops.get_fg(f_d, a0, bgs, 2))

"""oracle.refport against the committed golden vectors, which are outputs of
the unmodified reference (tests/golden/make_golden.py).  Bit-exact everywhere
except the stages that end in cv2's HSV2BGR, which is +-1 LSB by cv2's own
inconsistency."""
import numpy as np
import pytest

from conftest import maxdiff
from oracle import cvmodel as M
from oracle import refport as R

BGCOL = np.array([60, 200, 40], np.uint8)


def test_primitives(golden):
    p = golden("primitives")
    m = p["mask"]
    for k, n in [(3, 2), (3, 5), (4, 2), (5, 3), (7, 10), (5, 10)]:
        assert np.array_equal(R.dilate_mask(m, k, n), p[f"dilate_{k}_{n}"])
        assert np.array_equal(R.erode_mask(m, k, n), p[f"erode_{k}_{n}"])
    assert np.array_equal(R.get_outer_boundary(m), p["outer_boundary"])
    assert [R.exist_foreground(m, t) for t in (0.001, 0.4, 0.5, 0.6)] == list(p["exist_fg"])
    assert np.array_equal(R.is_pixel_inrange(p["img"], p["bgcolor"], (10, 100, 180)), p["inrange_color"])
    assert np.array_equal(R.is_pixel_inrange(p["img"], np.array([5, 9, 250], np.uint8), (60, 255, 255)), p["inrange_color_wide"])
    assert np.array_equal(R.is_pixel_inrange(p["img"], p["bgimg"], (20, 20, 120)), p["inrange_image"])
    assert np.array_equal(R.is_pixel_inrange(p["img"], p["bgimg"], (10, 100, 180)), p["inrange_image2"])
    for th, tw, h, w, L in p["target_sizes"]:
        assert R.get_target_size(h, w, L) == (th, tw)


def cf_tables(c, tag):
    lb = np.stack([R.gmm_lut(c[f"{tag}_bg{i}_means"], c[f"{tag}_bg{i}_covs"], c[f"{tag}_bg{i}_weights"]) for i in range(3)])
    lf = np.stack([R.gmm_lut(c[f"{tag}_fg{i}_means"], c[f"{tag}_fg{i}_covs"], c[f"{tag}_fg{i}_weights"]) for i in range(3)])
    bgh = R.bg_color_hsv([np.atleast_1d(c[f"{tag}_bg{i}_means"])[0] for i in range(3)])
    return lb, lf, bgh


@pytest.mark.parametrize("tag", ["x2", "x4", "frac", "portrait"])
def test_colorfilter(golden, tag):
    c = golden("colorfilter")
    L = int(c[f"{tag}_L"])
    lb, lf, bgh = cf_tables(c, tag)
    fr, seg = c[f"{tag}_frame"], c[f"{tag}_seg"]
    a, bgimg, _ = R.cf_forward_predict(fr, seg, lb, lf, bgh, L)
    assert np.array_equal(a, c[f"{tag}_alpha_pred"])
    assert (a > 128).sum() > 1000 and (a < 128).sum() > 1000
    assert maxdiff(bgimg, c[f"{tag}_bgimg"]) <= 1
    a2, _, _ = R.cf_forward_predict(c[f"{tag}_frame2"], c[f"{tag}_seg2"], lb, lf, bgh, L)
    assert np.array_equal(a2, c[f"{tag}_alpha_pred2"])
    hsv = M.bgr2hsv(fr)
    th, tw = R.get_target_size(*fr.shape[:2], L)
    hl, sl = M.resize_linear(hsv, tw, th), M.resize_linear(seg, tw, th)
    raw = R.alpha_from_luts(hl, lb, lf)
    assert np.array_equal(raw, c[f"{tag}_alpha_raw"])
    assert np.array_equal(R.cf_postprocess(raw, sl), c[f"{tag}_alpha_post"])
    assert np.array_equal(R.get_color_prior(hl, sl < 128, 30)[0], c[f"{tag}_prior30"])
    assert np.array_equal(R.get_color_prior(hl, sl < 128, 6)[0], c[f"{tag}_prior6"])


def test_colorfilter_early_outs(golden):
    c = golden("colorfilter")
    fr = c["early_frame"]
    z = np.zeros(fr.shape[:2], np.uint8)
    a, b, _ = R.cf_forward_predict(fr, z, None, None, None, 48)
    assert np.array_equal(a, c["early_nofg_alpha"]) and np.array_equal(b, c["early_nofg_bg"])
    a, b, _ = R.cf_forward_predict(fr, z + 255, None, None, None, 48)
    assert np.array_equal(a, c["early_nobg_alpha"]) and np.array_equal(b, c["early_nobg_bg"])


@pytest.mark.parametrize("tag", ["x2", "x4", "frac", "portrait", "up"])
def test_trimap(golden, tag):
    t = golden("trimap")
    L = int(t[f"{tag}_L"])
    fr = t[f"{tag}_frame"]
    assert np.array_equal(R.generate_trimap(t[f"{tag}_mask"], L), t[f"{tag}_plain"])
    assert np.array_equal(R.generate_trimap(t[f"{tag}_soft"], L), t[f"{tag}_plain_soft"])
    assert np.array_equal(R.generate_trimap_withbg(t[f"{tag}_soft"], fr, BGCOL, L), t[f"{tag}_withcolor"])
    assert np.array_equal(R.generate_trimap_withbg(t[f"{tag}_soft"], fr, t[f"{tag}_bgimg"], L), t[f"{tag}_withimage"])
    assert np.array_equal(R.generate_trimap_withbg(t[f"{tag}_leak"], fr, BGCOL, L), t[f"{tag}_leak_withcolor"])
    assert np.array_equal(R.generate_trimap_withbg(t[f"{tag}_ring"], fr, BGCOL, L), t[f"{tag}_ring_withcolor"])
    assert np.array_equal(R.generate_trimap_withbg(t[f"{tag}_ring"], fr, t[f"{tag}_bgimg"], L), t[f"{tag}_ring_withimage"])
    assert set(np.unique(t[f"{tag}_plain"])) <= {0, 128, 255}


def test_trimap_empty(golden):
    t = golden("trimap")
    z = np.zeros((40, 64), np.uint8)
    assert np.array_equal(R.generate_trimap_withbg(z, np.zeros((40, 64, 3), np.uint8), np.array([1, 2, 3], np.uint8), 64),
                          t["empty_withcolor"])


def test_composite(golden):
    k = golden("composite")
    fr, al, bg = k["frame"], k["alpha"], k["bg"]
    # +-1: end in HSV2BGR
    assert maxdiff(R.get_fg(fr, al, bg), k["get_fg"]) <= 1
    assert maxdiff(R.get_bg(al, bg), k["get_bg"]) <= 1
    assert maxdiff(R.get_fg(fr, al, R.patch_bg(bg, fr, al, "lt128")), k["green_patch_fg"]) <= 1
    assert maxdiff(R.get_fg(fr, al, R.patch_bg(bg, fr, al, "eq0")), k["bg_patch_fg"]) <= 1
    assert (R.get_fg(fr, al, bg) != k["get_fg"]).mean() < 1e-3
    # exact: float64 sequences
    assert np.array_equal(R.get_fg_naive(fr, al), k["get_fg_naive"])
    assert np.array_equal(R.fuse_fgbg(fr, bg, al), k["fuse_fgbg"])
    assert np.array_equal(R.composite_fgbg(fr, al, k["newbg"]), k["composite"])
    assert np.array_equal(R.composite_fgbg(fr, al, k["newbg"], True), k["composite_ext"])
    assert np.array_equal(R.composite_fgbg(fr, al, k["tallbg"]), k["composite_tall"])
    assert np.array_equal(R.replace_blend(fr, np.stack([al] * 3, -1), bg), k["replace"])
    assert np.array_equal(R.replace_blend(fr, al, bg), k["replace"])
    assert np.array_equal(R.fuse_bg(bg, k["bg_always"], 0.1), k["fused_bg"])
    assert np.array_equal(R.bgdiff_gate(fr, k["near_bg"], al, 25), k["gate"])
    assert np.array_equal(R.binarise_dilate(al), k["binarise_dilate"])


def test_temporal(golden):
    t = golden("temporal")
    bg, ma = R.masked_temporal_mean(t["frames"], t["masks"])
    assert np.array_equal(bg, t["mean_bg"]) and np.array_equal(ma, t["mean_mask_always"])
    assert np.array_equal(R.temporal_median(t["frames"]), t["median_even"])
    assert np.array_equal(R.temporal_median(t["frames"][:23]), t["median_odd"])
    assert np.array_equal(R.temporal_median(t["rnd"]), t["median_rnd"])


def test_geometry(golden):
    """replacement-path geometry: shift_fg bit-exact; rescale_fg within the documented tie tolerance of cv2's IPP cubic
    (oracle/cvmodel.py:resize_cubic_crop)."""
    g = golden("geometry")
    fg, m3, bg = g["fg"], g["mask3"], g["bg"]
    for i, (dx, dy) in enumerate(g["shifts"]):
        assert np.array_equal(R.shift_fg(fg, dx, dy), g[f"shift_fg_{i}"]), (dx, dy)
        assert np.array_equal(R.shift_fg(m3[..., 0], dx, dy), g[f"shift_mask_{i}"]), (dx, dy)
    for tag, sc in (("12", 1.2), ("11", 1.1)):
        for mine, ref in ((R.rescale_fg(fg, sc), g[f"rescale_fg_{tag}"]), (R.rescale_fg(m3[..., 0], sc), g[f"rescale_mask_{tag}"])):
            d = np.abs(mine.astype(int) - ref.astype(int))
            assert d.max() <= 1 and (d > 0).mean() <= 1e-4
    d = np.abs(R.replace_frame(fg, m3, bg, 3, -2, 1.2).astype(int) - g["replace_frame"].astype(int))
    assert d.max() <= 2 and (d > 0).mean() <= 1e-3


def test_color_correct(golden):
    """color_correct outputs of the unmodified reference (green-clip frames, three working resolutions, two background
    colours): bit-exact."""
    g = golden("geometry")
    for ci, col in enumerate(g["cc_colors"]):
        for L in g["cc_long_sides"]:
            for i in range(2):
                got = R.color_correct(g["cc_frames"][i], g["cc_alpha"][i], col, target_long_side=int(L))
                assert np.array_equal(got, g[f"cc_{ci}_{int(L)}_{i}"]), (ci, int(L), i)


@pytest.mark.parametrize("i", range(5))
def test_background_agent_mean_pcov(golden, i):
    """BackgroundAgent.forward (bgmodel/agent.py:159-208): 'pcov' bit-exact; 'mean' ends in cv2's HSV2BGR of a constant
    image, whose scalar tail (the last columns of every row) rounds where the SIMD body truncates: <= 1 LSB
    there before, <= 2 LSB after the final resize."""
    g = golden("bgmodel")
    h, w, L, kind = (int(v) for v in g["cases"][i])
    img, m = g[f"img_{i}"], g[f"mask_{i}"]
    assert np.array_equal(R.background_forward(img, m, "pcov", input_long_side=L), g[f"pcov_{i}"])
    got = R.background_forward(img, m, "mean", input_long_side=L)
    d = np.abs(got.astype(int) - g[f"mean_{i}"].astype(int))
    assert d.max() <= 2      # where: the hole pixels of the scalar-tail columns (image width mod the SIMD width)
    # the early-outs
    assert R.background_forward(img, np.zeros_like(m), "mean") is img
    z = R.background_forward(img, np.full_like(m, 255), "pcov")
    assert z.dtype == np.float64 and z.shape == img.shape and not z.any()
    with pytest.raises(NameError):
        R.background_forward(img, m, "telea")


@pytest.mark.parametrize("i", range(5))
def test_regionfill(golden, i):
    """regionfill (utils/region_fill.py:7-63) at factor 1 (bg.py:79) and 0.5 (bgmodel/agent.py:150): the same sparse
    system handed to the same scipy solver, the float64 cv2.resize calls modelled (cvmodel.resize_linear_f64): equal to
    1e-9.  BackgroundAgent 'rf' ends in HSV2BGR: <= 2 LSB after the final resize, as for 'mean'."""
    g = golden("regionfill")
    h, w, L, kind = (int(v) for v in g["cases"][i])
    img, m = g[f"img_{i}"], g[f"mask_{i}"]
    for f in (1.0, 0.5):
        got = R.regionfill(img[:, :, i % 3], m > 0, f)
        assert np.abs(got - g[f"fill_{i}_{int(f * 10)}"]).max() <= 1e-9, f
    d = np.abs(R.background_forward(img, m, "rf", input_long_side=L).astype(int) - g[f"rf_{i}"].astype(int))
    assert d.max() <= 2
    assert R.regionfill(img[:, :, 0], np.zeros_like(m)).dtype == np.uint8        # :8-9: the input's copy


@pytest.mark.parametrize("i", range(6))
def test_remove_invalid_objects(golden, i):
    """the closed-form contour model (oracle/refport.py:contour_objects) against the reference's cv2.findContours /
    contourArea / drawContours loop: bit-exact for the three configurations the repository ships"""
    sys_path_golden()
    from make_golden import OBJ_CFGS
    g = golden("objects")
    a, seg = g[f"alpha_{i}"], g[f"seg_{i}"]
    for c, cfg in enumerate(OBJ_CFGS):
        assert np.array_equal(R.remove_invalid_objects(cfg, a), g[f"self_{i}_{c}"]), c
        assert np.array_equal(R.remove_invalid_objects(cfg, a, seg), g[f"seg_{i}_{c}"]), c
    assert any((g[f"self_{i}_{c}"] != a).any() for c in range(3))      # something was removed


def sys_path_golden():
    import os
    import sys
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if p not in sys.path:
        sys.path.insert(0, p)

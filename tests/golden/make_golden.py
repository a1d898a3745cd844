"""Generate the committed golden fixtures by executing the UNMODIFIED reference
(/root/reference) under a non-invasive import shim, on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs + reference outputs) and MANIFEST.json with
the library versions the outputs were produced with -- the reference pins none
of its dependencies, so parity is defined against exactly these versions.
The inline lines of tools/unscreen/{bg,bg_offline}.py and tools/replace/
replace.py cannot be imported (scripts with hard-coded data roots); they are
restated here with the reference's own utils, citing the lines.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))


def load_reference(root="/root/reference"):
    np.float = float  # removed numpy aliases the reference still uses
    np.int = int      # (np.bool still exists in numpy 2: leave it alone)
    for n in ("mmcv", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    import unscreen.utils as U
    from unscreen.colorfiltering import ColorFilteringAgent
    from unscreen.trimap import TrimapAgent
    from unscreen.bgmodel import BackgroundAgent
    return U, ColorFilteringAgent, TrimapAgent, BackgroundAgent


def main():
    import cv2
    import sklearn
    import torch
    from video_unscreen_b200 import synth

    U, CF, TA, BA = load_reference()
    rng = np.random.default_rng(1234)
    out = {}

    # ---- primitives through the reference's wrappers -------------------
    prim = {}
    m = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    prim["mask"] = m
    for k, n in [(3, 2), (3, 5), (4, 2), (5, 3), (7, 10), (5, 10)]:
        prim[f"dilate_{k}_{n}"] = U.dilate_mask(m, k, n)
        prim[f"erode_{k}_{n}"] = U.erode_mask(m, k, n)
    prim["outer_boundary"] = U.get_outer_boundary(m)
    prim["exist_fg"] = np.array([U.exist_foreground(m, t) for t in (0.001, 0.4, 0.5, 0.6)])
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    bgi = np.clip(img.astype(np.int16) + rng.integers(-30, 31, img.shape), 0, 255).astype(np.uint8)
    prim["img"] = img
    prim["bgimg"] = bgi
    prim["bgcolor"] = np.array([60, 200, 40], np.uint8)
    prim["inrange_color"] = U.is_pixel_inrange(img, prim["bgcolor"], (10, 100, 180))
    prim["inrange_color_wide"] = U.is_pixel_inrange(img, np.array([5, 9, 250], np.uint8), (60, 255, 255))
    prim["inrange_image"] = U.is_pixel_inrange(img, bgi, (20, 20, 120))
    prim["inrange_image2"] = U.is_pixel_inrange(img, bgi, (10, 100, 180))
    prim["target_sizes"] = np.array([list(U.get_target_size(h, w, L)) + [h, w, L] for h, w, L in
                                     [(1080, 1920, 960), (2160, 3840, 960), (1920, 1080, 960),
                                      (720, 1280, 960), (97, 131, 64), (333, 517, 200), (480, 480, 100)]])
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **prim)

    # ---- colour filtering ----------------------------------------------
    cf = {}
    for tag, (h, w, L) in {"x2": (270, 480, 240), "x4": (360, 640, 160), "frac": (250, 333, 200),
                           "portrait": (480, 270, 240)}.items():
        frame, seg = synth.green_frame(h, w, t=3, n=20, seed=5)
        agent = CF(input_long_side=L)
        np.random.seed(0)
        a_fit, bg_fit, _ = agent.forward(frame.copy(), seg.copy(), 3)
        a_pred, bg_pred, _ = agent.forward(frame.copy(), seg.copy(), 0)
        frame2, seg2 = synth.green_frame(h, w, t=9, n=20, seed=5)
        a_pred2, _, _ = agent.forward(frame2.copy(), seg2.copy(), 0)
        cf[f"{tag}_frame"] = frame
        cf[f"{tag}_seg"] = seg
        cf[f"{tag}_frame2"] = frame2
        cf[f"{tag}_seg2"] = seg2
        cf[f"{tag}_L"] = np.array(L)
        cf[f"{tag}_alpha_fit"] = a_fit
        cf[f"{tag}_alpha_pred"] = a_pred
        cf[f"{tag}_alpha_pred2"] = a_pred2
        cf[f"{tag}_bgimg"] = bg_pred
        for nm, gm in (("bg", agent.bg_gmms), ("fg", agent.fg_gmms)):
            for c in range(3):
                cf[f"{tag}_{nm}{c}_means"] = gm[c].means_.squeeze()
                cf[f"{tag}_{nm}{c}_covs"] = gm[c].covariances_.squeeze()
                cf[f"{tag}_{nm}{c}_weights"] = gm[c].weights_.squeeze()
        # intermediate pieces for finer-grained oracle checks
        hsv = cv2.cvtColor(frame, cv2.COLOR_BGR2HSV)
        th, tw = U.get_target_size(h, w, L)
        hsv_lo = cv2.resize(hsv, (tw, th))
        seg_lo = cv2.resize(seg, (tw, th))
        a_raw, _ = agent.get_alpha_by_gmm(hsv_lo)
        cf[f"{tag}_alpha_raw"] = a_raw
        cf[f"{tag}_alpha_post"] = agent.postprocess(a_raw.copy(), seg_lo)
        cf[f"{tag}_prior30"] = agent.get_color_prior(hsv_lo, seg_lo < 128, 30)
        cf[f"{tag}_prior6"] = agent.get_color_prior(hsv_lo, seg_lo < 128, 6)
    # early-outs (agent.py:303-307)
    frame, seg = synth.green_frame(64, 96, seed=2)
    agent = CF(input_long_side=48)
    e1 = agent.forward(frame, np.zeros_like(seg), 0)
    e2 = agent.forward(frame, np.full_like(seg, 255), 0)
    cf["early_frame"] = frame
    cf["early_nofg_alpha"], cf["early_nofg_bg"] = e1[0], e1[1]
    cf["early_nobg_alpha"], cf["early_nobg_bg"] = e2[0], e2[1]
    np.savez_compressed(os.path.join(HERE, "colorfilter.npz"), **cf)

    # ---- trimap -----------------------------------------------------------
    tri = {}
    for tag, (h, w, L) in {"x2": (270, 480, 240), "x4": (360, 640, 160), "frac": (250, 333, 200),
                           "portrait": (480, 270, 240), "up": (120, 200, 320)}.items():
        frame, seg = synth.green_frame(h, w, t=2, n=10, seed=8)
        ta = TA(input_long_side=L)
        soft = seg.copy()
        soft[::7, ::5] = rng.integers(0, 256, soft[::7, ::5].shape, dtype=np.uint8)
        bgcol = np.array([60, 200, 40], np.uint8)
        bgimg = np.clip(np.full((h, w, 3), (60, 200, 40), np.int16) + rng.integers(-5, 6, (h, w, 3)), 0, 255).astype(np.uint8)
        tri[f"{tag}_frame"], tri[f"{tag}_mask"], tri[f"{tag}_soft"] = frame, seg, soft
        tri[f"{tag}_bgimg"] = bgimg
        tri[f"{tag}_L"] = np.array(L)
        tri[f"{tag}_plain"] = ta.forward(seg.copy())
        tri[f"{tag}_plain_soft"] = ta.forward(soft.copy())
        tri[f"{tag}_withcolor"] = ta.forward(soft.copy(), frame.copy(), bgcol)
        tri[f"{tag}_withimage"] = ta.forward(soft.copy(), frame.copy(), bgimg.copy())
        # a mask that leaks into the screen => fuzzy ratio > 0.1 => plain branch
        leak = U.dilate_mask(seg, 5, 10)
        tri[f"{tag}_leak"] = leak
        tri[f"{tag}_leak_withcolor"] = ta.forward(leak.copy(), frame.copy(), bgcol)
        # a thin ring of screen pixels inside the mask => ratio < 0.1 => ensemble branch
        ring = U.dilate_mask(seg, 3, 4)
        ring[::9, ::11] = np.maximum(ring[::9, ::11], (rng.integers(0, 8, ring[::9, ::11].shape) == 0).astype(np.uint8) * 200)
        tri[f"{tag}_ring"] = ring
        tri[f"{tag}_ring_withcolor"] = ta.forward(ring.copy(), frame.copy(), bgcol)
        tri[f"{tag}_ring_withimage"] = ta.forward(ring.copy(), frame.copy(), bgimg.copy())
        print(tag, "ensemble-vs-plain differing px:", int((tri[f"{tag}_ring_withcolor"] != ta.forward(ring.copy())).sum()))
    tri["empty_withcolor"] = TA(input_long_side=64).forward(np.zeros((40, 64), np.uint8), np.zeros((40, 64, 3), np.uint8), np.array([1, 2, 3], np.uint8))
    np.savez_compressed(os.path.join(HERE, "trimap.npz"), **tri)

    # ---- compositing -------------------------------------------------------
    comp = {}
    h, w = 180, 320
    frame, seg = synth.green_frame(h, w, t=1, n=10, seed=11)
    alpha = rng.integers(0, 256, (h, w), dtype=np.uint8)
    alpha[:40] = 0
    alpha[-40:] = 255
    bg = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    newbg = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    comp.update(frame=frame, alpha=alpha, bg=bg, newbg=newbg)
    comp["get_fg"] = U.get_fg(frame.copy(), alpha.copy(), bg.copy())
    comp["get_bg"] = U.get_bg(alpha.copy(), bg.copy())
    comp["get_fg_naive"] = U.get_fg_naive(frame.copy(), alpha.copy())
    comp["fuse_fgbg"] = U.fuse_fgbg(frame.copy(), bg.copy(), alpha.copy())
    comp["composite"] = U.composite_fgbg(frame.copy(), alpha.copy(), newbg.copy())
    comp["composite_ext"] = U.composite_fgbg(frame.copy(), alpha.copy(), newbg.copy(), extend=True)
    tall = rng.integers(0, 256, (400, 150, 3), dtype=np.uint8)
    comp["tallbg"] = tall
    comp["composite_tall"] = U.composite_fgbg(frame.copy(), alpha.copy(), tall.copy())
    # tools/replace/replace.py:74-76
    mask3 = np.stack([alpha] * 3, -1)
    nbm = mask3.astype(np.float) / 255
    comp["replace"] = (frame.astype(np.float) * nbm + bg.astype(np.float) * (1 - nbm)).astype(np.uint8)
    # tools/unscreen/green.py:125-126
    b2 = bg.copy()
    b2[alpha < 128] = frame[alpha < 128]
    comp["green_patch_fg"] = U.get_fg(frame.copy(), alpha.copy(), b2)
    # tools/unscreen/bg_offline.py:171-172
    b3 = bg.copy()
    b3[alpha == 0] = frame[alpha == 0]
    comp["bg_patch_fg"] = U.get_fg(frame.copy(), alpha.copy(), b3)
    # tools/unscreen/bg_offline.py:150-160
    always = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    comp["bg_always"] = always
    beta = 0.1
    fused = ((bg.copy().astype(np.float32) * beta + (1 - beta) * always.copy().astype(np.float32))).astype(np.uint8)
    comp["fused_bg"] = fused
    near = np.clip(frame.astype(np.int16) + rng.integers(-40, 41, frame.shape), 0, 255).astype(np.uint8)
    comp["near_bg"] = near
    raw = (np.abs(frame.astype(np.float32) - near.copy().astype(np.float32))).astype(np.uint8)
    g = cv2.cvtColor(raw, cv2.COLOR_BGR2GRAY)
    g[g > 25] = 255
    g = np.clip((g.astype(np.float32)), 0, 255).astype(np.uint8)
    g = U.dilate_mask(g, 4, 2)
    comp["gate"] = alpha.copy() * (g // 255)
    # tools/unscreen/bg.py:74-77
    ab = alpha.copy()
    ab[ab > 128] = 255
    ab[ab <= 128] = 0
    comp["binarise_dilate"] = U.dilate_mask(ab, 3, 2)
    np.savez_compressed(os.path.join(HERE, "composite.npz"), **comp)

    # ---- temporal ----------------------------------------------------------
    tmp = {}
    n, h, w = 24, 60, 80
    frames, masks, _ = synth.bgstep_clip(n, h, w, seed=3)
    tmp["frames"], tmp["masks"] = frames, masks
    # tools/unscreen/bg_offline.py:106-125 (dead code there), single-channel masks stacked to 3
    raw = np.zeros((h, w, 3))
    cont = np.zeros((h, w, 3))
    for fid in range(n):
        sm = np.stack([masks[fid]] * 3, -1)
        sm = U.dilate_mask(sm, 3, 2)
        no_mask = frames[fid] * (np.ones_like(sm) - sm // 255).astype(np.float32)
        cont += (sm < 250).astype(np.float32)
        raw += no_mask.astype(np.float32)
    mask_always = ((cont <= 10) * 255).astype(np.uint8)
    cc = cont.copy()
    cc[cont == 0] = 1
    bg_always = np.clip((raw / cc), 0, 255).astype(np.uint8)
    bg_always[mask_always == 255] = 0
    tmp["mean_bg"] = bg_always
    tmp["mean_mask_always"] = mask_always[..., 0]
    # NEW SPEC (no reference code): np.median
    tmp["median_even"] = np.median(frames, axis=0).astype(np.uint8)
    tmp["median_odd"] = np.median(frames[:23], axis=0).astype(np.uint8)
    rnd = synth.random_clip(10, 33, 47, seed=4)
    tmp["rnd"] = rnd
    tmp["median_rnd"] = np.median(rnd, axis=0).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "temporal.npz"), **tmp)

    geometry(U)
    bgmodel(BA)
    objects(U)
    write_manifest()


def geometry(U):
    """replacement-path geometry: shift_fg / rescale_fg of the reference and the frame of
    tools/replace/replace.py:69-76 restated with them.  `python make_golden.py geometry` regenerates this file alone."""
    rng = np.random.default_rng(11)
    geo = {}
    fg = rng.integers(0, 256, (54, 96, 3), dtype=np.uint8)
    m3 = np.repeat(rng.integers(0, 256, (54, 96, 1), dtype=np.uint8), 3, axis=2)
    bg = rng.integers(0, 256, (54, 96, 3), dtype=np.uint8)
    geo["fg"], geo["mask3"], geo["bg"] = fg, m3, bg
    shifts = np.array([[3, -2], [0.5, 0.5], [-6.37, 2.81], [0, 0], [40.25, -30.75]], np.float64)
    geo["shifts"] = shifts
    for i, (dx, dy) in enumerate(shifts):
        geo[f"shift_fg_{i}"] = U.shift_fg(fg, dx=dx, dy=dy)
        geo[f"shift_mask_{i}"] = U.shift_fg(m3[..., 0], dx=dx, dy=dy)
    for tag, sc in (("12", 1.2), ("11", 1.1)):
        geo[f"rescale_fg_{tag}"] = U.rescale_fg(fg, scale_factor=sc)
        geo[f"rescale_mask_{tag}"] = U.rescale_fg(m3[..., 0], scale_factor=sc)
    f = U.rescale_fg(U.shift_fg(fg, dx=3, dy=-2), scale_factor=1.2)
    m = U.rescale_fg(U.shift_fg(m3, dx=3, dy=-2), scale_factor=1.2)
    nb = m.astype(np.float64) / 255
    geo["replace_frame"] = (f.astype(np.float64) * nb + bg.astype(np.float64) * (1 - nb)).astype(np.uint8)
    # color_correct (imgprocess.py:263-300) on frames of the synthetic green clip, working resolutions 1x, 1/2, ragged
    from video_unscreen_b200 import synth
    frames, segs = synth.green_clip(2, 108, 192, seed=12)
    alpha = np.minimum(segs, np.where(rng.random(segs.shape) < 0.3, rng.integers(0, 256, segs.shape), 255).astype(np.uint8))
    geo["cc_frames"], geo["cc_alpha"] = frames, alpha
    geo["cc_colors"] = np.array([[60, 200, 40], [200, 30, 30]], np.uint8)
    geo["cc_long_sides"] = np.array([96, 192, 70])
    for ci, col in enumerate(geo["cc_colors"]):
        for L in geo["cc_long_sides"]:
            for i in range(2):
                geo[f"cc_{ci}_{int(L)}_{i}"] = U.color_correct(frames[i].copy(), alpha[i].copy(), col.copy(), target_long_side=int(L))
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **geo)


def bgmodel_case(h, w, seed, kind):
    """image + person-like mask for the BackgroundAgent fixtures: blurred random image; kind 1 adds isolated mask pixels,
    kind 2 a block touching the image corner (the hole's box is clipped by the image)"""
    import cv2
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3)
    yy, xx = np.mgrid[0:h, 0:w]
    m = ((((xx - w * 0.5) / (w * 0.2)) ** 2 + ((yy - h * 0.55) / (h * 0.3)) ** 2) <= 1).astype(np.uint8) * 255
    if kind == 1:
        m[rng.random((h, w)) < 0.002] = 255
    if kind == 2:
        m[:h // 3, :w // 4] = 255
    return img, m


def bgmodel(BA):
    """BackgroundAgent.forward, methods 'mean' and 'pcov' (SURVEY.md 8f rank 4).  `python make_golden.py bgmodel`."""
    out = {}
    cases = [(108, 192, 96, 0), (108, 192, 96, 1), (135, 240, 270, 2), (150, 100, 75, 0), (96, 160, 160, 1)]
    out["cases"] = np.array(cases)
    for i, (h, w, L, kind) in enumerate(cases):
        img, m = bgmodel_case(h, w, 100 + i, kind)
        out[f"img_{i}"], out[f"mask_{i}"] = img, m
        ag = BA(input_long_side=L)
        for method in ("mean", "pcov"):
            out[f"{method}_{i}"] = ag.forward(img.copy(), m.copy(), method)
    np.savez_compressed(os.path.join(HERE, "bgmodel.npz"), **out)


def regionfill(U, BA):
    """regionfill (utils/region_fill.py) at both factors its callers use, and BackgroundAgent.forward(method='rf')
    (SURVEY.md 8f rank 4).  `python make_golden.py regionfill`.  The comparison of the device results with
    these is a tolerance (tests/test_gpu_parity.py)."""
    out = {}
    cases = [(108, 192, 96, 0), (150, 100, 75, 1), (135, 240, 270, 2), (200, 130, 90, 1), (121, 187, 187, 0)]
    out["cases"] = np.array(cases)
    for i, (h, w, L, kind) in enumerate(cases):
        img, m = bgmodel_case(h, w, 300 + i, kind)
        out[f"img_{i}"], out[f"mask_{i}"] = img, m
        for f in (1.0, 0.5):
            out[f"fill_{i}_{int(f * 10)}"] = U.regionfill(img[:, :, i % 3].copy(), m > 0, f)
        out[f"rf_{i}"] = BA(input_long_side=L).forward(img.copy(), m.copy(), "rf")
    np.savez_compressed(os.path.join(HERE, "regionfill.npz"), **out)


OBJ_CFGS = [{'objectremoval': {'score_map_center': {'landscape': [0.5, 0.5], 'portrait': [0.6, 0.5]}, 'saliency_thr': t, 'consensus_thr': 0.5}}
            for t in (0.005, 0.00001, 0.001)]      # configs/green.json, configs/bg.json, the function's default


def objects_case(h, w, seed):
    """blobby matte (thresholded low-pass noise, grey values, sometimes salted with isolated pixels) + an unrelated blobby
    segmentation mask: components of every size, holes, islands in holes, contours with area around 100"""
    import cv2
    rng = np.random.default_rng(seed)
    x = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), float(rng.choice([2, 4, 8])))
    on = x > np.quantile(x, rng.choice([0.5, 0.7, 0.9]))
    a = on.astype(np.uint8) * rng.integers(1, 256, (h, w), dtype=np.uint8)
    if seed % 3 == 0:
        a[rng.random((h, w)) < 0.01] = 200
    if seed % 4 == 1:
        a[h // 4: h // 2, w // 4: w // 2] = 0           # a big hole with whatever is left inside
        a[h // 3, w // 3] = 99
    seg = (cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 8) > 0.5).astype(np.uint8) * 255
    return a, seg


def objects(U):
    """remove_invalid_objects.  `python make_golden.py objects`."""
    out = {}
    sizes = [(120, 200), (200, 120), (135, 240), (96, 96), (64, 300), (180, 320)]
    out["sizes"] = np.array(sizes)
    for i, (h, w) in enumerate(sizes):
        a, seg = objects_case(h, w, 40 + i)
        out[f"alpha_{i}"], out[f"seg_{i}"] = a, seg
        for c, cfg in enumerate(OBJ_CFGS):
            out[f"self_{i}_{c}"] = U.remove_invalid_objects(cfg, a.copy())
            out[f"seg_{i}_{c}"] = U.remove_invalid_objects(cfg, a.copy(), seg.copy())
    np.savez_compressed(os.path.join(HERE, "objects.npz"), **out)


def write_manifest():
    import cv2
    import sklearn
    import torch
    manifest = {
        "generator": "tests/golden/make_golden.py",
        "reference": "AnyiRao/video_unscreen @ /root/reference (unmodified, import shim)",
        "versions": {"python": sys.version.split()[0], "numpy": np.__version__, "cv2": cv2.__version__,
                     "torch": torch.__version__, "sklearn": sklearn.__version__},
        "cpu_capability": torch.backends.cpu.get_cpu_capability(),
        "files": sorted(f for f in os.listdir(HERE) if f.endswith(".npz")),
    }
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    if sys.argv[1:] == ["geometry"]:
        geometry(load_reference()[0])
        write_manifest()
    elif sys.argv[1:] == ["objects"]:
        objects(load_reference()[0])
        write_manifest()
    elif sys.argv[1:] == ["regionfill"]:
        ref = load_reference()
        regionfill(ref[0], ref[3])
        write_manifest()
    elif sys.argv[1:] == ["bgmodel"]:
        bgmodel(load_reference()[3])
        write_manifest()
    else:
        main()

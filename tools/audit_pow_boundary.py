#!/usr/bin/env python
"""How exposed is the matte to the one irreproducible operation of the reference, torch.pow(x, 1/3.) (SLEEF's 1-ULP powf,
colorfiltering/agent.py:250-251)?  The matte is a pure function of a pixel's (H, S, V): for EVERY one of the 180 x 256 x
256 inputs this evaluates the oracle with the cube roots one float32 ULP down / exact / one ULP up and counts the
inputs whose uint8 result depends on it; then the share of such pixels on the frames of the BASELINE clips (synthetic
green-screen frames, working resolution).  VERDICT r1, weak 5.  CPU only:  python tools/audit_pow_boundary.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cvmodel as M  # noqa: E402
from oracle import refport as R  # noqa: E402
from video_unscreen_b200 import synth  # noqa: E402


def alpha_variants(hsv, lb, lf):
    """(alpha with the correctly rounded cube roots, set of inputs whose alpha changes when either root moves by one ULP)"""
    f = np.float32
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    bg = ((f(1.0) * lb[0][h]).astype(f) * lb[1][s]).astype(f) * lb[2][v]
    fg = ((f(1.0) * lf[0][h]).astype(f) * lf[1][s]).astype(f) * lf[2][v]
    b0, f0 = R.pow_third(bg.astype(f)), R.pow_third(fg.astype(f))

    def alpha(b, g):
        den = ((b + g).astype(f) + f(1e-6)).astype(f)
        return np.clip(((g / den).astype(f) * f(255)).astype(f), 0, 255).astype(np.uint8)
    base = alpha(b0, f0)
    unstable = np.zeros(base.shape, bool)
    for db in (-1, 0, 1):
        for dg in (-1, 0, 1):
            if db == 0 and dg == 0:
                continue
            b = np.nextafter(b0, f(np.inf) if db > 0 else f(-np.inf)) if db else b0
            g = np.nextafter(f0, f(np.inf) if dg > 0 else f(-np.inf)) if dg else f0
            unstable |= alpha(b.astype(f), g.astype(f)) != base
    return base, unstable


def main():
    c = np.load(os.path.join(ROOT, "tests", "golden", "colorfilter.npz"))
    out = {}
    for tag in ("x2", "x4"):
        lb = np.stack([R.gmm_lut(c[f"{tag}_bg{i}_means"], c[f"{tag}_bg{i}_covs"], c[f"{tag}_bg{i}_weights"]) for i in range(3)])
        lf = np.stack([R.gmm_lut(c[f"{tag}_fg{i}_means"], c[f"{tag}_fg{i}_covs"], c[f"{tag}_fg{i}_weights"]) for i in range(3)])
        unstable_lut = np.zeros((180, 256, 256), bool)
        for h in range(180):
            hsv = np.stack(np.meshgrid(np.array([h]), np.arange(256), np.arange(256), indexing="ij"), -1).astype(np.int64)[0]
            _, u = alpha_variants(hsv, lb, lf)
            unstable_lut[h] = u
        hh, ww = (1080, 1920) if tag == "x2" else (2160, 3840)
        px = tot = 0
        for t in range(3):
            frame, _ = synth.green_frame(hh, ww, t=t, n=6, seed=0)
            lo = M.resize_linear(M.bgr2hsv(frame), 960, 540)
            px += int(unstable_lut[lo[..., 0], lo[..., 1], lo[..., 2]].sum())
            tot += lo.shape[0] * lo.shape[1]
        out[tag] = {"inputs": 180 * 256 * 256, "inputs_depending_on_a_1ulp_cube_root": int(unstable_lut.sum()),
                    "share_of_inputs": float(unstable_lut.mean()), "baseline_clip_pixels_checked": tot,
                    "baseline_clip_pixels_depending": px, "share_of_pixels": px / tot}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

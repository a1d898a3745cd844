#!/usr/bin/env python
"""Single-image background inpainting on one 1080p frame: BackgroundAgent 'mean' / 'pcov' / 'rf'
through the numpy API, and bg.py:79's full-resolution regionfill of the three planes; the oracle port (numpy / scipy)
timed beside them.  One JSON line per measurement.  python tools/bench_bgmodel.py [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from make_golden import bgmodel_case
    from oracle import refport as R
    from video_unscreen_b200 import ops
    from video_unscreen_b200.unscreen.bgmodel import BackgroundAgent
    h, w = 1080, 1920
    img, m = bgmodel_case(h, w, 7, 0)
    ag = BackgroundAgent()

    def wall(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3, out

    for method in ("mean", "pcov", "rf"):
        ms, got = wall(lambda: ag.forward(img, m, method), args.reps)
        rec = {"what": f"BackgroundAgent.forward(method='{method}'), 1080p numpy frame in / out", "gpu_ms": ms}
        if not args.no_cpu:
            t0 = time.perf_counter()
            want = R.background_forward(img, m, method)
            rec["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
            rec["max_abs_diff"] = int(np.abs(got.astype(int) - want.astype(int)).max())
            rec["mismatch_fraction"] = float((got != want).mean())
        print(json.dumps(rec), flush=True)
    # bg.py:74-79: dilated binary matte, the three planes at full resolution
    alpha = R.dilate_mask(m, 3, 2)
    planes = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).cuda()
    md = torch.from_numpy(alpha).cuda()
    src = planes.to(torch.float64)
    its = ops.laplace_fill(src, md)[1]
    ms, got = wall(lambda: ops.regionfill(planes, md, 1.0), args.reps)
    rec = {"what": "regionfill of the B, G, R planes at 1080p (bg.py:79), device tensors", "gpu_ms": ms, "cg_iterations": its,
           "hole_pixels": int((alpha > 0).sum())}
    if not args.no_cpu:
        t0 = time.perf_counter()
        want = R.regionfill(img[:, :, 0], alpha, 1.0)
        rec["cpu_oracle_ms_per_plane"] = (time.perf_counter() - t0) * 1e3
        g0 = got[0].cpu().numpy()
        rec["max_abs_diff_float"] = float(np.abs(g0 - want).max())
        rec["uint8_flips"] = int((g0.astype(np.uint8) != want.astype(np.uint8)).sum())
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()

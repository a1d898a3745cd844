#!/usr/bin/env python
"""chunk size x stream count sweep of the chunked clip pipelines (device-resident, CUDA events): python tools/sweep_chunks.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as BC  # noqa: E402
from video_unscreen_b200 import clip  # noqa: E402
from video_unscreen_b200.unscreen.trimap import TrimapAgent  # noqa: E402

D = BC.Dist()
torch.cuda.set_device(0)


def t(fn, steps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


n, h, w = 300, 1080, 1920
fr, sg, fr_h, sg_h = BC.green_clip_dev(D, n, h, w, distinct=6)
cf, ta = BC.fitted_agent(fr_h[0], sg_h[0]), TrimapAgent()
col = cf.bg_color_bgr()
alpha = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
tri = torch.empty_like(alpha)
for chunk in (25, 50, 75, 100, 150, 300):
    for streams in (1, 2, 3, 4):
        ms = t(lambda: clip.cf_trimap_clip(fr, sg, cf, ta, col, chunk=chunk, out_alpha=alpha, out_trimap=tri, streams=streams))
        print(json.dumps({"workload": "cf_trimap_1080p", "chunk": chunk, "streams": streams, "ms": round(ms, 4)}), flush=True)
del fr, sg, alpha, tri
n, h, w = 48, 2160, 3840
fr, sg, fr_h, sg_h = BC.green_clip_dev(D, n, h, w, distinct=3)
cf = BC.fitted_agent(fr_h[0], sg_h[0])
col = cf.bg_color_bgr()
tile = torch.from_numpy(np.tile(col, (1, 4, 1))).cuda()
res = [None]
for chunk in (8, 12, 16, 24, 48):
    for streams in (1, 2, 3, 4):
        def step():
            res[0] = clip.green_clip(fr, sg, cf, ta, chunk=chunk, bg_color=col, bg_tile=tile, streams=streams, out=res[0])
        print(json.dumps({"workload": "green_4k", "chunk": chunk, "streams": streams, "ms": round(t(step), 4)}), flush=True)
del fr, sg, res
n, h, w = 240, 2160, 3840
fr = BC.make_clip_device(n, h, w, 1, torch.device("cuda"))
mk = BC.make_masks_device(n, h, w, torch.device("cuda"))
res = [None]
for chunk in (12, 24, 48, 80):
    for streams in (1, 2, 3):
        def step():
            res[0] = clip.bgstep_clip(fr, mk, ta, thr=25, chunk=chunk, streams=streams, out=res[0])
        print(json.dumps({"workload": "bgstep_4k_240", "chunk": chunk, "streams": streams, "ms": round(t(step, 3), 4)}), flush=True)

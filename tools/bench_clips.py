#!/usr/bin/env python
"""Throughput of the batched clip pipelines for the other BASELINE.json configs
(config 0 cf+trimap 1080p, config 2 full green pipeline 4K, config 3 person
replacement 1080p, config 4 bg_step median+trimap+composite 4K), device
resident, CUDA events, with the oracle port timed on a few frames beside it.
One JSON line per workload.  bench.py (the driver contract) stays on config 1."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import refport as R  # noqa: E402
from video_unscreen_b200 import _lib, clip, ops, synth  # noqa: E402
from video_unscreen_b200.unscreen.colorfiltering import ColorFilteringAgent  # noqa: E402
from video_unscreen_b200.unscreen.trimap import TrimapAgent  # noqa: E402


def timed(fn, steps, warmup=2):
    L = _lib.lib()
    if GRAPH:
        l0 = L.vu_launch_count()
        g = clip.Graphed(fn)
        per_step = int(L.vu_launch_count() - l0) // 2      # one eager run + one captured run
        fn = g.replay
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    l0 = L.vu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps, (per_step if GRAPH else int(L.vu_launch_count() - l0) // steps)


GRAPH = False


def green_clip_dev(n, h, w, distinct=6):
    fr, sg = zip(*[synth.green_frame(h, w, t=t, n=distinct, seed=0) for t in range(distinct)])
    fr = torch.from_numpy(np.stack(fr)).cuda()
    sg = torch.from_numpy(np.stack(sg)).cuda()
    idx = torch.arange(n, device="cuda") % distinct
    return fr[idx].contiguous(), sg[idx].contiguous(), np.stack(fr.cpu().numpy()), np.stack(sg.cpu().numpy())


def fitted_agent(frame, seg):
    ag = ColorFilteringAgent()
    np.random.seed(0)
    ag.forward(frame, seg, 3)
    return ag


def report(name, desc, frames, ms, launches, algo_bytes, cpu_fps, cpu_note, peak):
    ach = algo_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": name, "description": desc, "value": frames / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                      "frames_per_step": frames, "gpu_launches_per_step": launches,
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                   "kernel": "whole pipeline (all launches of the step)", "algorithmic_bytes_per_step": algo_bytes},
                      "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": bench.host_threads(), "kind": "port", "sample": cpu_note}}),
          flush=True)


def main():
    chunk4k = int(os.environ.get("VU_CHUNK4K", "24"))   # frames per chunk of the 4K pipelines
    streams = int(os.environ.get("VU_STREAMS", "2"))     # chunks overlap on this many streams
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--graph", action="store_true", help="capture each pipeline as one CUDA graph and time its replays")
    args = ap.parse_args()
    global GRAPH
    GRAPH = args.graph
    peak, _ = bench.measured_peak()
    ta = TrimapAgent()
    want = lambda k: not args.only or k in args.only.split(",")

    if want("cf_trimap_1080p"):
        n, h, w = 300, 1080, 1920
        fr, sg, fr_h, sg_h = green_clip_dev(n, h, w)
        cf = fitted_agent(fr_h[0], sg_h[0])
        col = cf.bg_color_bgr()
        alpha = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        tri = torch.empty_like(alpha)

        chunk = int(os.environ.get("VU_CHUNK", "50"))

        def step():
            clip.cf_trimap_clip(fr, sg, cf, ta, col, chunk=chunk, out_alpha=alpha, out_trimap=tri, streams=streams)
        ms, launches = timed(step, args.steps)
        lb, lf, bgh = cf.tables()
        t0 = time.perf_counter()
        a_o, _, _ = R.cf_forward_predict(fr_h[1], sg_h[1], lb, lf, bgh, 960)
        t_o = R.generate_trimap_withbg(a_o, fr_h[1], col, 960)
        dt = time.perf_counter() - t0
        assert np.array_equal(alpha[1].cpu().numpy(), a_o) and np.array_equal(tri[1].cpu().numpy(), t_o)
        report("cf_trimap_1080p", "BASELINE configs[0]: colour filtering predict + trimap with bg colour, 300 x 1080p", n, ms, launches,
               n * 6 * h * w, 1 / dt, "oracle cf_forward_predict + generate_trimap_withbg on 1 frame (numpy, 1 thread); output bit-exact", peak)
        del fr, sg, alpha, tri

    if want("green_4k"):
        n, h, w = 48, 2160, 3840
        fr, sg, fr_h, sg_h = green_clip_dev(n, h, w, distinct=3)
        cf = fitted_agent(fr_h[0], sg_h[0])

        col = cf.bg_color_bgr()
        tile = torch.from_numpy(np.tile(col, (1, 4, 1))).cuda()

        def step():
            return clip.green_clip(fr, sg, cf, ta, chunk=chunk4k, bg_color=col, bg_tile=tile, streams=streams)
        ms, launches = timed(step, max(2, args.steps // 2), warmup=1)
        a_d, t_d, f_d, b_d = [x[1].cpu().numpy() for x in (step() if not GRAPH else clip.green_clip(fr, sg, cf, ta, chunk=chunk4k, bg_color=col, bg_tile=tile, streams=streams))]
        lb, lf, bgh = cf.tables()
        t0 = time.perf_counter()
        a_o, _, _ = R.cf_forward_predict(fr_h[1], sg_h[1], lb, lf, bgh, 960)
        t_o = R.generate_trimap_withbg(a_o, fr_h[1], col, 960)
        b_o = R.patch_bg(np.broadcast_to(col, fr_h[1].shape), fr_h[1], a_o, "lt128")
        f_o = R.get_fg(fr_h[1], a_o, b_o)
        dt = time.perf_counter() - t0
        assert np.array_equal(a_d, a_o) and np.array_equal(t_d, t_o) and np.array_equal(b_d, b_o) and np.array_equal(f_d, f_o)
        report("green_4k", "BASELINE configs[2]: cf predict -> trimap -> patched bg -> get_fg at 4K (CNN stages skipped), 48 frames", n, ms, launches,
               n * 12 * h * w, 1 / dt, "oracle cf_forward_predict + generate_trimap_withbg + patch + get_fg on 1 frame (numpy, 1 thread); alpha, trimap, "
               "bg, fg bit-exact", peak)
        del fr, sg

    if want("replace_1080p"):
        n, h, w = 300, 1080, 1080 * 16 // 9
        g = torch.Generator(device="cuda").manual_seed(3)
        fg = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
        al = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda", generator=g)
        bg = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
        out = [None]

        def step():
            out[0] = clip.replace_clip(fg, al, bg)
        ms, launches = timed(step, args.steps)
        f0, a0, b0 = fg[0].cpu().numpy(), al[0].cpu().numpy(), bg.cpu().numpy()
        t0 = time.perf_counter()
        ref = R.replace_blend(f0, a0, b0)
        dt = time.perf_counter() - t0
        assert np.array_equal(out[0][0].cpu().numpy(), ref)
        report("replace_1080p", "BASELINE configs[3]: person-replacement blend (replace.py:74-76), 300 x 1080p, shared background", n, ms, launches,
               n * 7 * h * w + 3 * h * w, 1 / dt, "oracle replace_blend on 1 frame (numpy float64, 1 thread); output bit-exact", peak)
        del fg, al, bg, out

    if want("replace_geo_1080p"):
        # the whole per-frame body of tools/replace/replace.py:69-76: shift + bicubic rescale of foreground and mask, blend
        n, h, w = 120, 1080, 1080 * 16 // 9
        g = torch.Generator(device="cuda").manual_seed(4)
        fg = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
        al = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda", generator=g)
        bg = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
        out = [None]
        for tag, dx, dy in (("int", 3, -2), ("frac", 0.5, 0.5)):   # the two correspondences hard-coded at replace.py:20-23
            def step():
                out[0] = clip.replace_clip(fg, al, bg, dx, dy, 1.2)
            ms, launches = timed(step, args.steps)
            f0, a0, b0 = fg[0].cpu().numpy(), al[0].cpu().numpy(), bg.cpu().numpy()
            t0 = time.perf_counter()
            ref = R.replace_frame(f0, a0, b0, dx, dy, 1.2)
            dt = time.perf_counter() - t0
            assert np.array_equal(out[0][0].cpu().numpy(), ref)
            report("replace_geo_1080p_" + tag, f"replace.py:69-76 with the geometric pre-steps (shift_fg dx={dx} dy={dy}, rescale_fg 1.2), "
                   "120 x 1080p, single-channel mask, shared background", n, ms, launches, n * 7 * h * w + 3 * h * w, 1 / dt,
                   "oracle replace_frame on 1 frame (numpy, 1 thread); output bit-exact", peak)
        for name, fn in (("shift_int", lambda: ops.shift(fg, 3, -2, 3)), ("shift_frac", lambda: ops.shift(fg, 0.5, 0.5, 3)),
                         ("rescale", lambda: ops.rescale_cubic(fg, 1.2, 3)), ("rescale_mask", lambda: ops.rescale_cubic(al, 1.2, 1))):
            ms, launches = timed(fn, args.steps)
            per = (1 if name == "rescale_mask" else 3) * h * w
            print(json.dumps({"kernel": name, "ms": round(ms, 4), "frames": n, "gbps": round(2 * n * per / ms / 1e6, 1),
                              "frac": round(2 * n * per / ms / 1e6 / peak, 3)}), flush=True)
        del fg, al, bg, out

    if want("color_correct_1080p"):
        n, h, w = 120, 1080, 1080 * 16 // 9
        fr, sg, _, _ = green_clip_dev(n, h, w)
        g = torch.Generator(device="cuda").manual_seed(5)
        al = torch.minimum(sg, torch.randint(0, 256, sg.shape, dtype=torch.uint8, device="cuda", generator=g) | 128)
        col = np.array([60, 200, 40], np.uint8)
        out = [None]

        def step():
            out[0] = clip.color_correct_clip(fr, al, col, chunk=int(os.environ.get("VU_CHUNK", "60")), streams=streams)
        ms, launches = timed(step, args.steps)
        f0, a0 = fr[0].cpu().numpy(), al[0].cpu().numpy()
        t0 = time.perf_counter()
        ref = R.color_correct(f0, a0, col)
        dt = time.perf_counter() - t0
        assert np.array_equal(out[0][0].cpu().numpy(), ref)
        report("color_correct_1080p", "color_correct (imgprocess.py:263-300, green.py:120), 120 x 1080p, working resolution 540 x 960", n, ms,
               launches, n * 5 * h * w, 1 / dt, "oracle color_correct on 1 frame (numpy, 1 thread); output bit-exact", peak)
        del fr, sg, al, out

    if want("masked_mean_1080p"):
        n, h, w = 300, 1080, 1920
        fr = bench.make_clip_device(n, h, w, 2, torch.device("cuda"))
        yy = torch.arange(h, device="cuda", dtype=torch.float32)[:, None]
        xx = torch.arange(w, device="cuda", dtype=torch.float32)[None, :]
        masks = torch.stack([((((xx - w * (0.15 + 0.7 * t / (n - 1))) / (w * 0.12)) ** 2 + ((yy - h / 2.0) / (h * 0.45)) ** 2) <= 1.0).to(torch.uint8) * 255
                             for t in range(n)])
        res = [None]

        def step():
            res[0] = ops.masked_temporal_mean_raw(fr, masks, 3, 2, 10)   # bg_offline.py:116-125, the dilation fused in
        ms, launches = timed(step, args.steps)
        rows = 64
        t0 = time.perf_counter()
        bg_o, always_o = R.masked_temporal_mean(fr[:, :rows].cpu().numpy(), masks[:, :rows].cpu().numpy())
        dt = (time.perf_counter() - t0) * (h / rows)
        # rows near the cut see a different dilation: compare away from it
        assert np.array_equal(res[0][0][:rows - 4].cpu().numpy(), bg_o[:rows - 4]) and np.array_equal(res[0][1][:rows - 4].cpu().numpy(), always_o[:rows - 4])
        report("masked_mean_1080p", "overall background of bg_offline.py:106-125 (dilate(3,2) of the masks + masked temporal mean), 300 x 1080p", n, ms,
               launches, n * 4 * h * w + 4 * h * w, n / dt, "oracle masked_temporal_mean on a 64-row strip (numpy, 1 thread), scaled; output bit-exact", peak)
        del fr, masks

    if want("bgstep_4k"):
        n, h, w = 120, 2160, 3840
        fr = bench.make_clip_device(n, h, w, 1, torch.device("cuda"))
        yy = torch.arange(h, device="cuda", dtype=torch.float32)[:, None]
        xx = torch.arange(w, device="cuda", dtype=torch.float32)[None, :]
        masks = torch.stack([((((xx - w * (0.15 + 0.7 * t / (n - 1))) / (w * 0.12)) ** 2 + ((yy - h / 2.0) / (h * 0.45)) ** 2) <= 1.0).to(torch.uint8) * 255
                             for t in range(n)])

        def step():
            return clip.bgstep_clip(fr, masks, ta, chunk=chunk4k, streams=streams)
        ms, launches = timed(step, max(2, args.steps // 2), warmup=1)
        bg_d, a_d, t_d, f_d = clip.bgstep_clip(fr, masks, ta, chunk=chunk4k, streams=streams)
        i = n // 2
        frame_h, mask_h, bg_h = fr[i].cpu().numpy(), masks[i].cpu().numpy(), bg_d.cpu().numpy()
        rows = 64                                             # the oracle median on a strip, scaled (np.partition over 120 frames)
        strip = fr[:, :rows].cpu().numpy()
        t0 = time.perf_counter()
        med_o = R.temporal_median(strip)
        dt_med = (time.perf_counter() - t0) * (h / rows)
        assert np.array_equal(med_o, bg_h[:rows])
        t0 = time.perf_counter()
        a_o = R.bgdiff_gate(frame_h, bg_h, mask_h, 25)
        t_o = R.generate_trimap(a_o, 960)
        f_o = R.get_fg(frame_h, a_o, R.patch_bg(bg_h, frame_h, a_o, "eq0"))
        dt = time.perf_counter() - t0
        assert np.array_equal(a_d[i].cpu().numpy(), a_o) and np.array_equal(t_d[i].cpu().numpy(), t_o) and np.array_equal(f_d[i].cpu().numpy(), f_o)
        report("bgstep_4k", "BASELINE configs[4] on one GPU tile: temporal median + difference gate + trimap + get_fg at 4K, 120 frames", n, ms, launches,
               n * 12 * h * w + 6 * h * w, n / (dt_med + n * dt),
               "oracle temporal_median on a 64-row strip (scaled to the frame) + bgdiff_gate + generate_trimap + patch + get_fg on 1 frame "
               "(numpy, 1 thread); background, alpha, trimap, fg bit-exact", peak)


if __name__ == "__main__":
    main()

"""The five BASELINE.json configs as bench workloads (used by bench.py; also runnable alone:
``python tools/bench_configs.py [--only cf_trimap_1080p,...]`` prints one JSON line per workload).

Every workload: seeded synthetic clip resident in HBM, the production clip pipeline of video_unscreen_b200.clip with the
production chunk sizes on two streams, CUDA events around ``steps`` repetitions, one frame (or a window of it) compared
bit-exactly with the oracle OUTSIDE the timed region, the oracle timed on the host cores beside it, and an end-to-end
figure with the clip in pinned host memory (H2D of the inputs and D2H of every output inside the timed region).

Multi-GPU (torchrun): per-frame configs are frame-sharded with shard.frame_ranges over a clip that grows with the world
size (weak scaling: a fixed number of frames per GPU); the bg_step config is ROW-TILE sharded over ONE clip (strong
scaling) with shard.row_tiles + shard.bgstep_halo.  No data-path collective: torch.distributed only carries the barrier
and the max / min over ranks of the timings.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P1080 = 1080 * 1920
P4K = 2160 * 3840

# name -> (BASELINE.json configs index, frames per GPU (None: see workload), H, W, description)
CONFIGS = {
    "cf_trimap_1080p": (0, 300, 1080, 1920, "green-screen colour filtering (predict) + trimap with bg colour, 300 x 1080p"),
    "median_1080p": (1, 300, 1080, 1920, "bg_step temporal-median background over 300 x 1080p"),
    "green_4k": (2, 48, 2160, 3840, "full green pipeline (cf -> trimap -> patched bg -> get_fg) at 4K, 48 frames"),
    "replace_1080p": (3, 300, 1080, 1920, "person-replacement blend (replace.py:74-76), 300 x 1080p, frame-sharded"),
    "bgstep_4k": (4, 500, 2160, 3840, "bg_step median + difference gate + trimap + get_fg at 4K, row-tile sharded"),
}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------------
# distributed helpers (no-ops for world == 1)
# ------------------------------------------------------------------------------------------------------------------

class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def minmax(self, v):
        """(min, max) of a python float over the ranks"""
        if self.world == 1:
            return v, v
        import torch.distributed as dist
        t = self.torch.tensor([v, -v], device=self.dev, dtype=self.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return -float(t[1].item()), float(t[0].item())

    def all_true(self, ok):
        if self.world == 1:
            return bool(ok)
        import torch.distributed as dist
        t = self.torch.tensor([0 if ok else 1], device=self.dev, dtype=self.torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item()) == 0


def timed(D, fn, steps, warmup):
    """-> (max over ranks of the total ms for ``steps`` runs, min over ranks, launches per step on this rank)"""
    torch = D.torch
    from video_unscreen_b200 import _lib
    L = _lib.lib()
    for _ in range(warmup):
        fn()
    D.barrier()
    l0 = L.vu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    D.barrier()
    ms = e0.elapsed_time(e1)
    lo, hi = D.minmax(ms)
    return hi, lo, int(L.vu_launch_count() - l0) // steps


def timed_e2e(D, fn, steps):
    """wall clock around ``steps`` runs of fn (which ends with its results on the host) -> max over ranks, seconds per step"""
    if steps <= 0:
        return float("nan")
    fn()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    D.torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    _, hi = D.minmax(dt)
    return hi / steps


def parallel_cpu(fn, items, threads):
    """fn over items on a thread pool (numpy / the oracle release the GIL in their heavy loops) -> (seconds, results)"""
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        out = list(ex.map(fn, items))
    return time.perf_counter() - t0, out


def pinned_like(torch, t):
    return torch.empty(t.shape, dtype=t.dtype, pin_memory=True)


def block(name, D, n_total, ms_total, steps, launches, algo_bytes_rank_step, peak, scaling, sharding, cpu, e2e, exact, extra=None):
    """one entry of the bench line's ``configs`` object.  The roofline fraction is per GPU: this rank's algorithmic bytes
    per step / (max-over-ranks step time) against one GPU's measured HBM peak."""
    idx, _, h, w, desc = CONFIGS[name]
    ms = ms_total / steps
    ach = algo_bytes_rank_step / (ms * 1e-3) / 1e9
    b = {"baseline_config": idx, "description": desc, "value": n_total / (ms * 1e-3), "unit": "frames/s", "frames_per_step": n_total,
         "height": h, "width": w, "n_gpus": D.world, "ms_per_step": ms, "steps": steps, "gpu_launches_per_step": launches,
         "scaling": scaling, "sharding": sharding,
         "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "algorithmic_bytes_per_gpu_step": int(algo_bytes_rank_step), "traffic": traffic(name),
                      "kernel": "whole pipeline (all launches of the step)"},
         "cpu_baseline": cpu, "e2e": e2e, "bit_exact": bool(exact)}
    if extra:
        b.update(extra)
    return b


def traffic(name):
    """dram bytes per step (sum over the step's kernels) from the committed ncu capture, or None"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        v = json.load(open(p)).get(name)
        return int(v) if v is not None else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------------
# synthetic clips on the device
# ------------------------------------------------------------------------------------------------------------------

def green_clip_dev(D, n, h, w, distinct, seed=0):
    """n frames cycling through ``distinct`` host-generated green-screen frames (synth.green_frame)"""
    torch = D.torch
    from video_unscreen_b200 import synth
    fr, sg = zip(*[synth.green_frame(h, w, t=t, n=distinct, seed=seed) for t in range(distinct)])
    fr_h, sg_h = np.stack(fr), np.stack(sg)
    idx = torch.arange(n, device=D.dev) % distinct
    return torch.from_numpy(fr_h).to(D.dev)[idx].contiguous(), torch.from_numpy(sg_h).to(D.dev)[idx].contiguous(), fr_h, sg_h


def make_clip_device(n, h, w, seed, device, rows=None, keep=None):
    """synthetic bg_step clip generated on the device (textured static background, per-frame noise in [-6,6], a moving
    ellipse covering each pixel in < 50 % of the frames) -- same structure as video_unscreen_b200.synth.bgstep_clip.
    ``rows`` = (a0, a1): every frame is generated whole (so the bytes do not depend on who holds which rows) and only
    rows [a0, a1) are kept.  ``keep`` = (r0, r1, c0, c1): additionally returns that window of every frame."""
    import torch
    g = torch.Generator(device=device).manual_seed(1000 + seed)
    tex = torch.randint(0, 256, (1, 3, h, w), device=device, generator=g, dtype=torch.uint8).float()
    bg = torch.nn.functional.avg_pool2d(tex, 11, stride=1, padding=5, count_include_pad=False)[0].permute(1, 2, 0)
    bg = bg.clamp(0, 255).to(torch.int16)
    a0, a1 = rows if rows is not None else (0, h)
    frames = torch.empty((n, a1 - a0, w, 3), dtype=torch.uint8, device=device)
    win = torch.empty((n, keep[1] - keep[0], keep[3] - keep[2], 3), dtype=torch.uint8, device=device) if keep else None
    yy = torch.arange(h, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(w, device=device, dtype=torch.float32)[None, :]
    person = torch.tensor([120, 140, 200], device=device, dtype=torch.int16)
    for t in range(n):
        noise = torch.randint(-6, 7, (h, w, 3), device=device, generator=g, dtype=torch.int16)
        f = (bg + noise).clamp_(0, 255)
        cx = w * (0.15 + 0.7 * t / max(n - 1, 1))
        ell = (((xx - cx) / (w * 0.1)) ** 2 + ((yy - h / 2.0) / (h * 0.4)) ** 2) <= 1.0
        pn = torch.randint(-40, 41, (h, w, 3), device=device, generator=g, dtype=torch.int16)
        f = torch.where(ell[..., None], (person + pn).clamp_(0, 255), f).to(torch.uint8)
        frames[t] = f[a0:a1]
        if keep:
            win[t] = f[keep[0]:keep[1], keep[2]:keep[3]]
    return (frames, win) if keep else frames


def make_masks_device(n, h, w, device, rows=None, keep=None):
    """coarse person masks for make_clip_device's clip: a slightly larger ellipse on the same track (0 / 255)"""
    import torch
    a0, a1 = rows if rows is not None else (0, h)
    yy = torch.arange(h, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(w, device=device, dtype=torch.float32)[None, :]
    masks = torch.empty((n, a1 - a0, w), dtype=torch.uint8, device=device)
    win = torch.empty((n, keep[1] - keep[0], keep[3] - keep[2]), dtype=torch.uint8, device=device) if keep else None
    for t in range(n):
        m = ((((xx - w * (0.15 + 0.7 * t / max(n - 1, 1))) / (w * 0.12)) ** 2 + ((yy - h / 2.0) / (h * 0.45)) ** 2) <= 1.0).to(torch.uint8) * 255
        masks[t] = m[a0:a1]
        if keep:
            win[t] = m[keep[0]:keep[1], keep[2]:keep[3]]
    return (masks, win) if keep else masks


def fitted_agent(frame, seg):
    """a ColorFilteringAgent whose mixtures were fitted once on frame 0 (np.random.seed(0), iters=3: SURVEY 8d config 1)"""
    from video_unscreen_b200.unscreen.colorfiltering import ColorFilteringAgent
    ag = ColorFilteringAgent()
    np.random.seed(0)
    ag.forward(frame, seg, 3)
    return ag


# ------------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------------

def run_cf_trimap_1080p(D, steps, warmup, peak, cpu=True, e2e_steps=2):
    """BASELINE configs[0] (green.py:99-114): every rank owns a 300-frame range of a (300 * world)-frame clip"""
    from oracle import refport as R
    from video_unscreen_b200 import clip, shard
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    torch = D.torch
    name = "cf_trimap_1080p"
    _, per, h, w, _ = CONFIGS[name]
    s, e = shard.my_frame_range(per * D.world, D.rank, D.world, align=30)
    n = e - s
    fr, sg, fr_h, sg_h = green_clip_dev(D, n, h, w, distinct=6)
    cf, ta = fitted_agent(fr_h[0], sg_h[0]), TrimapAgent()     # same seed on every rank: identical tables (else: broadcast them)
    col = cf.bg_color_bgr()
    alpha = torch.empty((n, h, w), dtype=torch.uint8, device=D.dev)
    tri = torch.empty_like(alpha)

    def step():
        clip.cf_trimap_clip(fr, sg, cf, ta, col, chunk=50, out_alpha=alpha, out_trimap=tri, streams=2)
    ms, ms_lo, launches = timed(D, step, steps, warmup)
    lb, lf, bgh = cf.tables()

    def oracle(i):
        a_o, _, _ = R.cf_forward_predict(fr_h[i % 6], sg_h[i % 6], lb, lf, bgh, 960)
        return a_o, R.generate_trimap_withbg(a_o, fr_h[i % 6], col, 960)
    a_o, t_o = oracle(1)
    exact = D.all_true(np.array_equal(alpha[1].cpu().numpy(), a_o) and np.array_equal(tri[1].cpu().numpy(), t_o)
                       and np.array_equal(alpha[n - 5].cpu().numpy(), oracle(n - 5)[0]))
    cpu_b = None
    if cpu and D.rank == 0:
        th = host_threads()
        dt, _ = parallel_cpu(oracle, list(range(th)), th)
        cpu_b = {"value": th / dt, "unit": "frames/s", "cores": th, "kind": "port",
                 "sample": f"{th} frames, one per host thread, in {dt:.1f} s: oracle.refport.cf_forward_predict + generate_trimap_withbg "
                           "(numpy restatement of colorfiltering/agent.py:285-354 + trimap/agent.py:63-101; NOT the reference's cv2/torch "
                           "build, which the survey timed at ~5.2 frames/s on 8 vCPUs)"}
    # end to end: clip in pinned host memory -> alpha, trimap in pinned host memory
    fr_p, sg_p = pinned_like(torch, fr).copy_(fr), pinned_like(torch, sg).copy_(sg)
    a_p, t_p = pinned_like(torch, alpha), pinned_like(torch, tri)

    def e2e_step():
        clip.streamed([fr_p, sg_p], [a_p, t_p], lambda s0, e0, f, m, outs: clip.cf_trimap_clip(
            f, m, cf, ta, col, chunk=50, out_alpha=outs[0], out_trimap=outs[1], streams=1), chunk=50)
    sec = timed_e2e(D, e2e_step, e2e_steps)
    exact = exact and (e2e_steps <= 0 or D.all_true(torch.equal(a_p, alpha.cpu()) and torch.equal(t_p, tri.cpu())))
    e2e = {"value": n * D.world / sec, "unit": "frames/s", "h2d_bytes_per_step": int(fr.numel() + sg.numel()),
           "d2h_bytes_per_step": int(alpha.numel() + tri.numel()), "ms_per_step": sec * 1e3, "steps": e2e_steps,
           "api": "clip.streamed(pinned host frames + masks -> clip.cf_trimap_clip -> pinned host alpha + trimap), chunks of 50 frames, "
                  "H2D / kernels / D2H on three streams"}
    return block(name, D, n * D.world, ms, steps, launches, n * 6 * h * w, peak, "weak", "frame ranges (shard.frame_ranges, align 30)",
                 cpu_b, e2e, exact, {"rank_ms_min_max": [ms_lo / steps, ms / steps], "chunk": 50, "streams": 2})


def run_green_4k(D, steps, warmup, peak, cpu=True, e2e_steps=2):
    """BASELINE configs[2] (green.py:99-126 without the CNN stages), 48 frames per GPU"""
    from oracle import refport as R
    from video_unscreen_b200 import clip, shard
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    torch = D.torch
    name = "green_4k"
    _, per, h, w, _ = CONFIGS[name]
    s, e = shard.my_frame_range(per * D.world, D.rank, D.world, align=1)
    n = e - s
    fr, sg, fr_h, sg_h = green_clip_dev(D, n, h, w, distinct=3)
    cf, ta = fitted_agent(fr_h[0], sg_h[0]), TrimapAgent()
    col = cf.bg_color_bgr()
    tile = torch.from_numpy(np.tile(col, (1, 4, 1))).to(D.dev)
    res = [None]

    def step():      # the results of the first call are reused: no allocation in the steady state
        res[0] = clip.green_clip(fr, sg, cf, ta, chunk=24, bg_color=col, bg_tile=tile, streams=2, out=res[0])
    ms, ms_lo, launches = timed(D, step, max(2, steps // 2), max(2, warmup // 2))
    steps_used = max(2, steps // 2)
    lb, lf, bgh = cf.tables()

    def oracle(i):
        f, m = fr_h[i % 3], sg_h[i % 3]
        a_o, _, _ = R.cf_forward_predict(f, m, lb, lf, bgh, 960)
        t_o = R.generate_trimap_withbg(a_o, f, col, 960)
        b_o = R.patch_bg(np.broadcast_to(col, f.shape), f, a_o, "lt128")
        return a_o, t_o, R.get_fg(f, a_o, b_o), b_o
    want = oracle(1)
    got = [x[1].cpu().numpy() for x in res[0]]
    exact = D.all_true(all(np.array_equal(g, w_) for g, w_ in zip(got, want)))
    cpu_b = None
    if cpu and D.rank == 0:
        th = min(host_threads(), 8)
        dt, _ = parallel_cpu(oracle, list(range(th)), th)
        cpu_b = {"value": th / dt, "unit": "frames/s", "cores": th, "kind": "port",
                 "sample": f"{th} 4K frames, one per host thread, in {dt:.1f} s: oracle cf_forward_predict + generate_trimap_withbg + patch + get_fg "
                           "(numpy restatement; the survey timed the reference's cv2/torch build at ~1.2 frames/s on 8 vCPUs)"}
    fr_p, sg_p = pinned_like(torch, fr).copy_(fr), pinned_like(torch, sg).copy_(sg)
    outs_p = [pinned_like(torch, x) for x in res[0]]

    def e2e_step():
        def body(s0, e0, f, m, outs):
            a, t, fg, bg = clip.green_clip(f, m, cf, ta, chunk=24, bg_color=col, bg_tile=tile, streams=1)
            for o, x in zip(outs, (a, t, fg, bg)):
                o.copy_(x)
        clip.streamed([fr_p, sg_p], outs_p, body, chunk=12)
    sec = timed_e2e(D, e2e_step, e2e_steps)
    exact = exact and (e2e_steps <= 0 or D.all_true(all(torch.equal(o[1], x[1].cpu()) for o, x in zip(outs_p, res[0]))))
    e2e = {"value": n * D.world / sec, "unit": "frames/s", "h2d_bytes_per_step": int(fr.numel() + sg.numel()),
           "d2h_bytes_per_step": int(sum(x.numel() for x in res[0])), "ms_per_step": sec * 1e3, "steps": e2e_steps,
           "api": "clip.streamed(pinned host frames + masks -> clip.green_clip -> pinned host alpha, trimap, fg, bg), chunks of 12 frames"}
    return block(name, D, n * D.world, ms, steps_used, launches, n * 12 * h * w, peak, "weak", "frame ranges (shard.frame_ranges)",
                 cpu_b, e2e, exact, {"rank_ms_min_max": [ms_lo / steps_used, ms / steps_used], "chunk": 24, "streams": 2})


def run_replace_1080p(D, steps, warmup, peak, cpu=True, e2e_steps=2):
    """BASELINE configs[3] (replace.py:74-76): 300 frames per GPU, frame-sharded, shared new background"""
    from oracle import refport as R
    from video_unscreen_b200 import clip, shard
    torch = D.torch
    name = "replace_1080p"
    _, per, h, w, _ = CONFIGS[name]
    s, e = shard.my_frame_range(per * D.world, D.rank, D.world, align=1)
    n = e - s
    g = torch.Generator(device=D.dev).manual_seed(3 + D.rank)
    fg = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=D.dev, generator=g)
    al = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=D.dev, generator=g)
    gb = torch.Generator(device=D.dev).manual_seed(33)
    bg = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=D.dev, generator=gb)
    out = torch.empty_like(fg)

    def step():
        clip.replace_clip(fg, al, bg, out=out)
    ms, ms_lo, launches = timed(D, step, steps, warmup)
    # the same clip under a realistic matte (alpha from the green clip): 0 outside a moving ellipse, 255 inside,
    # a soft ring between: groups of sixteen pixels with alpha all 0 / all 255 are copies of one image (exact), the other
    # image is not read.  Reported beside the uniform-random alpha above, which is the kernel's worst case.
    yy = torch.arange(h, device=D.dev, dtype=torch.float32)[:, None]
    xx = torch.arange(w, device=D.dev, dtype=torch.float32)[None, :]
    al2 = torch.empty_like(al)
    for t in range(n):
        cx = w * (0.3 + 0.4 * t / max(n - 1, 1))
        r = torch.sqrt(((xx - cx) / (w * 0.156)) ** 2 + ((yy - h / 2.0) / (h * 0.417)) ** 2)
        al2[t] = ((1.03 - r) / 0.06 * 255).clamp_(0, 255).to(torch.uint8)
    out2 = torch.empty_like(fg)

    def step2():
        clip.replace_clip(fg, al2, bg, out=out2)
    ms2, _, _ = timed(D, step2, steps, warmup)
    a2_h = al2[n // 2].cpu().numpy()
    matte_exact = D.all_true(np.array_equal(out2[n // 2].cpu().numpy(), R.replace_blend(fg[n // 2].cpu().numpy(), a2_h, bg.cpu().numpy())))
    soft = float(((a2_h > 0) & (a2_h < 255)).mean())
    del al2, out2
    f_h, a_h, b_h = fg[:16].cpu().numpy(), al[:16].cpu().numpy(), bg.cpu().numpy()
    oracle = lambda i: R.replace_blend(f_h[i % 16], a_h[i % 16], b_h)
    exact = D.all_true(np.array_equal(out[1].cpu().numpy(), oracle(1)) and np.array_equal(out[15].cpu().numpy(), oracle(15)))
    cpu_b = None
    if cpu and D.rank == 0:
        th = host_threads()
        dt, _ = parallel_cpu(oracle, list(range(2 * th)), th)
        cpu_b = {"value": 2 * th / dt, "unit": "frames/s", "cores": th, "kind": "port",
                 "sample": f"{2 * th} frames on {th} host threads in {dt:.1f} s: oracle.refport.replace_blend (the float64 numpy expression of "
                           "replace.py:74-76 itself)"}
    fg_p, al_p, out_p = pinned_like(torch, fg).copy_(fg), pinned_like(torch, al).copy_(al), pinned_like(torch, out)

    def e2e_step():
        clip.streamed([fg_p, al_p], [out_p], lambda s0, e0, f, a, outs: clip.replace_clip(f, a, bg, out=outs[0]), chunk=50)
    sec = timed_e2e(D, e2e_step, e2e_steps)
    exact = exact and (e2e_steps <= 0 or D.all_true(torch.equal(out_p[7], out[7].cpu()) and torch.equal(out_p[n - 1], out[n - 1].cpu())))
    e2e = {"value": n * D.world / sec, "unit": "frames/s", "h2d_bytes_per_step": int(fg.numel() + al.numel()),
           "d2h_bytes_per_step": int(out.numel()), "ms_per_step": sec * 1e3, "steps": e2e_steps,
           "api": "clip.streamed(pinned host fg + mask -> clip.replace_clip -> pinned host composite), chunks of 50 frames; the new background "
                  "stays on the device"}
    return block(name, D, n * D.world, ms, steps, launches, n * 7 * h * w + 3 * h * w, peak, "weak", "frame ranges (shard.frame_ranges)",
                 cpu_b, e2e, exact and matte_exact,
                 {"rank_ms_min_max": [ms_lo / steps, ms / steps], "alpha": "uniform random bytes (worst case: every pixel takes the float64 path)",
                  "realistic_matte": {"ms_per_step": ms2 / steps, "value": n * D.world / (ms2 / steps * 1e-3), "unit": "frames/s",
                                      "frac_of_10P_roofline": (n * 7 * h * w + 3 * h * w) / (ms2 / steps * 1e-3) / 1e9 / peak,
                                      "soft_pixels": soft, "bit_exact": bool(matte_exact),
                                      "what": "alpha = 0 outside a moving ellipse, 255 inside, a soft ring between (6 % of the radius)"}})


def run_bgstep_4k(D, steps, warmup, peak, cpu=True, e2e_steps=1, frames=None):
    """BASELINE configs[4] (bg_offline.py:150-172 after a temporal-median background): ONE seeded 4K clip, every rank owns
    a row tile (+ halo) of ALL frames: strong scaling.  500 frames by default so that the clip and its outputs fit one
    GPU at N = 1 (the named 2000-frame clip is 150 GB with its outputs); ``frames`` = 2000 at N = 8 is the named config."""
    from oracle import refport as R
    from video_unscreen_b200 import clip
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    torch = D.torch
    name = "bgstep_4k"
    _, n_default, h, w, _ = CONFIGS[name]
    n = int(frames or n_default)
    ta = TrimapAgent()
    r0, r1, ht, hb, scale, _ = clip.bgstep_tile_geometry(h, w, ta, D.rank, D.world)
    a0, a1 = r0 - ht, r1 + hb
    # rank 0 keeps a window of every frame around its lower seam (the middle of the frame when it owns all rows) for the oracle
    seam = r1 if r1 < h else h // 2
    keep = (seam - 64, seam + 64, w // 2 - 256, w // 2 + 256) if D.rank == 0 else None
    fr = make_clip_device(n, h, w, 1, D.dev, rows=(a0, a1), keep=keep)
    mk = make_masks_device(n, h, w, D.dev, rows=(a0, a1), keep=keep)
    if keep:
        (fr, fr_win), (mk, mk_win) = fr, mk
    res, bufs = [None], [None]
    chunk = 24 * D.world     # a row tile is 1 / world of a frame: chunks of the same number of pixels as 24 whole frames

    def step():      # the result buffers (tile plus halo) of the first call are reused: no allocation in the steady state
        res[0] = clip.bgstep_clip_tile(fr, mk, ta, D.rank, D.world, thr=25, chunk=chunk, rows=(a0, a1, h), out=bufs[0])
        if bufs[0] is None:
            bufs[0] = tuple(x._base if x._base is not None else x for x in res[0][1:])
    st = max(4, steps // 2)
    ms, ms_lo, launches = timed(D, step, st, 2)
    (_, _), bg_t, a_t, t_t, f_t = res[0]
    exact, cpu_b = True, None
    if D.rank == 0:
        # oracle on the 128 x 512 window (working resolution 32 x 128, aligned to the frame's grid); rows / columns further
        # than the stencils' reach (28 rows, 28 columns) from the window's edges equal the whole frame's
        fw, mw = fr_win.cpu().numpy(), mk_win.cpu().numpy()
        t0 = time.perf_counter()
        bg_o = R.temporal_median(fw)
        t_med = time.perf_counter() - t0
        ys, xs = slice(keep[0] + 32 - r0, seam - r0 if r1 < h else keep[1] - 32 - r0), slice(keep[2] + 32, keep[3] - 32)
        wy = slice(32, (seam - keep[0]) if r1 < h else 96)
        exact = np.array_equal(bg_t[ys, xs].cpu().numpy(), bg_o[wy, 32:-32])
        t_pf = 0.0
        for i in (0, n // 2, n - 1):
            t0 = time.perf_counter()
            a_o = R.bgdiff_gate(fw[i], bg_o, mw[i], 25)
            t_o = R.generate_trimap(a_o, 128)
            f_o = R.get_fg(fw[i], a_o, R.patch_bg(bg_o, fw[i], a_o, "eq0"))
            t_pf += time.perf_counter() - t0
            exact = exact and np.array_equal(a_t[i][ys, xs].cpu().numpy(), a_o[wy, 32:-32]) and \
                np.array_equal(t_t[i][ys, xs].cpu().numpy(), t_o[wy, 32:-32]) and np.array_equal(f_t[i][ys, xs].cpu().numpy(), f_o[wy, 32:-32])
        if cpu:
            # scale the window's oracle time to the whole clip: median per pixel, per-frame stages per pixel and frame
            frac = (128 * 512) / (h * w)
            t_clip = t_med / frac + (t_pf / 3) * n / frac
            cpu_b = {"value": n / t_clip, "unit": "frames/s", "cores": 1, "kind": "port",
                     "sample": f"128 x 512 window of all {n} frames: oracle temporal_median {t_med:.1f} s + (bgdiff_gate + generate_trimap + patch + "
                               f"get_fg) on 3 frames {t_pf:.2f} s, scaled by pixels to the whole clip; one host thread"}
    exact = D.all_true(exact)
    # end to end on the first 96 frames of this rank's rows (pinning the whole tile set of a 500-frame 4K clip is 16 GB per step)
    ne = min(n, 96)
    fr_p, mk_p = pinned_like(torch, fr[:ne]).copy_(fr[:ne]), pinned_like(torch, mk[:ne]).copy_(mk[:ne])
    outs_p = None

    def e2e_step():
        nonlocal outs_p
        f_d, m_d = fr_p.to(D.dev, non_blocking=True), mk_p.to(D.dev, non_blocking=True)
        _, bg_e, a_e, t_e, f_e = clip.bgstep_clip_tile(f_d, m_d, ta, D.rank, D.world, thr=25, chunk=chunk, rows=(a0, a1, h))
        if outs_p is None:
            outs_p = [pinned_like(torch, x) for x in (bg_e, a_e, t_e, f_e)]
        for o, x in zip(outs_p, (bg_e, a_e, t_e, f_e)):
            o.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
    sec = timed_e2e(D, e2e_step, e2e_steps)
    e2e = {"value": ne / sec, "unit": "frames/s", "h2d_bytes_per_step": int(fr_p.numel() + mk_p.numel()),
           "d2h_bytes_per_step": int(sum(o.numel() for o in outs_p or [])), "ms_per_step": sec * 1e3, "steps": e2e_steps, "frames": ne,
           "api": f"first {ne} frames of the rank's rows: pinned host tile -> clip.bgstep_clip_tile -> pinned host background, alpha, trimap, fg "
                  "(the median needs every frame before the per-frame stages start: no overlap of copies and kernels)"}
    rows_alg = r1 - r0
    algo = n * 12 * rows_alg * w + 6 * rows_alg * w
    # data-independent bound: every mask pixel set (the kernel's all-zero-mask tile shortcut never fires; the person masks
    # above cover 17 % of the frame)
    mk.fill_(255)
    ms_d, _, _ = timed(D, step, st, 1)
    dense = {"ms_per_step": ms_d / st, "value": n / (ms_d / st * 1e-3), "frac": algo / (ms_d / st * 1e-3) / 1e9 / peak,
             "what": "the same step with every segmentation-mask pixel set: no tile takes the all-zero-mask shortcut"}
    return block(name, D, n, ms, st, launches, algo, peak, "strong", "row tiles of ONE clip (shard.row_tiles + shard.bgstep_halo: 28 / 24 halo rows "
                 "read from the rank's own rows, no exchange)", cpu_b, e2e, exact,
                 {"frames": n, "rank_ms_min_max": [ms_lo / st, ms / st], "rows_per_gpu": rows_alg, "halo_rows": [ht, hb], "chunk": chunk, "streams": 2, "dense_masks": dense,
                  "note": "the named config is 2000 frames on 8 GPUs (bench.py runs it as bgstep_4k_2000 when --gpus 8); 500 frames keep the "
                          "clip plus outputs within one GPU at N = 1"})


RUNNERS = {"cf_trimap_1080p": run_cf_trimap_1080p, "green_4k": run_green_4k, "replace_1080p": run_replace_1080p, "bgstep_4k": run_bgstep_4k}


def per_frame_latency(D, reps=5):
    """the reference's per-frame numpy API (what tools/unscreen/green.py:99-126 calls), host arrays in and out, beside the
    survey's CPU timings of the reference"""
    from video_unscreen_b200 import synth
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    from video_unscreen_b200.unscreen.utils import get_fg
    torch = D.torch
    out = {}
    for tag, h, w in (("1080p", 1080, 1920), ("4k", 2160, 3840)):
        frame, seg = synth.green_frame(h, w, t=1, n=6, seed=0)
        cf, ta = fitted_agent(frame, seg), TrimapAgent()
        col = cf.bg_color_bgr()
        t = {}
        for _ in range(2):
            a, bgimg, _ = cf.forward(frame, seg, 0)
            tri = ta.forward(a, frame, col)
            get_fg(frame, a, bgimg)
        torch.cuda.synchronize()
        for key, fn in (("cf_forward_predict_ms", lambda: cf.forward(frame, seg, 0)),
                        ("trimap_forward_withbg_ms", lambda: ta.forward(a, frame, col)),
                        ("get_fg_ms", lambda: get_fg(frame, a, bgimg))):
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            t[key] = (time.perf_counter() - t0) / reps * 1e3
        out[tag] = t
    out["reference_cpu_ms_survey_8vcpu"] = {"1080p": {"cf_forward_predict_ms": 162, "trimap_forward_withbg_ms": 30, "get_fg_ms": 37},
                                            "4k": {"cf_forward_predict_ms": 408, "trimap_forward_withbg_ms": 116, "get_fg_ms": 329}}
    out["note"] = "numpy uint8 frames in, numpy out, one call per frame: H2D + kernels + D2H + the host-side synchronisations, wall clock"
    return out


def main():
    import argparse

    import torch

    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--frames", type=int, default=0, help="bgstep_4k: frames of the clip (default 500)")
    args = ap.parse_args()
    D = Dist()
    torch.cuda.set_device(D.local)
    if D.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=D.dev)
    peak, _ = bench.measured_peak()
    for name, fn in RUNNERS.items():
        if args.only and name not in args.only.split(","):
            continue
        kw = {"frames": args.frames} if name == "bgstep_4k" and args.frames else {}
        if args.no_e2e:
            kw["e2e_steps"] = 0
        b = fn(D, args.steps, args.warmup, peak, cpu=not args.no_cpu, **kw)
        if D.rank == 0:
            print(json.dumps({"workload": name, **b}), flush=True)
        torch.cuda.empty_cache()
    if (not args.only or "per_frame" in args.only) and D.rank == 0:
        print(json.dumps({"workload": "per_frame", **per_frame_latency(D)}), flush=True)
    if D.world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

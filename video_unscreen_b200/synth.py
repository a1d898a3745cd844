"""Seeded synthetic clips shared by the tests, the oracle runs and ``bench.py``
(SURVEY.md section 8d).  numpy only, so the same bytes can be produced in the
build container and on the GPU box."""
import numpy as np

GREEN_BG = (60, 200, 40)       # BGR
PERSON = (120, 140, 200)       # BGR


def ellipse_mask(h, w, cx, cy, ax, ay):
    yy, xx = np.mgrid[0:h, 0:w]
    return (((xx - cx) / float(ax)) ** 2 + ((yy - cy) / float(ay)) ** 2) <= 1.0


def _noisy(rng, shape, base, amp):
    n = rng.integers(-amp, amp + 1, size=shape, dtype=np.int16)
    return np.clip(n + np.asarray(base, np.int16), 0, 255).astype(np.uint8)


def _erode_box(m, r):
    """cheap separable box erosion used only to mimic a coarse CNN mask."""
    h, w = m.shape
    out = m.copy()
    p = np.pad(m, r, constant_values=False)
    for d in range(2 * r + 1):
        out &= p[r:r + h, d:d + w]
    p = np.pad(out, r, constant_values=False)
    for d in range(2 * r + 1):
        out &= p[d:d + h, r:r + w]
    return out


def green_frame(h, w, t=0, n=1, seed=0):
    """one green-screen frame + coarse segmentation mask (0/255)."""
    rng = np.random.default_rng([seed, t])
    frame = _noisy(rng, (h, w, 3), GREEN_BG, 12)
    cx = w / 2.0 + 0.1 * w * np.sin(2 * np.pi * t / max(n, 1))
    ell = ellipse_mask(h, w, cx, h / 2.0, w * 0.156, h * 0.417)
    person = _noisy(rng, (h, w, 3), PERSON, 40)
    frame[ell] = person[ell]
    seg = _erode_box(ell, max(1, min(h, w) // 180)).astype(np.uint8) * 255
    return frame, seg


def green_clip(n, h, w, seed=0):
    frames = np.empty((n, h, w, 3), np.uint8)
    segs = np.empty((n, h, w), np.uint8)
    for t in range(n):
        frames[t], segs[t] = green_frame(h, w, t, n, seed)
    return frames, segs


def _lowpass(img, r):
    x = img.astype(np.float32)
    for ax in (0, 1):
        c = np.cumsum(np.concatenate([np.repeat(np.take(x, [0], ax), r + 1, ax), x,
                                      np.repeat(np.take(x, [-1], ax), r, ax)], ax), axis=ax)
        n = x.shape[ax]
        hi = np.take(c, np.arange(2 * r + 1, 2 * r + 1 + n), ax)
        lo = np.take(c, np.arange(0, n), ax)
        x = (hi - lo) / (2 * r + 1)
    return x


def textured_background(h, w, seed=0, r=5):
    rng = np.random.default_rng([seed, 7777])
    tex = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    return np.clip(_lowpass(tex, r) * 1.0, 0, 255).astype(np.uint8)


def bgstep_clip(n, h, w, seed=0, out=None, masks_out=None):
    """static textured background + per-frame noise [-6,6] + a moving ellipse
    that covers any pixel in < 50 % of the frames; masks = the ellipse."""
    bg = textured_background(h, w, seed)
    frames = out if out is not None else np.empty((n, h, w, 3), np.uint8)
    masks = masks_out if masks_out is not None else np.empty((n, h, w), np.uint8)
    for t in range(n):
        rng = np.random.default_rng([seed, 1, t])
        f = rng.integers(-6, 7, size=(h, w, 3), dtype=np.int16) + bg.astype(np.int16)
        f = np.clip(f, 0, 255).astype(np.uint8)
        cx = w * (0.15 + 0.7 * t / max(n - 1, 1))
        ell = ellipse_mask(h, w, cx, h / 2.0, w * 0.1, h * 0.4)
        person = _noisy(rng, (h, w, 3), PERSON, 40)
        f[ell] = person[ell]
        frames[t] = f
        masks[t] = ell.astype(np.uint8) * 255
    return frames, masks, bg


def random_clip(n, h, w, seed=0):
    """uniform random uint8 frames: worst case for off-by-one detection."""
    return np.random.default_rng([seed, 99]).integers(0, 256, (n, h, w, 3), dtype=np.uint8)

"""Probe: model of the IPP cubic path that cv2.resize(INTER_CUBIC) takes in this container."""
import numpy as np, cv2

def coeffs(x, dt):
    A = dt(-0.75); x = dt(x); one = dt(1)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (one - x) - (A + 3)) * (one - x) * (one - x) + 1
    return np.array([c0, c1, c2, one - c0 - c1 - c2], dt)

def axis(dst, src, f, dt):
    scale = 1.0 / f
    idx = np.empty((dst, 4), np.int64); co = np.empty((dst, 4), dt)
    for d in range(dst):
        p = (d + 0.5) * scale - 0.5
        s = int(np.floor(p)); fr = p - s
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, src - 1); co[d] = coeffs(fr, dt)
    return idx, co

def model(img, f, dt, order):
    h, w = img.shape[:2]
    dh, dw = int(round(h * f)), int(round(w * f))
    yi, yc = axis(dh, h, f, dt); xi, xc = axis(dw, w, f, dt)
    im = img.astype(dt)
    if im.ndim == 2: im = im[..., None]
    if order == "hv":
        hor = sum(im[:, xi[:, k]] * xc[:, k][None, :, None] for k in range(4))
        v = sum(hor[yi[:, k]] * yc[:, k][:, None, None] for k in range(4))
    else:
        ver = sum(im[yi[:, k]] * yc[:, k][:, None, None] for k in range(4))
        v = sum(ver[:, xi[:, k]] * xc[:, k][None, :, None] for k in range(4))
    return v.reshape(dh, dw, *img.shape[2:])

rng = np.random.default_rng(0)
img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
ref = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC).astype(int)
for dt in (np.float64, np.float32):
    for order in ("hv", "vh"):
        v = model(img, 1.2, dt, order)
        for name, q in (("rint", np.rint(v)), ("floor+.5", np.floor(v + dt(0.5))), ("trunc", np.trunc(v))):
            o = np.clip(q, 0, 255).astype(int)
            d = np.abs(o - ref)
            print(dt.__name__, order, name, int((d > 0).sum()), int(d.max()))
# where do the f64 mismatches sit?
v = model(img, 1.2, np.float64, "hv"); o = np.clip(np.rint(v), 0, 255).astype(int)
bad = np.argwhere(o != ref)
print("bad rows hist (first 12):", np.bincount(bad[:, 0] % 6, minlength=6), "cols:", np.bincount(bad[:, 1] % 6, minlength=6))
fracs = np.abs(v - np.rint(v))[o != ref]
print("distance to integer at mismatches: min %.4f median %.4f" % (fracs.min(), np.median(fracs)))

"""Probe: which float32 evaluation order reproduces cv2's IPP cubic bit for bit (if any)?"""
import numpy as np, cv2, itertools
f32 = np.float32

def coeffs(x, style):
    A = f32(-0.75); x = f32(x); one = f32(1)
    if style == "cv":
        c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
        c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
        c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
        c3 = one - c0 - c1 - c2
    elif style == "cv4":     # all four from their own polynomial
        c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
        c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
        y = one - x
        c2 = ((A + f32(2)) * y - (A + f32(3))) * y * y + one
        c3 = ((A * (y + one) - f32(5) * A) * (y + one) + f32(8) * A) * (y + one) - f32(4) * A
    else:                    # exact in f64, rounded to f32
        A = -0.75; x = float(x)
        c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c1 = ((A + 2) * x - (A + 3)) * x * x + 1
        c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        c3 = 1 - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], f32)

def axis(dst, src, style, fstyle):
    idx = np.empty((dst, 4), np.int64); co = np.empty((dst, 4), f32)
    for d in range(dst):
        if fstyle == "f64":
            p = (d + 0.5) / 1.2 - 0.5; s = int(np.floor(p)); fr = p - s
        elif fstyle == "f64mul":
            p = (d + 0.5) * (1 / 1.2) - 0.5; s = int(np.floor(p)); fr = p - s
        else:
            p = f32((d + 0.5) * (1 / 1.2) - 0.5); s = int(np.floor(p)); fr = f32(p - f32(s))
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, src - 1); co[d] = coeffs(fr, style)
    return idx, co

def acc(terms, ws, mode):
    # terms: list of 4 f32 arrays, ws: list of 4 broadcastable f32 weight arrays
    if mode == "seq":
        r = terms[0] * ws[0]
        for k in range(1, 4): r = (r + terms[k] * ws[k]).astype(f32)
        return r
    if mode == "fma":
        r = (terms[0] * ws[0]).astype(f32)
        for k in range(1, 4): r = (terms[k].astype(np.float64) * ws[k].astype(np.float64) + r.astype(np.float64)).astype(f32)
        return r
    if mode == "pair":
        return ((terms[0] * ws[0] + terms[1] * ws[1]).astype(f32) + (terms[2] * ws[2] + terms[3] * ws[3]).astype(f32)).astype(f32)
    if mode == "f64":
        return sum(terms[k].astype(np.float64) * ws[k].astype(np.float64) for k in range(4))

rng = np.random.default_rng(1)
imgs = [rng.integers(0, 256, (540, 960, 3), dtype=np.uint8) for _ in range(3)]
refs = [cv2.resize(i, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC).astype(np.int64) for i in imgs]
for style, fstyle, order, mode in itertools.product(("cv", "cv4", "exact"), ("f64", "f32"), ("hv", "vh"), ("seq", "fma", "pair", "f64")):
    bad = 0
    for img, ref in zip(imgs, refs):
        h, w = img.shape[:2]; dh, dw = ref.shape[:2]
        yi, yc = axis(dh, h, style, fstyle); xi, xc = axis(dw, w, style, fstyle)
        im = img.astype(f32)
        if order == "hv":
            hor = acc([im[:, xi[:, k]] for k in range(4)], [xc[:, k][None, :, None] for k in range(4)], mode)
            if mode == "f64": hor = hor  # keep f64
            v = acc([hor[yi[:, k]] for k in range(4)], [yc[:, k][:, None, None] for k in range(4)], mode)
        else:
            ver = acc([im[yi[:, k]] for k in range(4)], [yc[:, k][:, None, None] for k in range(4)], mode)
            v = acc([ver[:, xi[:, k]] for k in range(4)], [xc[:, k][None, :, None] for k in range(4)], mode)
        o = np.clip(np.rint(v), 0, 255).astype(np.int64)
        bad += int((o != ref).sum())
    print(style, fstyle, order, mode, bad, flush=True)

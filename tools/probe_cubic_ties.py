"""Probe: exact rational evaluation of the 1.2x cubic; what does cv2 (IPP path) do at exact .5 ties?"""
import numpy as np, cv2
from fractions import Fraction as F

def coeffs_exact(k):        # x = k/12, returns the 4 coefficients * 6912 as integers
    A = F(-3, 4); x = F(k, 12)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c3 = 1 - c0 - c1 - c2
    out = [c * 6912 for c in (c0, c1, c2, c3)]
    assert all(o.denominator == 1 for o in out)
    return [int(o) for o in out]

def axis(dst, src):
    idx = np.empty((dst, 4), np.int64); co = np.empty((dst, 4), np.int64)
    for d in range(dst):
        # p = (d + 0.5)/1.2 - 0.5 = (10 d + 5)/12 - 6/12 = (10 d - 1)/12
        num = 10 * d - 1
        s = num // 12; k = num - 12 * s
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, src - 1); co[d] = coeffs_exact(k)
    return idx, co

rng = np.random.default_rng(1)
tot = 0; res = {"up": 0, "down": 0}; non_tie_bad = 0; evenodd = {"even": 0, "odd": 0}
for trial in range(6):
    img = rng.integers(0, 256, (540, 960, 3), dtype=np.uint8)
    ref = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC).astype(np.int64)
    h, w = img.shape[:2]; dh, dw = ref.shape[:2]
    yi, yc = axis(dh, h); xi, xc = axis(dw, w)
    im = img.astype(np.int64)
    hor = sum(im[:, xi[:, k]] * xc[:, k][None, :, None] for k in range(4))
    V = sum(hor[yi[:, k]] * yc[:, k][:, None, None] for k in range(4))       # value * 6912^2
    D = 6912 * 6912
    fl = V // D; rem = V - fl * D
    tie = (2 * rem == D)
    exact_round = np.clip(np.where(2 * rem > D, fl + 1, fl), 0, 255)         # ties -> down here
    inrange = (fl >= 0) & (fl < 255)
    t = tie & inrange
    up = (ref == fl + 1) & t; down = (ref == fl) & t
    res["up"] += int(up.sum()); res["down"] += int(down.sum()); tot += int(t.sum())
    evenodd["even"] += int((ref[t] % 2 == 0).sum()); evenodd["odd"] += int((ref[t] % 2 == 1).sum())
    non_tie_bad += int(((ref != exact_round) & ~tie).sum())
print("ties", tot, res, evenodd, "non-tie mismatches", non_tie_bad)

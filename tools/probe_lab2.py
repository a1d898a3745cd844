"""Probe 2: BGR2Lab integer-table model with OpenCV's own cubeRoot algorithm (float32 in, double rational polynomial)."""
import numpy as np, cv2
f32 = np.float32
gamma_shift, lab_shift = 3, 12
lab_shift2 = lab_shift + gamma_shift

def cv_cuberoot(value):
    value = np.asarray(value, f32)
    bits = value.view(np.int32).astype(np.int64)
    ix = bits & 0x7fffffff
    ex = (ix >> 23) - 127
    shx = np.fmod(ex, 3)              # C remainder (sign of dividend)
    shx = shx - np.where(shx >= 0, 3, 0)
    ex = (ex - shx) // 3              # exact division
    frb = ((ix & ((1 << 23) - 1)) | ((shx + 127) << 23)).astype(np.int32)
    fr = frb.view(f32).astype(np.float64)
    num = ((((45.2548339756803022511987494 * fr + 192.2798368355061050458134625) * fr + 119.1654824285581628956914143) * fr
            + 13.43250139086239872172837314) * fr + 0.1636161226585754240958355063)
    den = ((((14.80884093219134573786480845 * fr + 151.9714051044435648658557668) * fr + 168.5254414101568283957668343) * fr
            + 33.9905941350215598754191872) * fr + 1.0)
    r = (num / den).astype(f32)
    rb = r.view(np.int32).astype(np.int64) + (ex << 23)
    out = rb.astype(np.int32).view(f32)
    return np.where(value == 0, f32(0), out)

i = np.arange(256)
x = (i / 255.0)
g = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
gtab = np.rint(255.0 * (1 << gamma_shift) * g).astype(np.int64)
n = 256 * 3 // 2 * (1 << gamma_shift)
scale = f32(1) / (f32(255) * f32(1 << gamma_shift))
xx = (scale * np.arange(n, dtype=f32)).astype(f32)
lthresh = f32(216) / f32(24389); lscale = f32(841) / f32(108); lbias = f32(16) / f32(116)
lin = (xx.astype(np.float64) * np.float64(lscale) + np.float64(lbias)).astype(f32)     # mulAdd: fused, one rounding (float64 holds it exactly enough)
ctab = np.rint(f32(1 << lab_shift2) * np.where(xx < lthresh, lin, cv_cuberoot(xx))).astype(np.int64)
M = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
wp = np.array([0.950456, 1.0, 1.088754])
Cf = np.rint((1 << lab_shift) * M / wp[:, None]).astype(np.int64)
def descale(v, n): return (v + (1 << (n - 1))) >> n
a = np.arange(1 << 24, dtype=np.uint32)
img = np.stack([a & 255, (a >> 8) & 255, (a >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
ref = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
B, G, R = (gtab[img[..., k]] for k in range(3))
f = [ctab[descale(R * Cf[r, 0] + G * Cf[r, 1] + B * Cf[r, 2], lab_shift)] for r in range(3)]
Lscale = (116 * 255 + 50) // 100
Lshift = -((16 * 255 * (1 << lab_shift2) + 50) // 100)
L = descale(Lscale * f[1] + Lshift, lab_shift2)
A = descale(500 * (f[0] - f[1]) + 128 * (1 << lab_shift2), lab_shift2)
Bq = descale(200 * (f[1] - f[2]) + 128 * (1 << lab_shift2), lab_shift2)
mine = np.clip(np.stack([L, A, Bq], -1), 0, 255).astype(np.uint8)
d = np.abs(mine.astype(int) - ref.astype(int))
for k, nm in enumerate("Lab"):
    print(nm, "mismatch", int((d[..., k] > 0).sum()), "max", int(d[..., k].max()))

# ---- per-entry fit: which table value does cv2 use? ----
xx64 = np.arange(n) / (255.0 * 8)
ctab1 = np.rint((1 << lab_shift2) * np.where(xx64 < 216 / 24389.0, xx64 * (841 / 108.0) + 16 / 116.0, np.cbrt(xx64))).astype(np.int64)
diff = np.nonzero(ctab1 != ctab)[0]
print("entries where the two tables differ:", len(diff), "first", diff[:20], "in linear part:", int((xx64[diff] < 216 / 24389.0).sum()))
idx = [descale(R * Cf[r, 0] + G * Cf[r, 1] + B * Cf[r, 2], lab_shift).astype(np.int16) for r in range(3)]
def mism(tab, sel):
    fx, fy, fz = (tab[idx[r][sel].astype(np.int64)] for r in range(3))
    L = descale(Lscale * fy + Lshift, lab_shift2); A = descale(500 * (fx - fy) + (128 << lab_shift2), lab_shift2); Bq = descale(200 * (fy - fz) + (128 << lab_shift2), lab_shift2)
    m = np.clip(np.stack([L, A, Bq], -1), 0, 255)
    return int((m != ref[sel]).sum())
best = ctab1.copy()
# candidates: differing entries + entries involved in v1 mismatches
f1 = [ctab1[idx[r].astype(np.int64)] for r in range(3)]
A1 = np.clip(descale(500 * (f1[0] - f1[1]) + (128 << lab_shift2), lab_shift2), 0, 255); B1 = np.clip(descale(200 * (f1[1] - f1[2]) + (128 << lab_shift2), lab_shift2), 0, 255)
badm = (A1 != ref[..., 1]) | (B1 != ref[..., 2])
cand = set(diff.tolist())
for r in range(3): cand |= set(np.unique(idx[r][badm]).tolist())
print("candidates", len(cand))
changed = {}
for e in sorted(cand):
    sel = (idx[0] == e) | (idx[1] == e) | (idx[2] == e)
    if not sel.any(): continue
    scores = {}
    for dlt in (-1, 0, 1):
        t = best.copy(); t[e] += dlt
        scores[dlt] = mism(t, sel)
    dl = min(scores, key=lambda k: (scores[k], abs(k)))
    if dl != 0:
        best[e] += dl; changed[e] = (dl, scores[0], scores[dl])
print("changed entries (entry: delta, mismatches before, after):", changed)
print("total mismatches with fitted table:", mism(best, np.ones(idx[0].shape, bool)))
print("fitted == cv-algorithm table at changed entries:", {e: bool(best[e] == ctab[e]) for e in changed})

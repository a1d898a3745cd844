"""Probe: does cv2.cvtColor(BGR2Lab) on uint8 follow OpenCV's integer-table model (RGB2Lab_b without interpolation)?"""
import numpy as np, cv2
gamma_shift, lab_shift = 3, 12
lab_shift2 = lab_shift + gamma_shift
i = np.arange(256)
x = (i / 255.0)
g = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
gtab = np.rint(255.0 * (1 << gamma_shift) * g).astype(np.int64)
j = np.arange(256 * 3 // 2 * (1 << gamma_shift))
xx = j / (255.0 * (1 << gamma_shift))
ctab = np.rint((1 << lab_shift2) * np.where(xx < 216 / 24389.0, xx * (841 / 108.0) + 16 / 116.0, np.cbrt(xx))).astype(np.int64)
M = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
wp = np.array([0.950456, 1.0, 1.088754])
Cf = np.rint((1 << lab_shift) * M / wp[:, None]).astype(np.int64)   # rows X,Y,Z; cols R,G,B
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

bad = np.argwhere(d[..., 1] > 0)
idx = [descale(R * Cf[r, 0] + G * Cf[r, 1] + B * Cf[r, 2], lab_shift) for r in range(3)]
ix = idx[0][bad[:, 0], bad[:, 1]]; iy = idx[1][bad[:, 0], bad[:, 1]]
print("a-mismatch: distinct X idx", len(np.unique(ix)), "distinct Y idx", len(np.unique(iy)))
ux, cx = np.unique(ix, return_counts=True); print("top X idx", sorted(zip(cx, ux))[-8:])
uy, cy = np.unique(iy, return_counts=True); print("top Y idx", sorted(zip(cy, uy))[-8:])
print("sign of (mine - ref) for a:", np.unique((mine.astype(int) - ref.astype(int))[..., 1][d[..., 1] > 0], return_counts=True))
# the exact pre-rounding remainder at mismatches
va = 500 * (f[0] - f[1]) + 128 * (1 << lab_shift2)
rem = (va & ((1 << lab_shift2) - 1))[d[..., 1] > 0]
print("remainder/32768 at a mismatches: min %.4f max %.4f" % (rem.min() / 32768, rem.max() / 32768))
print("sample colours (B,G,R):", img[bad[:5, 0], bad[:5, 1]].tolist())

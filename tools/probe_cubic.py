"""Probe: is cv2.resize(INTER_CUBIC, fx=fy=1.2) on uint8 reproduced by the generic fixed-point model
(11-bit coefficients, int32 horizontal pass, (v + 2^21) >> 22 vertical cast)?"""
import numpy as np, cv2, sys

def cubic_coeffs(x):
    A = -0.75
    x = np.float32(x)
    c = np.empty(4, np.float32)
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c[3] = np.float32(1) - c[0] - c[1] - c[2]
    return c

def axis(dst, src, fx_scale):
    # scale = 1/fx as double; sx = floor((d+0.5)*scale - 0.5) with float fx
    scale = 1.0 / fx_scale
    idx = np.empty((dst, 4), np.int64); co = np.empty((dst, 4), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f)); f = np.float32(f - np.float32(s))
        c = cubic_coeffs(f)
        ci = np.rint(c * np.float32(2048)).astype(np.int64)
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, src - 1)
        co[d] = ci
    return idx, co

def model(img, f=1.2, variant=0):
    h, w = img.shape[:2]
    dh, dw = int(round(h * f)), int(round(w * f))
    yi, yc = axis(dh, h, f); xi, xc = axis(dw, w, f)
    im = img.astype(np.int64)
    if im.ndim == 2: im = im[..., None]
    hor = sum(im[:, xi[:, k]] * xc[:, k][None, :, None] for k in range(4))      # h x dw x c
    if variant == 0:
        v = sum(hor[yi[:, k]] * yc[:, k][:, None, None] for k in range(4))
        out = np.clip((v + (1 << 21)) >> 22, 0, 255)
    else:   # SIMD float path: rint(sum(float(row)*beta*scale))
        sc = np.float32(1.0 / (2048 * 2048))
        acc = np.zeros((dh, dw, im.shape[2]), np.float32)
        for k in range(4):
            b = (yc[:, k].astype(np.float32) * sc)[:, None, None]
            acc = (hor[yi[:, k]].astype(np.float32) * b + acc).astype(np.float32)
        out = np.clip(np.rint(acc), 0, 255)
    return out.astype(np.uint8).reshape(dh, dw, *img.shape[2:])

rng = np.random.default_rng(0)
for shape in ((270, 480, 3), (135, 241, 3), (100, 90), (1080, 1920, 3)):
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC)
    for v in (0, 1):
        m = model(img, 1.2, v)
        if m.shape != ref.shape: print(shape, "shape", m.shape, ref.shape); continue
        d = np.abs(m.astype(int) - ref.astype(int))
        print(shape, "variant", v, "mismatch", int((d > 0).sum()), "of", d.size, "max", int(d.max()))

print("--- IPP off")
cv2.ipp.setUseIPP(False)
img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
ref = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC)
for v in (0, 1):
    d = np.abs(model(img, 1.2, v).astype(int) - ref.astype(int)); print("variant", v, int((d > 0).sum()), int(d.max()))
cv2.setUseOptimized(False)
ref2 = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC)
for v in (0, 1):
    d = np.abs(model(img, 1.2, v).astype(int) - ref2.astype(int)); print("noopt variant", v, int((d > 0).sum()), int(d.max()))
cv2.setUseOptimized(True); cv2.ipp.setUseIPP(True)
ref3 = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_CUBIC)
print("ipp vs noipp differ:", int((ref3 != ref).sum()), " noopt vs noipp:", int((ref2 != ref).sum()))

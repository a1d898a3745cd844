"""Closed-form numpy models of the OpenCV primitives on the hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  No cv2 import here: these
are the specifications the CUDA kernels are checked against, and they are
themselves checked bit-exactly against cv2 4.13.0 in
``tests/test_oracle_cvmodel.py`` (exhaustive 2^24 colours for the colour
conversions, random images for resize / morphology).

The reference reaches these primitives at (paths relative to the reference
root): ``cv2.cvtColor`` unscreen/colorfiltering/agent.py:310,352 and
unscreen/utils/fgfuncs.py:36,39,55,100-101,109,129,136; ``cv2.resize``
unscreen/colorfiltering/agent.py:315-316,342, unscreen/trimap/agent.py:52,59;
``cv2.dilate/erode/getStructuringElement`` unscreen/utils/maskprocess.py:16-18,
31-33; ``cv2.inRange`` unscreen/utils/fgfuncs.py:60.
"""
import numpy as np

# --------------------------------------------------------------------------
# colour conversions
# --------------------------------------------------------------------------

_HSV_SHIFT = 12


def hsv_div_tables():
    """sdiv/hdiv fixed-point reciprocal tables of cv2's 8-bit BGR2HSV
    (round-half-even construction, H range 180)."""
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, np.int32)
    hdiv = np.zeros(256, np.int32)
    sdiv[1:] = np.rint((255 << _HSV_SHIFT) / i).astype(np.int32)
    hdiv[1:] = np.rint((180 << _HSV_SHIFT) / (6.0 * i)).astype(np.int32)
    return sdiv, hdiv


_SDIV, _HDIV = hsv_div_tables()


def bgr2hsv(img):
    """cv2.cvtColor(img, COLOR_BGR2HSV) for uint8 (H in [0,179])."""
    img = np.asarray(img)
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(b, g), r)
    mn = np.minimum(np.minimum(b, g), r)
    d = v - mn
    s = (d * _SDIV[v] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    # hue numerator, priority r, then g, then b
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (h * _HDIV[d] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT  # arithmetic shift
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


# cv2's 8-bit BGR2Lab (sRGB, D65): integer tables.  gamma table: 3 fractional
# bits; cube-root table: 15 fractional bits over 3072 entries; 12-bit matrix.
_LAB_GAMMA_SHIFT = 3
_LAB_SHIFT = 12
_LAB_SHIFT2 = _LAB_SHIFT + _LAB_GAMMA_SHIFT


def lab_tables():
    """(gamma[256], cbrt[3072], coeffs[3][3] for R,G,B) of cv2's RGB2Lab_b.
    cv2 fills the cube-root table with its own float cube-root routine; the
    correctly rounded values below differ from it in exactly two entries that
    any colour reaches (49 and 628), fixed here by hand -- pinned by the
    exhaustive 2^24 test in tests/test_oracle_cvmodel.py."""
    i = np.arange(256) / 255.0
    g = np.where(i <= 0.04045, i / 12.92, ((i + 0.055) / 1.055) ** 2.4)
    gamma = np.rint(255.0 * (1 << _LAB_GAMMA_SHIFT) * g).astype(np.int64)
    x = np.arange(256 * 3 // 2 * (1 << _LAB_GAMMA_SHIFT)) / (255.0 * (1 << _LAB_GAMMA_SHIFT))
    f = np.where(x < 216 / 24389.0, x * (841 / 108.0) + 16 / 116.0, np.cbrt(x))
    cbrt = np.rint((1 << _LAB_SHIFT2) * f).astype(np.int64)
    cbrt[49] -= 1
    cbrt[628] += 1
    m = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
    wp = np.array([0.950456, 1.0, 1.088754])
    coeffs = np.rint((1 << _LAB_SHIFT) * m / wp[:, None]).astype(np.int64)
    return gamma, cbrt, coeffs


_LAB_GAMMA, _LAB_CBRT, _LAB_COEFFS = lab_tables()


def bgr2lab(img):
    """cv2.cvtColor(img, COLOR_BGR2Lab) for uint8."""
    img = np.asarray(img)
    b, g, r = (_LAB_GAMMA[img[..., k]] for k in range(3))

    def descale(v, n):
        return (v + (1 << (n - 1))) >> n

    fx, fy, fz = (_LAB_CBRT[descale(r * _LAB_COEFFS[k, 0] + g * _LAB_COEFFS[k, 1] + b * _LAB_COEFFS[k, 2], _LAB_SHIFT)] for k in range(3))
    lscale = (116 * 255 + 50) // 100
    lshift = -((16 * 255 * (1 << _LAB_SHIFT2) + 50) // 100)
    L = descale(lscale * fy + lshift, _LAB_SHIFT2)
    a = descale(500 * (fx - fy) + (128 << _LAB_SHIFT2), _LAB_SHIFT2)
    bb = descale(200 * (fy - fz) + (128 << _LAB_SHIFT2), _LAB_SHIFT2)
    return np.clip(np.stack([L, a, bb], -1), 0, 255).astype(np.uint8)


def bgr2gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8 (15-bit coefficients)."""
    img = np.asarray(img)
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def hsv2bgr_f32(hsv):
    """The float32 value cv2's HSV2BGR computes before the final cast to u8
    (scaled by 255).  cv2 is not self-consistent about that cast: whole-image
    (SIMD) calls truncate, 1-pixel (scalar tail) calls round-half-even
     -- hence the reference's composites are +-1 LSB."""
    hsv = np.asarray(hsv)
    f = np.float32
    h = hsv[..., 0].astype(f) * f(6.0 / 180.0)
    s = hsv[..., 1].astype(f) * f(1.0 / 255.0)
    v = hsv[..., 2].astype(f) * f(1.0 / 255.0)
    h = np.fmod(h, f(6.0)).astype(f)
    sec = np.floor(h).astype(np.int32)
    fr = (h - sec.astype(f)).astype(f)
    bad = (sec < 0) | (sec >= 6)
    sec = np.where(bad, 0, sec)
    fr = np.where(bad, f(0), fr).astype(f)
    one = f(1.0)
    tab = np.stack([
        v,
        (v * (one - s)).astype(f),
        (v * (one - (s * fr).astype(f))).astype(f),
        (v * (one - (s * (one - fr)).astype(f))).astype(f),
    ], axis=-1)
    idx = _SECTOR[sec]  # (...,3)
    out = np.take_along_axis(tab, idx, axis=-1)
    grey = (hsv[..., 1] == 0)
    out = np.where(grey[..., None], v[..., None], out)
    return (out * f(255.0)).astype(f)


def hsv2bgr(hsv, rounding="trunc"):
    """cv2.cvtColor(hsv, COLOR_HSV2BGR) for uint8; ``rounding`` is 'trunc'
    (whole images) or 'rint' (single pixels)."""
    x = hsv2bgr_f32(hsv)
    if rounding == "trunc":
        return np.clip(x, 0, 255).astype(np.uint8)
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def in_range(img, lo, hi):
    """cv2.inRange: 255 where lo<=img<=hi on every channel (inclusive)."""
    img = np.asarray(img).astype(np.int32)
    lo = np.asarray(lo).astype(np.int32)
    hi = np.asarray(hi).astype(np.int32)
    ok = np.all((img >= lo) & (img <= hi), axis=-1)
    return (ok * 255).astype(np.uint8)


# --------------------------------------------------------------------------
# resize
# --------------------------------------------------------------------------

_COEF_BITS = 11
_COEF_SCALE = 1 << _COEF_BITS


def _linear_axis(dst, src, horizontal):
    """index / weight tables of cv2's fixed-point bilinear resize for one axis.
    Horizontal: fraction reset at both borders.  Vertical: index clamp only."""
    inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    i0 = np.floor(f).astype(np.int32)
    fr = (f - i0.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = i0 < 0
        i0 = np.where(lo, 0, i0)
        fr = np.where(lo, np.float32(0), fr)
        hi = i0 >= src - 1
        i0 = np.where(hi, src - 1, i0)
        fr = np.where(hi, np.float32(0), fr).astype(np.float32)
        i1 = np.minimum(i0 + 1, src - 1)
    else:
        i1 = np.clip(i0 + 1, 0, src - 1)
        i0 = np.clip(i0, 0, src - 1)
    w0 = np.rint((np.float32(1.0) - fr).astype(np.float32) * np.float32(_COEF_SCALE)).astype(np.int32)
    w1 = np.rint(fr * np.float32(_COEF_SCALE)).astype(np.int32)
    return i0, i1, w0, w1


def resize_linear(src, dw, dh):
    """cv2.resize(src, (dw, dh)) with the default INTER_LINEAR on uint8.
    Exact 2x down-scaling in both axes silently becomes INTER_AREA (rounded
    2x2 mean)."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    if dw == sw and dh == sh:
        return src.copy()
    if sw == 2 * dw and sh == 2 * dh:
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8)
    x0, x1, a0, a1 = _linear_axis(dw, sw, True)
    y0, y1, b0, b1 = _linear_axis(dh, sh, False)
    s = src.astype(np.int32)
    if s.ndim == 3:
        a0 = a0[None, :, None]
        a1 = a1[None, :, None]
        b0 = b0[:, None, None]
        b1 = b1[:, None, None]
    else:
        a0 = a0[None, :]
        a1 = a1[None, :]
        b0 = b0[:, None]
        b1 = b1[:, None]
    rows = s[:, x0] * a0 + s[:, x1] * a1  # (sh, dw[, c]) scaled by 2048
    r0 = rows[y0]
    r1 = rows[y1]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def _linear_axis_f(dst, src, scale):
    """index / weight tables of cv2's bilinear resize of CV_64F data for one axis (fraction reset at both borders).  This
    depth goes to IPP (ippiResizeLinear_64f), whose weights are doubles: the model agrees with cv2 4.13 to ~1e-12, with
    float32 weights (cv2's own generic loop) it would be off by ~1e-3."""
    d = np.arange(dst, dtype=np.float64)
    f = (d + 0.5) * scale - 0.5
    i0 = np.floor(f).astype(np.int64)
    fr = f - i0
    lo = i0 < 0
    i0 = np.where(lo, 0, i0)
    fr = np.where(lo, 0.0, fr)
    hi = i0 >= src - 1
    i0 = np.where(hi, src - 1, i0)
    fr = np.where(hi, 0.0, fr)
    i1 = np.minimum(i0 + 1, src - 1)
    return i0, i1, 1.0 - fr, fr


def scaled_size(n, f):
    """cv2.resize(..., (0, 0), fx=f): saturate_cast<int>(n * f) = round half to even"""
    return int(np.rint(n * f))


def resize_linear_f64(src, dw, dh, fx=None, fy=None):
    """cv2.resize of a float64 single-channel array with the default INTER_LINEAR (to ~1e-12, see _linear_axis_f).
    ``fx`` / ``fy`` given (the ``(0, 0), fx=, fy=`` form): the sampling scale is 1 / fx whatever the rounded size is, and
    an exact 2 x 2 scale becomes INTER_AREA -- 2 x 2 means, and for an odd source the last row / column averages the
    pixels that exist (in float32, as cv2's tail loop does)."""
    src = np.asarray(src, np.float64)
    sh, sw = src.shape
    scale_x = 1.0 / fx if fx is not None else 1.0 / (float(dw) / sw)
    scale_y = 1.0 / fy if fy is not None else 1.0 / (float(dh) / sh)
    if dw == sw and dh == sh:
        return src.copy()
    if abs(scale_x - 2) < 2.2e-16 and abs(scale_y - 2) < 2.2e-16:
        out = np.zeros((dh, dw), np.float64)
        fw, fh = sw // 2, sh // 2
        out[:fh, :fw] = (src[0:2 * fh:2, 0:2 * fw:2] + src[0:2 * fh:2, 1:2 * fw:2] + src[1:2 * fh:2, 0:2 * fw:2] + src[1:2 * fh:2, 1:2 * fw:2]) * 0.25
        for y in range(dh):
            for x in range(dw):
                if y < fh and x < fw:
                    continue
                blk = src[2 * y:2 * y + 2, 2 * x:2 * x + 2]
                out[y, x] = np.float32(blk.sum()) / np.float32(blk.size) if blk.size else 0.0
        return out
    x0, x1, a0, a1 = _linear_axis_f(dw, sw, scale_x)
    y0, y1, b0, b1 = _linear_axis_f(dh, sh, scale_y)
    rows = src[:, x0] * a0[None, :] + src[:, x1] * a1[None, :]
    return rows[y0] * b0[:, None] + rows[y1] * b1[:, None]


def resize_nearest(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=INTER_NEAREST)."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    ifx = 1.0 / (float(dw) / float(sw))
    ify = 1.0 / (float(dh) / float(sh))
    xs = np.minimum(np.floor(np.arange(dw) * ifx).astype(np.int64), sw - 1)
    ys = np.minimum(np.floor(np.arange(dh) * ify).astype(np.int64), sh - 1)
    return src[ys][:, xs].copy()


# --------------------------------------------------------------------------
# geometric pre-steps of the replacement path
# --------------------------------------------------------------------------


def warp_translate(src, dx, dy):
    """cv2.warpAffine(src, np.float32([[1, 0, dx], [0, 1, dy]]), (W, H)) on
    uint8 (default INTER_LINEAR, BORDER_CONSTANT 0) -- unscreen/utils/
    imgprocess.py:55-64 (``shift_fg``), SURVEY.md A.8.

    The matrix is stored as float32, inverted in double (dst(x, y) <- src(x -
    dx, y - dy)), and evaluated in fixed point: 10 fractional bits per
    coordinate term, each term rounded half-to-even on its own, +16 and >> 5
    to a 5-bit sub-pixel fraction; the 2x2 weights are 32 * (32 - fx | fx) *
    (32 - fy | fy) (exact in the 15-bit remap table, so the table's
    sum-correction never fires); dst = (sum(w * tap) + 16384) >> 15; taps
    outside the image read 0."""
    src = np.asarray(src)
    h, w = src.shape[:2]
    b1 = -float(np.float32(dx))
    b2 = -float(np.float32(dy))
    X = (int(np.rint(b1 * 1024.0)) + 16 + 1024 * np.arange(w, dtype=np.int64)) >> 5
    Y = (np.rint((np.arange(h, dtype=np.float64) + b2) * 1024.0).astype(np.int64) + 16) >> 5
    # cv2 stores the integer part as int16 (saturating)
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    fx = (X & 31)[None, :]
    fy = (Y & 31)[:, None]
    s = src.astype(np.int64)
    if s.ndim == 2:
        s = s[..., None]

    def tap(yy, xx):
        ok = ((yy >= 0) & (yy < h))[:, None] & ((xx >= 0) & (xx < w))[None, :]
        v = s[np.clip(yy, 0, h - 1)][:, np.clip(xx, 0, w - 1)]
        return v * ok[..., None]

    acc = (tap(sy, sx) * (32 * (32 - fx) * (32 - fy))[..., None] + tap(sy, sx + 1) * (32 * fx * (32 - fy))[..., None]
           + tap(sy + 1, sx) * (32 * (32 - fx) * fy)[..., None] + tap(sy + 1, sx + 1) * (32 * fx * fy)[..., None])
    out = ((acc + 16384) >> 15).astype(np.uint8)
    return out.reshape(src.shape)


def cubic_axis(dst, src, factor, off=0, count=None):
    """tap indices (clamped) and float32 weights of the bicubic (a = -0.75)
    up-scale by ``factor`` for destination positions off .. off+count-1.
    Coefficients are evaluated in float64 from the float64 source position and
    rounded to float32 once."""
    count = dst - off if count is None else count
    d = np.arange(off, off + count, dtype=np.float64)
    p = (d + 0.5) / float(factor) - 0.5
    s = np.floor(p)
    x = p - s
    A = -0.75
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c3 = 1 - c0 - c1 - c2
    co = np.stack([c0, c1, c2, c3], 1).astype(np.float32)
    idx = np.clip(s.astype(np.int64)[:, None] + np.arange(-1, 3)[None, :], 0, src - 1)
    return idx, co


def rescale_size(n, factor):
    """cv2.resize(..., fx=factor): dsize = saturate_cast<int>(n * factor)
    (round half to even)."""
    return int(np.rint(n * float(factor)))


def fma32(a, b, c):
    """float32 fused multiply-add, correctly rounded: the product of two
    float32 is exact in float64; the float64 sum is made exact with TwoSum,
    rounded to odd, and only then rounded to float32 (round-to-odd keeps the
    second rounding honest: float64 carries more than two extra bits)."""
    p = np.asarray(a, np.float32).astype(np.float64) * np.asarray(b, np.float32).astype(np.float64)
    c = np.asarray(c, np.float32).astype(np.float64)
    p, c = np.broadcast_arrays(p, c)
    t = p + c
    bb = t - p
    e = (p - (t - bb)) + (c - bb)
    fix = (e != 0) & ((t.view(np.int64) & 1) == 0)
    t = np.where(fix, np.nextafter(t, np.where(e > 0, np.inf, -np.inf)), t)
    return t.astype(np.float32)


def resize_cubic_crop(src, factor):
    """cv2.resize(src, None, fx=factor, fy=factor, interpolation=INTER_CUBIC)
    followed by the centre crop back to the source size --
    unscreen/utils/imgprocess.py:40-52 (``rescale_fg``), factor >= 1.

    PARITY NOTE.  cv2 4.13.0 in the build container dispatches this call to
    Intel IPP's closed-source float32 cubic; with IPP switched off
    (cv2.ipp.setUseIPP(False)) the same call runs OpenCV's 11-bit fixed-point
    kernel and 8 % of the values differ by 1.  The model here is the one the
    default (IPP) path follows: separable float bicubic, a = -0.75, replicated
    borders, round half to even, saturate.  It is evaluated in float32 with a
    fixed order (horizontal then vertical; taps -1, 0, +1, +2: one multiply,
    then three fused multiply-adds), which agrees with the IPP result
    everywhere except at values whose exact result is within float32 rounding
    noise of a .5 tie: < 2e-5 of the values, off by 1 (tools/probe_cubic_*.py;
    exact rational evaluation shows IPP rounds true ties to even, and none of
    48 float32 evaluation orders reproduces it on the near-ties).  The CUDA
    kernel is bit-exact against THIS model."""
    src = np.asarray(src)
    h, w = src.shape[:2]
    dh, dw = rescale_size(h, factor), rescale_size(w, factor)
    h_off = int((dh - h) / 2)
    w_off = int((dw - w) / 2)
    yi, yc = cubic_axis(dh, h, factor, h_off, h)
    xi, xc = cubic_axis(dw, w, factor, w_off, w)
    s = src.astype(np.float32)
    if s.ndim == 2:
        s = s[..., None]
    hor = s[:, xi[:, 0]] * xc[:, 0][None, :, None]
    for k in range(1, 4):
        hor = fma32(s[:, xi[:, k]], xc[:, k][None, :, None], hor)
    v = hor[yi[:, 0]] * yc[:, 0][:, None, None]
    for k in range(1, 4):
        v = fma32(hor[yi[:, k]], yc[:, k][:, None, None], v)
    assert hor.dtype == np.float32 and v.dtype == np.float32
    out = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return out.reshape(src.shape)


# --------------------------------------------------------------------------
# morphology
# --------------------------------------------------------------------------


def ellipse_se(k):
    """cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)) as a 0/1 array; the
    anchor is (k//2, k//2).  k=3 is the 5-tap cross."""
    r = k // 2
    c = k // 2
    inv_r2 = 1.0 / (float(r) * r) if r else 0.0
    se = np.zeros((k, k), np.uint8)
    for i in range(k):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            j1 = max(c - dx, 0)
            j2 = min(c + dx + 1, k)
            se[i, j1:j2] = 1
    return se


def se_offsets(k):
    se = ellipse_se(k)
    a = k // 2
    return [(i - a, j - a) for i in range(k) for j in range(k) if se[i, j]]


def _morph_once(img, offs, dilate):
    h, w = img.shape
    pad = max(max(abs(dy), abs(dx)) for dy, dx in offs)
    fill = 0 if dilate else 255
    p = np.full((h + 2 * pad, w + 2 * pad), fill, np.uint8)
    p[pad:pad + h, pad:pad + w] = img
    out = None
    for dy, dx in offs:
        v = p[pad + dy:pad + dy + h, pad + dx:pad + dx + w]
        if out is None:
            out = v.copy()
        elif dilate:
            np.maximum(out, v, out=out)
        else:
            np.minimum(out, v, out=out)
    return out


def morph(img, k, iters, dilate):
    """cv2.dilate / cv2.erode(img, ellipse_se(k), iterations=iters) on a
    single-channel uint8 image (taps outside the image are ignored)."""
    img = np.ascontiguousarray(img)
    offs = se_offsets(k)
    for _ in range(iters):
        img = _morph_once(img, offs, dilate)
    return img


def dilate(img, k, iters):
    if img.ndim == 3:
        return np.stack([morph(img[..., c], k, iters, True) for c in range(img.shape[2])], -1)
    return morph(img, k, iters, True)


def erode(img, k, iters):
    if img.ndim == 3:
        return np.stack([morph(img[..., c], k, iters, False) for c in range(img.shape[2])], -1)
    return morph(img, k, iters, False)

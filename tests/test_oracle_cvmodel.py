"""oracle.cvmodel (closed-form numpy) against cv2 itself -- known-answer tests
for the third-party primitives the reference's hot path is made of."""
import numpy as np
import pytest

from oracle import cvmodel as M

cv2 = pytest.importorskip("cv2")


def all_colours():
    a = np.arange(1 << 24, dtype=np.uint32)
    return np.stack([a & 255, (a >> 8) & 255, (a >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)


def test_bgr2hsv_exhaustive():
    img = all_colours()
    assert np.array_equal(M.bgr2hsv(img), cv2.cvtColor(img, cv2.COLOR_BGR2HSV))


def test_bgr2gray_exhaustive():
    img = all_colours()
    assert np.array_equal(M.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_hsv2bgr_within_one_lsb():
    hsv = all_colours()[::3, ::3].copy()
    hsv[..., 0] %= 180
    d = np.abs(M.hsv2bgr(hsv).astype(int) - cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR).astype(int))
    assert d.max() <= 1          # cv2 is not self-consistent here
    assert (d > 0).mean() < 1e-2   # scalar-tail pixels round instead of truncating


@pytest.mark.parametrize("sh,sw,dh,dw", [(1080, 1920, 540, 960), (2160, 3840, 540, 960), (540, 960, 1080, 1920),
                                         (270, 480, 1080, 1920), (720, 1280, 540, 960), (960, 540, 480, 270),
                                         (333, 517, 200, 311), (200, 311, 333, 517), (541, 961, 270, 480)])
@pytest.mark.parametrize("ch", [1, 3])
def test_resize(sh, sw, dh, dw, ch):
    rng = np.random.default_rng(sh + dw + ch)
    s = rng.integers(0, 256, (sh, sw, ch) if ch > 1 else (sh, sw), dtype=np.uint8)
    assert np.array_equal(M.resize_linear(s, dw, dh), cv2.resize(s, (dw, dh)))
    assert np.array_equal(M.resize_nearest(s, dw, dh), cv2.resize(s, (dw, dh), interpolation=cv2.INTER_NEAREST))


@pytest.mark.parametrize("k,n", [(3, 2), (3, 5), (4, 2), (5, 3), (5, 10), (7, 10)])
def test_morphology(k, n):
    rng = np.random.default_rng(k * 100 + n)
    s = rng.integers(0, 256, (135, 240), dtype=np.uint8)
    ke = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    assert np.array_equal(ke, M.ellipse_se(k))
    assert np.array_equal(M.dilate(s, k, n), cv2.dilate(s, ke, iterations=n))
    assert np.array_equal(M.erode(s, k, n), cv2.erode(s, ke, iterations=n))


def test_cross_iterations_equal_diamond():
    rng = np.random.default_rng(5)
    s = rng.integers(0, 256, (60, 70), dtype=np.uint8)
    for r in (2, 5):
        offs = [(dy, dx) for dy in range(-r, r + 1) for dx in range(-r, r + 1) if abs(dy) + abs(dx) <= r]
        assert np.array_equal(M._morph_once(s, offs, True), M.dilate(s, 3, r))
        assert np.array_equal(M._morph_once(s, offs, False), M.erode(s, 3, r))


def test_inrange():
    rng = np.random.default_rng(9)
    s = rng.integers(0, 256, (50, 60, 3), dtype=np.uint8)
    lo, hi = np.array([10, 100, 20]), np.array([90, 255, 200])
    assert np.array_equal(M.in_range(s, lo, hi), cv2.inRange(s, lo, hi))


@pytest.mark.parametrize("shape", [(61, 83, 3), (64, 48), (135, 240, 3)])
def test_warp_translate(shape):
    """shift_fg's cv2.warpAffine translation: bit-exact, including shifts that no phase table of the
    survey's probe covered (arbitrary fractions, out-of-frame shifts, tiny shifts)."""
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    for dx, dy in [(3, -2), (0.5, 0.5), (0, 0), (-7.3, 4.9), (12.015625, -0.984375), (1e-3, 33.333), (-100, 2), (0.484375, 0.515625),
                   (5.7, 200), (2.25, -3.75), (-0.015, -0.016), (1e-30, -1e-30)]:
        ref = cv2.warpAffine(img, np.float32([[1, 0, dx], [0, 1, dy]]), (shape[1], shape[0]))
        assert np.array_equal(M.warp_translate(img, dx, dy), ref), (dx, dy)


CUBIC_TIE_TOL = 1e-4   # fraction of values allowed to differ (by one) from cv2's IPP float cubic; observed < 1.5e-5


@pytest.mark.parametrize("shape", [(270, 480, 3), (135, 241, 3), (100, 90)])
def test_resize_cubic_crop(shape):
    """rescale_fg's cv2.resize(INTER_CUBIC) + centre crop.  cv2 dispatches to Intel IPP's closed float kernel here; the
    model agrees except where the exact value sits within float rounding noise of a .5 tie (see the model's parity
    note): at most one LSB, on fewer than CUBIC_TIE_TOL of the values."""
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    h, w = shape[:2]
    for f in (1.2, 1.1, 1.5, 1.0):
        r = cv2.resize(img, None, fx=f, fy=f, interpolation=cv2.INTER_CUBIC)
        assert r.shape[:2] == (M.rescale_size(h, f), M.rescale_size(w, f))
        ho, wo = int((r.shape[0] - h) / 2), int((r.shape[1] - w) / 2)
        ref = r[ho:ho + h, wo:wo + w].astype(int)
        d = np.abs(M.resize_cubic_crop(img, f).astype(int) - ref)
        assert d.max() <= 1 and (d > 0).mean() <= CUBIC_TIE_TOL, (f, int((d > 0).sum()))


def test_bgr2lab_exhaustive():
    """cv2's 8-bit BGR2Lab over all 2^24 colours: bit-exact (color_correct, imgprocess.py:285-287)."""
    img = all_colours()
    assert np.array_equal(M.bgr2lab(img), cv2.cvtColor(img, cv2.COLOR_BGR2Lab))

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

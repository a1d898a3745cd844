"""Frame I/O mirror (reference unscreen/utils/fileio.py:31-62) against cv2.imread / cv2.imwrite.

JPEG decoders are not bit-identical (inverse DCT rounding, chroma up-sampling), so the contract is a tolerance, stated
here: decoded frames within a mean |difference| of 1.5 grey levels of cv2's (and, for 4:4:4 files, which have no chroma
up-sampling, within 1 level on average and 8 levels everywhere: measured 4); encoded frames decode (by cv2) to within 1 dB of the PSNR cv2.imwrite reaches
at the same quality.  Lossless formats go through cv2 and are identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")
pytest.importorskip("torchvision")

from video_unscreen_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def fio():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from video_unscreen_b200.unscreen.utils import fileio
    return fileio


def frames_of(n, h, w, seed):
    fr, _ = synth.green_clip(n, h, w, seed=seed)
    return [cv2.GaussianBlur(f, (0, 0), 1.2) for f in fr]       # camera-like: no pixel-level noise for the codec to chew on


def psnr(a, b):
    return 10 * np.log10(255.0 ** 2 / max(np.mean((a.astype(np.float64) - b) ** 2), 1e-12))


def test_layout_kernels_roundtrip(fio):
    from video_unscreen_b200 import ops
    rng = np.random.default_rng(0)
    for h, w in ((1, 1), (7, 13), (270, 481)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        pl = ops.bgr_to_planar_rgb(torch.from_numpy(img).cuda())
        assert np.array_equal(pl.cpu().numpy(), np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1)))
        assert np.array_equal(ops.planar_rgb_to_bgr(pl).cpu().numpy(), img)


@pytest.mark.parametrize("sampling", ["420", "444"])
def test_parallel_read_img_close_to_cv2(fio, tmp_path, sampling):
    frames = frames_of(5, 270, 480, seed=3)
    flag = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR", None)
    if sampling == "444" and flag is None:
        pytest.skip("this cv2 cannot write 4:4:4 JPEG")
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"{i:05d}.jpg")
        params = [cv2.IMWRITE_JPEG_QUALITY, 95] + ([flag, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444] if sampling == "444" else [])
        assert cv2.imwrite(p, f, params)
        paths.append(p)
    png = str(tmp_path / "mask.png")
    cv2.imwrite(png, frames[0])
    got = fio.parallel_read_img(paths + [png], batch=2)
    assert len(got) == 6 and all(isinstance(g, np.ndarray) and g.dtype == np.uint8 for g in got)
    assert np.array_equal(got[5], cv2.imread(png))
    for p, g in zip(paths, got):
        ref = cv2.imread(p)
        assert g.shape == ref.shape
        d = np.abs(g.astype(np.int16) - ref)
        assert d.mean() < 1.5, (sampling, d.mean(), d.max())
        if sampling == "444":
            assert d.mean() < 1.0 and d.max() <= 8, (d.mean(), d.max())
    dev = fio.parallel_read_img(paths[:2], on_device=True)
    assert all(t.is_cuda and t.shape == (270, 480, 3) for t in dev)
    assert np.array_equal(dev[1].cpu().numpy(), got[1])


def test_save_img_jpeg_quality_and_downscale(fio, tmp_path):
    img = frames_of(1, 360, 640, seed=9)[0]
    ours, ref = str(tmp_path / "ours.jpg"), str(tmp_path / "ref.jpg")
    fio.save_img(img, ours)
    cv2.imwrite(ref, img)
    a, b = cv2.imread(ours), cv2.imread(ref)
    assert a.shape == img.shape
    assert psnr(a, img) > psnr(b, img) - 1.0, (psnr(a, img), psnr(b, img))
    # down-scaled, lossless container: identical to the reference's cv2.resize + cv2.imwrite
    fio.save_img(img, str(tmp_path / "small.png"), downsacle=3)
    assert np.array_equal(cv2.imread(str(tmp_path / "small.png")), cv2.resize(img, (640 // 3, 360 // 3)))
    alpha = img[:, :, 1].copy()
    fio.save_img(alpha, str(tmp_path / "alpha.png"), downsacle=2)
    assert np.array_equal(cv2.imread(str(tmp_path / "alpha.png"), cv2.IMREAD_GRAYSCALE), cv2.resize(alpha, (320, 180)))
    fio.save_img(alpha, str(tmp_path / "alpha.jpg"))
    assert psnr(cv2.imread(str(tmp_path / "alpha.jpg"), cv2.IMREAD_GRAYSCALE), alpha) > 35

"""The C-ABI library loads and exports every symbol include/vu_b200.h declares;
host-side logic that needs no GPU (no compute calls here)."""
import ctypes
import os

import numpy as np
import pytest

from video_unscreen_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with `python __graft_entry__.py`"
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in vu_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table out of sync with the header"


def test_abi_version_and_status_strings():
    L = _lib.lib()
    assert L.vu_abi_version() == 1
    assert L.vu_status_string(0) == b"ok"
    assert b"unsupported" in L.vu_status_string(-2)
    with pytest.raises(_lib.VuError):
        _lib.check(-1)


def test_invalid_arguments_are_refused_without_a_gpu():
    L = _lib.lib()
    null = ctypes.c_void_p(0)
    assert L.vu_bgr2hsv_u8(null, null, 10, null) == -1
    assert L.vu_temporal_median_u8(null, 3, 10, null, null) == -1
    assert L.vu_morph_u8(null, null, 1, 4, 4, 3, 1, 0, null, 0, null) == -1


def test_product_package_does_not_import_the_oracle():
    import pathlib
    root = pathlib.Path(_lib.__file__).parent
    for p in root.rglob("*.py"):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, p


def test_host_helpers():
    from video_unscreen_b200.unscreen.utils import get_target_size
    from video_unscreen_b200.unscreen.utils.fgfuncs import bgr2hsv_pixel
    from oracle import cvmodel as M
    assert get_target_size(1080, 1920, 960) == (540, 960)
    assert get_target_size(2160, 3840, 960) == (540, 960)
    assert get_target_size(1920, 1080, 960) == (960, 540)
    rng = np.random.default_rng(0)
    px = rng.integers(0, 256, (500, 3), dtype=np.uint8)
    want = M.bgr2hsv(px[None])[0]
    got = np.array([bgr2hsv_pixel(p) for p in px])
    assert np.array_equal(got, want)


def test_agents_mirror_reference_signatures():
    import inspect
    from video_unscreen_b200.unscreen.bgmodel import BackgroundAgent
    from video_unscreen_b200.unscreen.colorfiltering import ColorFilteringAgent
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    assert list(inspect.signature(ColorFilteringAgent.__init__).parameters)[1:] == [
        "input_long_side", "bg_ncomp", "fg_ncomp", "max_num_samples", "color_prior_winsize", "use_opencv_gmm"]
    assert list(inspect.signature(TrimapAgent.__init__).parameters)[1:] == ["input_long_side", "kernelsize", "iters", "color_winsize"]
    assert list(inspect.signature(BackgroundAgent.__init__).parameters)[1:] == [
        "input_long_side", "dilation_ksize", "dilation_iters", "boundary_ksize", "boundary_iters", "pcov_ksize"]
    with pytest.raises(AssertionError):
        ColorFilteringAgent(input_long_side=960.0)
    with pytest.raises(NameError):
        BackgroundAgent().forward(None, None, method="nope")
    ag = ColorFilteringAgent()
    assert not ag.is_trained() and len(ag.bg_gmms) == 3 and len(ag.fg_gmms) == 3


def test_generated_lab_tables_match_the_oracle():
    """video_unscreen_b200/csrc/vu_lab_tables.inc (tools/gen_lab_tables.py) holds the tables of oracle.cvmodel.bgr2lab,
    which tests/test_oracle_cvmodel.py pins against cv2 over all 2^24 colours."""
    import re

    from oracle import cvmodel as M
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(ROOT, "video_unscreen_b200", "csrc", "vu_lab_tables.inc")).read().replace("\\\n", " ")
    tabs = {}
    for name, body in re.findall(r"#define (\w+)\s+([0-9,\s]+)", text):
        tabs[name] = np.array([int(v) for v in body.replace(" ", "").strip(",").split(",") if v])
    gamma, cbrt, coeffs = M.lab_tables()
    assert np.array_equal(tabs["VU_LAB_GAMMA_TABLE"], gamma) and np.array_equal(tabs["VU_LAB_CBRT_TABLE"], cbrt)
    src = open(os.path.join(ROOT, "video_unscreen_b200", "csrc", "vu_colorcorrect.cu")).read()
    m = re.search(r"#define VU_LAB_COEFFS \{([0-9,\s]+)\}", src)
    assert [int(v) for v in m.group(1).split(",")] == coeffs.reshape(-1).tolist()

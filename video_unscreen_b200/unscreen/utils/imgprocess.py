"""size math, plain resize and the geometric pre-steps of the replacement path
(reference: unscreen/utils/imgprocess.py)."""
from ... import ops
from ..._io import back, to_dev

__all__ = ["get_target_size", "adaptive_resize", "color_correct", "rescale_fg", "shift_fg"]


def get_target_size(h, w, target_long_side, division=1):
    """reference unscreen/utils/imgprocess.py:164-192 (host-side integer math)."""
    if h > w:
        target_h = target_long_side
        target_w = int(float(target_long_side) * w / h)
        if target_w % division != 0:
            target_w = (target_w // division + 1) * division
    else:
        target_w = target_long_side
        target_h = int(float(target_long_side) * h / w)
        if target_h % division != 0:
            target_h = (target_h // division + 1) * division
    return target_h, target_w


def adaptive_resize(img, img_target):
    """reference unscreen/utils/imgprocess.py:33-37: cv2.resize to the target's size."""
    t, as_np = to_dev(img)
    th, tw = img_target.shape[0], img_target.shape[1]
    out = ops.resize_linear_image(t, th, tw) if t.ndim == 3 else ops.resize_linear_mask(t, th, tw)
    return back(out, as_np)


def _channels(t):
    """the reference passes HWC images (fg, 3-channel JPEG masks) and HW maps through the same functions."""
    if t.ndim == 2:
        return 1
    if t.ndim == 3 and t.shape[2] == 3:
        return 3
    raise ValueError(f"expected an HW or HWx3 uint8 array, got {tuple(t.shape)}")


def rescale_fg(img, scale_factor=1.1):
    """reference unscreen/utils/imgprocess.py:40-52: bicubic up-scale about the centre, cropped to the input size.
    Parity with cv2's IPP cubic: equal except for < 1e-5 of the values, off by one."""
    t, as_np = to_dev(img)
    return back(ops.rescale_cubic(t, scale_factor, _channels(t)), as_np)


def shift_fg(img, dx=0, dy=0):
    """reference unscreen/utils/imgprocess.py:55-64: cv2.warpAffine translation, zero border (bit-exact)."""
    t, as_np = to_dev(img)
    return back(ops.shift(t, dx, dy, _channels(t)), as_np)


def color_correct(img, alpha, bg_color, target_long_side=960, mean_exp=0.95):
    """reference unscreen/utils/imgprocess.py:263-300 (green.py:120): alpha times the normalised Lab chroma distance
    to the background colour.  Bit-exact float32 sequence; the loop's mean is accumulated in float64 on the device."""
    t, as_np = to_dev(img)
    a, _ = to_dev(alpha)
    th, tw = get_target_size(t.shape[0], t.shape[1], target_long_side)
    col = bg_color.cpu().numpy() if hasattr(bg_color, "cpu") else bg_color
    return back(ops.color_correct(t, a, col, th, tw, mean_exp), as_np)

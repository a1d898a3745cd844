"""size math and plain resize (reference: unscreen/utils/imgprocess.py)."""
from ... import ops
from ..._io import back, to_dev

__all__ = ["get_target_size", "adaptive_resize"]


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

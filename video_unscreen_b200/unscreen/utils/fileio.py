"""Frame I/O on the device (reference: unscreen/utils/fileio.py:31-62).

The reference decodes a whole clip with ``Pool(48)`` x ``cv2.imread`` and writes three JPEGs per frame with
``cv2.imwrite``; once the kernels run at the rates of DESIGN.md section 4 every loop of tools/ is bound by exactly that.
Here the files are read on threads and decoded in batches by nvJPEG on the GPU (the codec is library code: torchvision's
bundled nvJPEG through ``torchvision.io``), the layout conversion to the reference's interleaved BGR is a kernel of this
library, and frames can stay on the device for the stages that follow.

JPEG is lossy and its decoders are not bit-identical: nvJPEG and cv2's libjpeg-turbo differ by a few LSB around edges (inverse
DCT and chroma up-sampling), so the contract here is a TOLERANCE (tests/test_gpu_io.py), not bit parity, and
``video_unscreen_b200.install()`` only swaps these in when asked (``install(io=True)``).  Not mirrored: EXIF orientation
(cv2.imread applies it; extracted video frames carry none) and ``save_video`` (mmcv / ffmpeg)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from ... import ops
from ..._io import device, to_dev

__all__ = ["parallel_read_img", "save_img"]

_JPEG = (".jpg", ".jpeg", ".jpe")


def _read_bytes(path):
    with open(path, "rb") as f:
        return torch.frombuffer(bytearray(f.read()), dtype=torch.uint8)


def parallel_read_img(framepaths, on_device=False, batch=64, threads=16):
    """reference fileio.py:31-39: the frames at ``framepaths`` as HxWx3 BGR uint8 arrays, in order.  JPEG files are decoded
    on the GPU, ``batch`` files per nvJPEG call; ``on_device=True`` returns CUDA tensors instead of numpy arrays.  Other
    formats go through cv2.imread like in the reference (PNG masks: lossless, identical)."""
    import torchvision.io as tio
    framepaths = list(framepaths)
    out = [None] * len(framepaths)
    dev = device()
    jpeg = [i for i, p in enumerate(framepaths) if os.path.splitext(p)[1].lower() in _JPEG]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for s in range(0, len(jpeg), batch):
            idx = jpeg[s:s + batch]
            datas = list(ex.map(_read_bytes, [framepaths[i] for i in idx]))
            planes = tio.decode_jpeg(datas, device=dev, mode=tio.ImageReadMode.RGB)
            for i, pl in zip(idx, planes):
                bgr = ops.planar_rgb_to_bgr(pl.contiguous())
                out[i] = bgr if on_device else bgr.cpu().numpy()
    rest = [i for i in range(len(framepaths)) if out[i] is None]
    if rest:
        import cv2
        for i in rest:
            img = cv2.imread(framepaths[i])
            out[i] = torch.from_numpy(img).to(dev) if (on_device and img is not None) else img
    return out


def save_img(img, save_path, downsacle=1, quality=95):
    """reference fileio.py:51-62: down-sample by an integer factor (cv2.resize) and write.  3-channel ``.jpg`` images are encoded by
    nvJPEG on the GPU at cv2.imwrite's default quality (95); anything else is written by cv2.imwrite."""
    assert isinstance(downsacle, int)
    t, _ = to_dev(img)
    if downsacle != 1:
        h, w = t.shape[:2]
        t = ops.resize_linear(t, h // downsacle, w // downsacle)
    if t.ndim == 3 and os.path.splitext(save_path)[1].lower() in _JPEG:     # single-channel images: cv2 (nvJPEG encodes 3 planes)
        import torchvision.io as tio
        data = tio.encode_jpeg(ops.bgr_to_planar_rgb(t), quality=quality)
        with open(save_path, "wb") as f:
            f.write(data.cpu().numpy().tobytes())
        return
    import cv2
    cv2.imwrite(save_path, t.cpu().numpy())

"""numpy <-> device plumbing for the reference-shaped API.

The reference's functions take and return numpy uint8 arrays; the B200 path
also accepts CUDA tensors (zero-copy, stays on the device).  Whatever kind
came in goes out.
"""
import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("video_unscreen_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(x):
    """-> (contiguous uint8 CUDA tensor, came_from_numpy)"""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.to(device())
        if x.dtype == torch.bool:
            x = x.to(torch.uint8)
        return x.contiguous(), False
    a = np.ascontiguousarray(x)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype != np.uint8:
        raise TypeError(f"expected uint8 data, got {a.dtype}")
    return torch.from_numpy(a).to(device(), non_blocking=False), True


def back(t, as_numpy, dtype=None):
    if as_numpy:
        a = t.cpu().numpy()
        return a.astype(dtype) if dtype is not None else a
    return t.to(torch.bool) if dtype is np.bool_ else t

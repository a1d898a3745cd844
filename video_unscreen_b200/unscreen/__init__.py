"""Drop-in mirror of the reference's ``unscreen`` package for the per-pixel
matte hot path: same import paths, names, signatures and conventions (numpy
uint8, HWC, BGR), computed by the sm_100a kernels of libvu_b200.so.

    from video_unscreen_b200.unscreen.colorfiltering import ColorFilteringAgent
    from video_unscreen_b200.unscreen.trimap import TrimapAgent
    from video_unscreen_b200.unscreen.bgmodel import BackgroundAgent
    from video_unscreen_b200.unscreen.utils import get_fg, dilate_mask, ...

``video_unscreen_b200.install()`` aliases it as ``unscreen`` so the reference's
tools/unscreen scripts pick it up unchanged.
"""

"""BackgroundAgent: signature kept, body out of scope.

The reference's agent (unscreen/bgmodel/agent.py:9-208) is single-image
spatial inpainting (boundary mean colour / iterated box filters / a sparse
Laplace solve with scipy).  It has no caller in tools/, is not a streaming
per-pixel kernel, and SURVEY.md section 8 (row a25) keeps only its signature
importable.  The temporal background estimators of the bg_step path live in
``unscreen.utils.temporal`` (temporal_median, masked_temporal_mean)."""


class BackgroundAgent():

    def __init__(self, input_long_side=540, dilation_ksize=5, dilation_iters=3, boundary_ksize=7, boundary_iters=10,
                 pcov_ksize=5):
        self.input_long_side = input_long_side
        self.dilation_ksize = dilation_ksize
        self.dilation_iters = dilation_iters
        self.boundary_ksize = boundary_ksize
        self.boundary_iters = boundary_iters
        self.pcov_ksize = pcov_ksize

    def forward(self, img, mask, method='rf'):
        if method not in ('mean', 'pcov', 'rf'):
            raise NameError(f'No such method for background inpainting: {method}')
        raise NotImplementedError(
            "BackgroundAgent.forward (single-image inpainting, reference bgmodel/agent.py:159-208) is outside the "
            "B200 hot path (SURVEY.md section 8 a25); use the reference's own implementation for it")

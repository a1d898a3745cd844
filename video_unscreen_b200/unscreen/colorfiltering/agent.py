"""ColorFilteringAgent on the B200 (reference: unscreen/colorfiltering/agent.py).

Per-pixel evaluation (BGR2HSV, down-scale, mixture evaluation, adaptive
threshold, d2/e2/e2/d2 morphology, up-scale) runs on the GPU.  The EM fit of
the six 1-D mixtures stays scikit-learn on the host, exactly as in the
reference (a tiny, order-sensitive, RNG-seeded statistical fit on <= 20k
samples): the GPU gathers the row-major strided samples the reference gathers
(vu_cf_samples) and hands scikit-learn those and the hue histogram.
"""
import cv2
import numpy as np
import torch
from sklearn import mixture

from ... import _lib, ops
from ..._io import back, to_dev
from ..utils.imgprocess import get_target_size


def gmm_parameters(gmm, use_opencv_gmm=False):
    if use_opencv_gmm:
        means = gmm.getMeans().squeeze()
        stds = np.sqrt(np.array(gmm.getCovs()).squeeze())
        weights = gmm.getWeights().squeeze()
    else:
        means = gmm.means_.squeeze()
        stds = np.sqrt(gmm.covariances_.squeeze())
        weights = gmm.weights_.squeeze()
    return means, stds, weights


def gmm_table(means, stds, weights):
    """the reference's get_prob_by_gmm (agent.py:201-230) evaluated on the 256
    possible uint8 samples with the same torch CPU ops in the same order: the
    per-pixel values of the reference are exactly entries of this table."""
    samples = np.arange(256, dtype=np.float64).reshape(1, -1)
    means = np.atleast_1d(np.asarray(means, np.float64))
    stds = np.atleast_1d(np.asarray(stds, np.float64))
    weights = np.atleast_1d(np.asarray(weights, np.float64))
    samples_t = torch.from_numpy(samples).float()
    means_t = torch.from_numpy(means[..., np.newaxis]).float()
    stds_t = torch.from_numpy(stds[..., np.newaxis]).float()
    weights_t = torch.from_numpy(weights).float().unsqueeze(dim=0)
    x = (samples_t - means_t) / stds_t
    y = 1. / (stds_t * np.sqrt(2 * np.pi)) * torch.exp(-1. / 2 * torch.pow(x, 2))
    return torch.mm(weights_t, y).reshape(256)


class ColorFilteringAgent():
    """same constructor, attributes and methods as the reference's agent
    (agent.py:49-70); ``forward`` accepts numpy arrays or CUDA tensors."""

    def __init__(self, input_long_side=960, bg_ncomp=(3, 5, 5), fg_ncomp=(10, 10, 10), max_num_samples=10000,
                 color_prior_winsize=30, use_opencv_gmm=False):
        assert isinstance(input_long_side, int)
        self.input_long_side = input_long_side
        assert len(bg_ncomp) == 3
        assert len(fg_ncomp) == 3
        self.bg_ncomp = bg_ncomp
        self.fg_ncomp = fg_ncomp
        assert isinstance(max_num_samples, int)
        assert max_num_samples > 2
        self.max_num_samples = max_num_samples
        assert isinstance(color_prior_winsize, int)
        assert color_prior_winsize > 0
        self.color_prior_winsize = color_prior_winsize
        self.use_opencv_gmm = use_opencv_gmm
        self.reset_gmms()

    # ---- model state (host) ------------------------------------------------
    def is_trained(self):
        return self._is_trained

    def reset_gmms(self):
        """agent.py:81-111."""
        self.bg_gmms, self.fg_gmms = [], []
        for i in range(3):
            if self.use_opencv_gmm:
                for lst, k in ((self.bg_gmms, self.bg_ncomp[i]), (self.fg_gmms, self.fg_ncomp[i])):
                    m = cv2.ml.EM_create()
                    m.setClustersNumber(k)
                    m.setCovarianceMatrixType(cv2.ml.EM_COV_MAT_SPHERICAL)
                    lst.append(m)
            else:
                self.bg_gmms.append(mixture.GaussianMixture(n_components=self.bg_ncomp[i], covariance_type='spherical', warm_start=True))
                self.fg_gmms.append(mixture.GaussianMixture(n_components=self.fg_ncomp[i], covariance_type='spherical', warm_start=True))
        self._is_trained = False
        self._luts_dev = None

    def set_tables(self, lut_bg, lut_fg, bg_hsv):
        """install already-evaluated mixture tables ((3,256) float32 each) and
        the component-0 background colour; used to run many agents / ranks from
        one fit without shipping sklearn objects around."""
        luts = np.concatenate([np.asarray(lut_bg, np.float32).reshape(3, 256), np.asarray(lut_fg, np.float32).reshape(3, 256)])
        self._luts_dev = torch.from_numpy(np.ascontiguousarray(luts)).cuda()
        self._bg_hsv = np.asarray(bg_hsv, np.uint8)
        self._is_trained = True

    def tables(self):
        """(lut_bg (3,256), lut_fg (3,256), bg_hsv (3,)) of the current mixtures."""
        lb = torch.stack([gmm_table(*gmm_parameters(g, self.use_opencv_gmm)) for g in self.bg_gmms])
        lf = torch.stack([gmm_table(*gmm_parameters(g, self.use_opencv_gmm)) for g in self.fg_gmms])
        hsv = []
        for i in range(3):  # agent.py:345-351: component 0 only
            m = self.bg_gmms[i].getMeans().squeeze() if self.use_opencv_gmm else self.bg_gmms[i].means_[0, 0]
            hsv.append(int(np.mean(m)))
        return lb.numpy(), lf.numpy(), np.array(hsv, np.uint8)

    def _refresh_tables(self):
        self.set_tables(*self.tables())

    def tables_dev(self):
        """[6,256] float32 device tensor: bg H,S,V then fg H,S,V mixture tables."""
        if self._luts_dev is None:
            self._refresh_tables()
        return self._luts_dev

    def lut3d_dev(self):
        """alpha tabulated over every (h<180, s, v) for the current mixtures
        (built once per fit on the device, 11.8 MB: L2-resident)."""
        self.tables_dev()
        if getattr(self, "_lut3d_for", None) is not self._luts_dev:
            self._lut3d = ops.cf_build_lut3d(self._luts_dev)
            self._lut3d_for = self._luts_dev
        return self._lut3d

    def bg_color_bgr(self):
        """the constant background colour of forward()'s bg_img as a (3,) uint8 BGR array"""
        self.tables_dev()
        key = bytes(np.asarray(self._bg_hsv, np.uint8))
        if getattr(self, "_bg_bgr_key", None) != key:     # one conversion per fit, not per frame
            px = torch.from_numpy(np.tile(self._bg_hsv, (1, 4, 1))).cuda()
            self._bg_bgr, self._bg_bgr_key = ops.hsv2bgr(px)[0, 0].cpu().numpy(), key
        return self._bg_bgr.copy()

    def _sample(self, channel, mask):
        samples = channel[mask].astype(float)
        if len(samples) > self.max_num_samples:
            samples = samples[::len(samples) // self.max_num_samples]
        return samples

    def get_color_prior(self, img_hsv, mask, color_prior_winsize=None):
        """agent.py:113-146 (host arrays)."""
        if color_prior_winsize is None:
            color_prior_winsize = self.color_prior_winsize
        samples = self._sample(img_hsv[:, :, 0], mask)
        hist, _ = np.histogram(samples, 256, [0, 256])
        peak = np.argmax(hist)
        h = img_hsv[:, :, 0]
        return (h > peak - color_prior_winsize // 2) & (h < peak + color_prior_winsize // 2)

    def _fit(self, gmms, img_hsv, mask):
        for i in range(3):
            samples = self._sample(img_hsv[:, :, i], mask)
            if self.use_opencv_gmm:
                gmms[i].trainEM(samples[..., np.newaxis])
            else:
                gmms[i].fit(samples[..., np.newaxis])
        self._is_trained = True
        self._luts_dev = None

    def fit_bg_gmms(self, img_hsv, mask, mask_by_prior=None):
        """agent.py:148-172."""
        if mask_by_prior is None:
            mask_by_prior = self.get_color_prior(img_hsv, mask)
        self._fit(self.bg_gmms, img_hsv, mask & mask_by_prior)

    def fit_fg_gmms(self, img_hsv, mask, mask_by_prior=None):
        """agent.py:174-199."""
        if mask_by_prior is None:
            mask_by_prior = self.get_color_prior(img_hsv, (1 - mask), self.color_prior_winsize // 5)
        if (mask & (1 - mask_by_prior)).sum() > max(self.fg_ncomp) * 5:
            mask = (mask & (1 - mask_by_prior).astype(bool))
        self._fit(self.fg_gmms, img_hsv, mask)

    def _fit_dev(self, hsv_lo, mask_lo):
        """one fit iteration of forward (agent.py:325-332) with the samples gathered on the device (vu_cf_samples:
        same pixels, same order, same stride as channel[mask][::step]); only the <= 20k samples per channel, the
        selection count and the 256-bin hue histogram come back for scikit-learn."""
        w_bg, w_fg = self.color_prior_winsize, self.color_prior_winsize // 5
        # get_color_prior(hsv, mask < 128, w) for both windows: same samples, same histogram, same peak
        _, _, hist = ops.cf_samples(hsv_lo, mask_lo, 0, self.max_num_samples)
        peak = int(np.argmax(hist))
        bg_prior = (peak - w_bg // 2, peak + w_bg // 2)
        fg_prior = (peak - w_fg // 2, peak + w_fg // 2)
        # fit_bg_gmms(hsv, mask < 128, bg_prior)
        samples, _, _ = ops.cf_samples(hsv_lo, mask_lo, 0, self.max_num_samples, prior=bg_prior)
        self._fit_samples(self.bg_gmms, samples)
        # fit_fg_gmms(hsv, mask > 128, fg_prior): outside the prior if that leaves enough pixels
        samples, total, _ = ops.cf_samples(hsv_lo, mask_lo, 1, self.max_num_samples, prior=fg_prior, invert=True)
        if not total > max(self.fg_ncomp) * 5:
            samples, _, _ = ops.cf_samples(hsv_lo, mask_lo, 1, self.max_num_samples)
        self._fit_samples(self.fg_gmms, samples)

    def _fit_samples(self, gmms, samples):
        for i in range(3):
            x = samples[i].astype(float)[..., np.newaxis]
            if self.use_opencv_gmm:
                gmms[i].trainEM(x)
            else:
                gmms[i].fit(x)
        self._is_trained = True
        self._luts_dev = None

    # ---- per-pixel evaluation (device) ---------------------------------------
    def _alpha_dev(self, hsv_dev):
        if self._luts_dev is None:
            self._refresh_tables()
        return ops.cf_alpha(hsv_dev, self._luts_dev)

    @staticmethod
    def _postprocess_dev(alpha_dev, mask_dev, thr_ratio=0.8):
        return ops.cf_postprocess(alpha_dev, mask_dev, thr_ratio)

    def get_alpha_by_gmm(self, img_hsv):
        """agent.py:232-257 -> (alpha, confidence).  The reference returns the
        bound method ``torch.std(prob).item`` as 'confidence' (never called);
        None is returned here."""
        t, as_np = to_dev(img_hsv)
        return back(self._alpha_dev(t), as_np), None

    def postprocess(self, alpha, mask, thr_ratio=0.8):
        """agent.py:259-283 (the input alpha is not mutated)."""
        a, as_np = to_dev(alpha)
        m, _ = to_dev(mask)
        return back(self._postprocess_dev(a, m, thr_ratio), as_np)

    def _degenerate(self, mask_dev):
        nfg = int(ops.count_cmp(mask_dev, _lib.CMP_GT, 128).item())
        nbg = int(ops.count_cmp(mask_dev, _lib.CMP_LT, 128).item())
        return nfg < max(self.fg_ncomp) * 5, nbg < max(self.bg_ncomp) * 5

    def forward(self, img, mask, iters=1):
        """agent.py:285-354 -> (alpha HxW, bg_img HxWx3, confidence)."""
        img_t, as_np = to_dev(img)
        mask_t, _ = to_dev(mask)
        if iters == 0 and (self._luts_dev is not None or self._is_trained):
            # predict only (the per-frame call of green.py:99 between refits): the batched path with the early-outs decided
            # on the device - no host round trip before the kernels, one read-back (flag + matte) after them
            from ... import clip
            alpha_d, flags = clip.cf_predict_clip(img_t[None], mask_t[None], self, chunk=1, streams=1, return_flags=True)
            flag = int(flags.cpu()[0])
            if flag == 1:
                return mask, img, 1.0
            if flag == 2:
                return mask, (np.zeros_like(img) if as_np else torch.zeros_like(img_t)), 1.0
            col = self.bg_color_bgr()
            if as_np:
                # a fresh constant image per call (callers write into it, green.py:125): a copy of a cached one (a numpy
                # broadcast fill of 3-byte pixels takes longer than all the kernels together)
                key = (tuple(img_t.shape), bytes(col))
                if getattr(self, "_bg_img_key", None) != key:
                    self._bg_img, self._bg_img_key = np.ascontiguousarray(np.broadcast_to(col, tuple(img_t.shape))), key
                return alpha_d[0].cpu().numpy(), self._bg_img.copy(), None
            return alpha_d[0], torch.from_numpy(col).to(img_t.device).expand(tuple(img_t.shape)).contiguous(), None
        no_fg, no_bg = self._degenerate(mask_t)
        if no_fg:
            return mask, img, 1.0
        if no_bg:
            return mask, (np.zeros_like(img) if as_np else torch.zeros_like(img_t)), 1.0

        hsv = ops.bgr2hsv(img_t)
        ori_h, ori_w = hsv.shape[:2]
        target_h, target_w = get_target_size(ori_h, ori_w, self.input_long_side)
        hsv_lo = ops.resize_linear_image(hsv, target_h, target_w)
        mask_lo = ops.resize_linear_mask(mask_t, target_h, target_w)

        if iters == 0:
            alpha = self._postprocess_dev(self._alpha_dev(hsv_lo), mask_lo)
        else:
            for _ in range(iters):
                self._fit_dev(hsv_lo, mask_lo)
                alpha = self._postprocess_dev(self._alpha_dev(hsv_lo), mask_lo)
                mask_lo = ops.binarise(alpha, 128)
                no_fg, no_bg = self._degenerate(mask_lo)
                if no_fg or no_bg:
                    break
        alpha = ops.resize_linear_mask(alpha, ori_h, ori_w)

        if self._luts_dev is None:
            self._refresh_tables()
        px = torch.from_numpy(np.tile(self._bg_hsv, (1, 4, 1))).cuda()
        bgr = ops.hsv2bgr(px)[0, 0]
        bg_img = bgr.expand(ori_h, ori_w, 3).contiguous()
        return back(alpha, as_np), back(bg_img, as_np), None

"""PixelwiseContrastiveLoss -- same class, method names, argument order and return tuples as the reference's
dense_correspondence/loss_functions/pixelwise_contrastive_loss.py, computed by the fused CUDA loss kernels.

Reference lines each method replaces are cited per method.  Like the reference, the methods that return
``num_hard_negatives`` return a Python int (which costs a device->host sync, exactly as
``len(torch.nonzero(...))`` does at pixelwise_contrastive_loss.py:210-211); the training path that matters,
``loss_composer.get_loss``, uses the fused sync-free kernel sequence instead.

Descriptor images: ``[1, W*H, D]`` (or ``[B, W*H, D]`` with ``[B, n]`` index tensors) float32 CUDA tensors,
normally the strided views produced by ``DenseCorrespondenceNetwork.process_network_output``.
"""
import torch

from . import _native as N
from .contrastive_ops import Term, contrastive_terms


# the reference's default loss configuration (config/dense_correspondence/training/training.yaml:51-61)
DEFAULT_LOSS_CONFIG = {
    "M_masked": 0.5, "M_background": 0.5, "M_pixel": 50,
    "match_loss_weight": 1.0, "non_match_loss_weight": 1.0,
    "use_l2_pixel_loss_on_masked_non_matches": False,
    "use_l2_pixel_loss_on_background_non_matches": False,
    "scale_by_hard_negatives": True,
    "scale_by_hard_negatives_DIFFERENT_OBJECT": True,
    "alpha_triplet": 0.1,
}


class PixelwiseContrastiveLoss(object):

    def __init__(self, image_shape, config=None):
        self.type = "pixelwise_contrastive"
        self.image_width = image_shape[1]
        self.image_height = image_shape[0]
        assert config is not None
        self._config = config
        self._debug_data = dict()
        self._debug = False

    @property
    def debug(self):
        return self._debug

    @debug.setter
    def debug(self, value):
        self._debug = value

    @property
    def config(self):
        return self._config

    @property
    def debug_data(self):
        return self._debug_data

    # ------------------------------------------------------------------ helpers
    def _terms(self, image_a_pred, image_b_pred, terms):
        sums, counts = contrastive_terms(image_a_pred, image_b_pred, self.image_width, terms)
        # reference semantics are per single pair; for B > 1 report the mean over pairs
        return sums.mean(0).to(torch.float32), counts

    # ------------------------------------------------------------------ reference API
    def get_loss_matched_and_non_matched_with_l2(self, image_a_pred, image_b_pred, matches_a, matches_b,
                                                 non_matches_a, non_matches_b, M_descriptor=None, M_pixel=None,
                                                 non_match_loss_weight=1.0, use_l2_pixel_loss=None):
        """pixelwise_contrastive_loss.py:35-101 -> (match_loss, non_match_loss_sum, num_hard_negatives)."""
        if M_descriptor is None:
            M_descriptor = self._config["M_descriptor"]
        if M_pixel is None:
            M_pixel = self._config["M_pixel"]
        if use_l2_pixel_loss is None:
            use_l2_pixel_loss = self._config['use_l2_pixel_loss_on_masked_non_matches']
        hinge = Term(non_matches_a, non_matches_b, N.TERM_HINGE, M_descriptor,
                     gt_b=matches_b if use_l2_pixel_loss else None, m_pixel=M_pixel)
        sums, counts = self._terms(image_a_pred, image_b_pred,
                                   [Term(matches_a, matches_b, N.TERM_MATCH), hinge])
        match_loss = sums[0] * (1.0 / matches_a.shape[-1])
        return match_loss, sums[1], int(counts[:, 1].sum().item())

    @staticmethod
    def get_triplet_loss(image_a_pred, image_b_pred, matches_a, matches_b, non_matches_a, non_matches_b, alpha):
        """pixelwise_contrastive_loss.py:103-129.  Not reachable from loss_composer.get_loss;
        composed from torch ops on the device -- no dedicated kernel."""
        num_matches = matches_a.size()[0]
        num_non_matches = non_matches_a.size()[0]
        multiplier = num_non_matches // num_matches
        matches_b_long = torch.t(matches_b.repeat(multiplier, 1)).contiguous().view(-1)
        a = torch.index_select(image_a_pred, 1, non_matches_a)
        b = torch.index_select(image_b_pred, 1, matches_b_long)
        nb = torch.index_select(image_b_pred, 1, non_matches_b)
        triplet_losses = (a - b).pow(2) - (a - nb).pow(2) + alpha
        return 1.0 / num_non_matches * torch.clamp(triplet_losses, min=0).sum()

    @staticmethod
    def match_loss(image_a_pred, image_b_pred, matches_a, matches_b):
        """pixelwise_contrastive_loss.py:131-167 -> (match_loss, matches_a_descriptors, matches_b_descriptors)."""
        sums, _ = contrastive_terms(image_a_pred, image_b_pred, 1, [Term(matches_a, matches_b, N.TERM_MATCH)])
        match_loss = (sums.mean(0)[0] * (1.0 / matches_a.shape[-1])).to(torch.float32)
        # the gathered descriptors are returned for API parity only (nothing on the hot path reads them)
        a = torch.index_select(image_a_pred, 1, matches_a) if matches_a.dim() == 1 else None
        b = torch.index_select(image_b_pred, 1, matches_b) if matches_b.dim() == 1 else None
        return match_loss, a, b

    @staticmethod
    def non_match_descriptor_loss(image_a_pred, image_b_pred, non_matches_a, non_matches_b, M=0.5, invert=False):
        """pixelwise_contrastive_loss.py:170-213 -> (loss_vector [n], num_hard_negatives, a_desc, b_desc).
        The per-pair VECTOR is only needed by callers outside the hot path, so it is composed from torch ops;
        the fused kernels never materialise it."""
        a = torch.index_select(image_a_pred, 1, non_matches_a).squeeze()
        b = torch.index_select(image_b_pred, 1, non_matches_b).squeeze()
        if len(non_matches_a) == 1:
            a = a.unsqueeze(0)
            b = b.unsqueeze(0)
        d = (a - b).norm(2, 1)
        loss_vec = torch.clamp(M - d, min=0).pow(2) if not invert else torch.clamp(d - M, min=0).pow(2)
        return loss_vec, len(torch.nonzero(loss_vec)), a, b

    def non_match_loss_with_l2_pixel_norm(self, image_a_pred, image_b_pred, matches_b, non_matches_a, non_matches_b,
                                          M_descriptor=0.5, M_pixel=None):
        """pixelwise_contrastive_loss.py:215-269 -> (sum_j l_j * w_j, num_hard_negatives)."""
        if M_descriptor is None:
            M_descriptor = self._config["M_descriptor"]
        if M_pixel is None:
            M_pixel = self._config["M_pixel"]
        sums, counts = self._terms(image_a_pred, image_b_pred,
                                   [Term(non_matches_a, non_matches_b, N.TERM_HINGE, M_descriptor, gt_b=matches_b,
                                         m_pixel=M_pixel)])
        num_hard_negatives = int(counts.sum().item())
        if self.debug:
            self._debug_data['num_hard_negatives'] = num_hard_negatives
            self._debug_data['fraction_hard_negatives'] = num_hard_negatives * 1.0 / non_matches_a.numel()
        return sums[0], num_hard_negatives

    def non_match_loss_descriptor_only(self, image_a_pred, image_b_pred, non_matches_a, non_matches_b,
                                       M_descriptor=0.5, invert=False):
        """pixelwise_contrastive_loss.py:271-304 -> (sum_j l_j, num_hard_negatives)."""
        if M_descriptor is None:
            M_descriptor = self._config["M_descriptor"]
        kind = N.TERM_HINGE_INV if invert else N.TERM_HINGE
        sums, counts = self._terms(image_a_pred, image_b_pred, [Term(non_matches_a, non_matches_b, kind, M_descriptor)])
        num_hard_negatives = int(counts.sum().item())
        if self._debug:
            self._debug_data['num_hard_negatives'] = num_hard_negatives
            self._debug_data['fraction_hard_negatives'] = num_hard_negatives * 1.0 / non_matches_a.numel()
        return sums[0], num_hard_negatives

    def l2_pixel_loss(self, matches_b, non_matches_b, M_pixel=None):
        """pixelwise_contrastive_loss.py:307-334 (the returned weight is NOT squared, despite its upstream name)."""
        if M_pixel is None:
            M_pixel = self._config['M_pixel']
        k = len(non_matches_b) // len(matches_b)
        gt = torch.t(matches_b.repeat(k, 1)).contiguous().view(-1, 1)
        gt_uv = self.flattened_pixel_locations_to_u_v(gt)
        s_uv = self.flattened_pixel_locations_to_u_v(non_matches_b.unsqueeze(1))
        w = 1.0 / M_pixel * torch.clamp((gt_uv - s_uv).float().norm(2, 1), max=M_pixel)
        return w, gt_uv, s_uv

    def flattened_pixel_locations_to_u_v(self, flat_pixel_locations):
        """pixelwise_contrastive_loss.py:338-352: n -> (n % W, n // W)."""
        uv = flat_pixel_locations.repeat(1, 2)
        uv[:, 0] = uv[:, 0] % self.image_width
        uv[:, 1] = uv[:, 1] // self.image_width
        return uv

    def get_l2_pixel_loss_original(self):
        pass

    def get_loss_original(self, image_a_pred, image_b_pred, matches_a, matches_b, non_matches_a, non_matches_b,
                          M_margin=0.5, non_match_loss_weight=1.0):
        """pixelwise_contrastive_loss.py:357-411 (legacy hinge on the SQUARED distance; unused by training.py)."""
        num_matches = matches_a.size()[0]
        num_non_matches = non_matches_a.size()[0]
        sums, _ = self._terms(image_a_pred, image_b_pred, [Term(matches_a, matches_b, N.TERM_MATCH)])
        match_loss = sums[0] * (1.0 / num_matches)
        na = torch.index_select(image_a_pred, 1, non_matches_a)
        nb = torch.index_select(image_b_pred, 1, non_matches_b)
        pw = torch.add(torch.neg((na - nb).pow(2).sum(dim=2)), M_margin)
        non_match_loss = non_match_loss_weight * 1.0 / num_non_matches * torch.clamp(pw, min=0).sum()
        return match_loss + non_match_loss, match_loss, non_match_loss

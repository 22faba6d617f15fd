"""CPU oracle for PixelwiseContrastiveLoss + loss_composer -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this file; the product never does.

Two independent restatements of the same reference source (which is Python-2 and
cannot be imported as it lies: TabError at pixelwise_contrastive_loss.py:8, print
statements in loss_composer.py:28):

* ``Torch*``  -- line-by-line Python-3 torch restatement (autograd provides the
  reference gradient).  ``/`` on ints is ``//`` (py2 semantics) wherever the
  reference divides integers.
* ``np_*``    -- numpy float64 restatement used to pin the torch one.

Reference lines followed (dense_correspondence/loss_functions/):
  pixelwise_contrastive_loss.py:131-167  match_loss
  pixelwise_contrastive_loss.py:170-213  non_match_descriptor_loss (+ nonzero count)
  pixelwise_contrastive_loss.py:215-269  non_match_loss_with_l2_pixel_norm
  pixelwise_contrastive_loss.py:271-304  non_match_loss_descriptor_only
  pixelwise_contrastive_loss.py:307-334  l2_pixel_loss   (value is NOT squared)
  pixelwise_contrastive_loss.py:338-352  flattened_pixel_locations_to_u_v
  pixelwise_contrastive_loss.py:35-101   get_loss_matched_and_non_matched_with_l2
  pixelwise_contrastive_loss.py:103-129  get_triplet_loss
  pixelwise_contrastive_loss.py:357-411  get_loss_original
  loss_composer.py:7-67                  get_loss dispatch
  loss_composer.py:70-143                get_within_scene_loss
  loss_composer.py:168-191               get_different_object_loss
  loss_composer.py:193-212               get_same_object_across_scene_loss (latent NameError kept)
  dataset/dense_correspondence_dataset_masked.py:209-223  empty_tensor / is_empty sentinel
  dataset/spartan_dataset_masked.py:31-36                 SpartanDatasetDataType

Parity status: PINNED to the executed reference.  The reference ships no golden vectors
for this path, so oracle/build_ref.py makes the reference's own files
importable (line-anchored py2->py3 patches, written to the git-ignored oracle/_ref/) and
tests/test_oracle_ref_cpu.py requires the torch restatement below to be BIT-EQUAL to the
executed reference (values, hard-negative counts, autograd gradients, every public method
and composer branch, the non-match sampler and the reprojection match finder); the numpy
restatement agrees to 1e-6; tests/golden/loss_*.npz hold the executed reference's outputs
(oracle/make_golden.py).
"""
import numpy as np
import torch


class SpartanDatasetDataType:
    SINGLE_OBJECT_WITHIN_SCENE = 0
    SINGLE_OBJECT_ACROSS_SCENE = 1
    DIFFERENT_OBJECT = 2
    MULTI_OBJECT = 3
    SYNTHETIC_MULTI_OBJECT = 4


def empty_tensor():
    return torch.LongTensor([-1])


def is_empty(tensor):
    return (len(tensor) == 1) and bool(tensor[0] == -1)


DEFAULT_LOSS_CONFIG = {  # config/dense_correspondence/training/training.yaml:51-61
    "M_masked": 0.5, "M_background": 0.5, "M_pixel": 50,
    "match_loss_weight": 1.0, "non_match_loss_weight": 1.0,
    "use_l2_pixel_loss_on_masked_non_matches": False,
    "use_l2_pixel_loss_on_background_non_matches": False,
    "scale_by_hard_negatives": True,
    "scale_by_hard_negatives_DIFFERENT_OBJECT": True,
    "alpha_triplet": 0.1,
}


class TorchPixelwiseContrastiveLoss(object):
    def __init__(self, image_shape, config=None):
        self.type = "pixelwise_contrastive"
        self.image_width = image_shape[1]
        self.image_height = image_shape[0]
        assert config is not None
        self._config = config
        self._debug_data = dict()
        self._debug = False

    @property
    def config(self):
        return self._config

    def get_loss_matched_and_non_matched_with_l2(self, image_a_pred, image_b_pred, matches_a, matches_b,
                                                 non_matches_a, non_matches_b, M_descriptor=None, M_pixel=None,
                                                 non_match_loss_weight=1.0, use_l2_pixel_loss=None):
        if M_descriptor is None:
            M_descriptor = self._config["M_descriptor"]
        if M_pixel is None:
            M_pixel = self._config["M_pixel"]
        if use_l2_pixel_loss is None:
            use_l2_pixel_loss = self._config["use_l2_pixel_loss_on_masked_non_matches"]
        match_loss, _, _ = self.match_loss(image_a_pred, image_b_pred, matches_a, matches_b)
        if use_l2_pixel_loss:
            non_match_loss, num_hard_negatives = self.non_match_loss_with_l2_pixel_norm(
                image_a_pred, image_b_pred, matches_b, non_matches_a, non_matches_b,
                M_descriptor=M_descriptor, M_pixel=M_pixel)
        else:
            non_match_loss, num_hard_negatives = self.non_match_loss_descriptor_only(
                image_a_pred, image_b_pred, non_matches_a, non_matches_b, M_descriptor=M_descriptor)
        return match_loss, non_match_loss, num_hard_negatives

    @staticmethod
    def get_triplet_loss(image_a_pred, image_b_pred, matches_a, matches_b, non_matches_a, non_matches_b, alpha):
        num_matches = matches_a.size()[0]
        num_non_matches = non_matches_a.size()[0]
        multiplier = num_non_matches // num_matches
        matches_b_long = torch.t(matches_b.repeat(multiplier, 1)).contiguous().view(-1)
        matches_a_descriptors = torch.index_select(image_a_pred, 1, non_matches_a)
        matches_b_descriptors = torch.index_select(image_b_pred, 1, matches_b_long)
        non_matches_b_descriptors = torch.index_select(image_b_pred, 1, non_matches_b)
        triplet_losses = (matches_a_descriptors - matches_b_descriptors).pow(2) - \
            (matches_a_descriptors - non_matches_b_descriptors).pow(2) + alpha
        return 1.0 / num_non_matches * torch.clamp(triplet_losses, min=0).sum()

    @staticmethod
    def match_loss(image_a_pred, image_b_pred, matches_a, matches_b):
        num_matches = matches_a.size()[0]
        a = torch.index_select(image_a_pred, 1, matches_a)
        b = torch.index_select(image_b_pred, 1, matches_b)
        if len(matches_a) == 1:
            a = a.unsqueeze(0)
            b = b.unsqueeze(0)
        match_loss = 1.0 / num_matches * (a - b).pow(2).sum()
        return match_loss, a, b

    @staticmethod
    def non_match_descriptor_loss(image_a_pred, image_b_pred, non_matches_a, non_matches_b, M=0.5, invert=False):
        a = torch.index_select(image_a_pred, 1, non_matches_a).squeeze()
        b = torch.index_select(image_b_pred, 1, non_matches_b).squeeze()
        if len(non_matches_a) == 1:
            a = a.unsqueeze(0)
            b = b.unsqueeze(0)
        if a.dim() == 1:       # D == 1 collapses under squeeze(); keep [n, D]
            a = a.view(len(non_matches_a), -1)
            b = b.view(len(non_matches_b), -1)
        d = (a - b).norm(2, 1)
        if not invert:
            loss_vec = torch.clamp(M - d, min=0).pow(2)
        else:
            loss_vec = torch.clamp(d - M, min=0).pow(2)
        num_hard_negatives = len(torch.nonzero(loss_vec))
        return loss_vec, num_hard_negatives, a, b

    def non_match_loss_with_l2_pixel_norm(self, image_a_pred, image_b_pred, matches_b, non_matches_a, non_matches_b,
                                          M_descriptor=0.5, M_pixel=None):
        if M_descriptor is None:
            M_descriptor = self._config["M_descriptor"]
        if M_pixel is None:
            M_pixel = self._config["M_pixel"]
        loss_vec, num_hard_negatives, _, _ = self.non_match_descriptor_loss(
            image_a_pred, image_b_pred, non_matches_a, non_matches_b, M=M_descriptor)
        pix, _, _ = self.l2_pixel_loss(matches_b, non_matches_b, M_pixel=M_pixel)
        return (loss_vec * pix).sum(), num_hard_negatives

    def non_match_loss_descriptor_only(self, image_a_pred, image_b_pred, non_matches_a, non_matches_b,
                                       M_descriptor=0.5, invert=False):
        if M_descriptor is None:
            M_descriptor = self._config["M_descriptor"]
        loss_vec, num_hard_negatives, _, _ = self.non_match_descriptor_loss(
            image_a_pred, image_b_pred, non_matches_a, non_matches_b, M=M_descriptor, invert=invert)
        return loss_vec.sum(), num_hard_negatives

    def l2_pixel_loss(self, matches_b, non_matches_b, M_pixel=None):
        if M_pixel is None:
            M_pixel = self._config["M_pixel"]
        k = len(non_matches_b) // len(matches_b)            # py2 int division
        gt = torch.t(matches_b.repeat(k, 1)).contiguous().view(-1, 1)
        gt_uv = self.flattened_pixel_locations_to_u_v(gt)
        s_uv = self.flattened_pixel_locations_to_u_v(non_matches_b.unsqueeze(1))
        val = 1.0 / M_pixel * torch.clamp((gt_uv - s_uv).float().norm(2, 1), max=M_pixel)
        return val, gt_uv, s_uv

    def flattened_pixel_locations_to_u_v(self, flat):
        uv = flat.repeat(1, 2)
        uv[:, 0] = uv[:, 0] % self.image_width
        uv[:, 1] = uv[:, 1] // self.image_width              # py2 int division
        return uv

    def get_loss_original(self, image_a_pred, image_b_pred, matches_a, matches_b, non_matches_a, non_matches_b,
                          M_margin=0.5, non_match_loss_weight=1.0):
        num_matches = matches_a.size()[0]
        num_non_matches = non_matches_a.size()[0]
        a = torch.index_select(image_a_pred, 1, matches_a)
        b = torch.index_select(image_b_pred, 1, matches_b)
        match_loss = 1.0 / num_matches * (a - b).pow(2).sum()
        na = torch.index_select(image_a_pred, 1, non_matches_a)
        nb = torch.index_select(image_b_pred, 1, non_matches_b)
        pw = (na - nb).pow(2).sum(dim=2)
        pw = torch.add(torch.neg(pw), M_margin)
        non_match_loss = non_match_loss_weight * 1.0 / num_non_matches * torch.max(torch.zeros_like(pw), pw).sum()
        return match_loss + non_match_loss, match_loss, non_match_loss


def zero_loss(like=None):
    if like is not None:
        return torch.zeros(1, dtype=torch.float32, device=like.device)
    return torch.zeros(1, dtype=torch.float32)


def get_within_scene_loss(pcl, image_a_pred, image_b_pred, matches_a, matches_b,
                          masked_non_matches_a, masked_non_matches_b,
                          background_non_matches_a, background_non_matches_b,
                          blind_non_matches_a, blind_non_matches_b):
    # loss_composer.py:70-143
    cfg = pcl._config
    match_loss, masked_nm, h_m = pcl.get_loss_matched_and_non_matched_with_l2(
        image_a_pred, image_b_pred, matches_a, matches_b, masked_non_matches_a, masked_non_matches_b,
        M_descriptor=cfg["M_masked"])
    if cfg["use_l2_pixel_loss_on_background_non_matches"]:
        bg_nm, h_b = pcl.non_match_loss_with_l2_pixel_norm(
            image_a_pred, image_b_pred, matches_b, background_non_matches_a, background_non_matches_b,
            M_descriptor=cfg["M_background"])
    else:
        bg_nm, h_b = pcl.non_match_loss_descriptor_only(
            image_a_pred, image_b_pred, background_non_matches_a, background_non_matches_b,
            M_descriptor=cfg["M_background"])
    blind_nm = zero_loss(image_a_pred)
    h_x = 1
    if not is_empty(blind_non_matches_a.data):
        blind_nm, h_x = pcl.non_match_loss_descriptor_only(
            image_a_pred, image_b_pred, blind_non_matches_a, blind_non_matches_b, M_descriptor=cfg["M_masked"])
    total_h = max(h_m + h_b, 1)
    if cfg["scale_by_hard_negatives"]:
        scale_factor = total_h
        masked_scaled = masked_nm * 1.0 / max(h_m, 1)
        bg_scaled = bg_nm * 1.0 / max(h_b, 1)
        blind_scaled = blind_nm * 1.0 / max(h_x, 1)
    else:
        n_m = max(len(masked_non_matches_a), 1)
        n_b = max(len(background_non_matches_a), 1)
        n_x = max(len(blind_non_matches_a), 1)
        scale_factor = n_m + n_b
        masked_scaled = masked_nm * 1.0 / n_m
        bg_scaled = bg_nm * 1.0 / n_b
        blind_scaled = blind_nm * 1.0 / n_x
    non_match_loss = 1.0 / scale_factor * (masked_nm + bg_nm)
    loss = cfg["match_loss_weight"] * match_loss + cfg["non_match_loss_weight"] * non_match_loss
    return loss, match_loss, masked_scaled, bg_scaled, blind_scaled


def get_different_object_loss(pcl, image_a_pred, image_b_pred, blind_non_matches_a, blind_non_matches_b):
    # loss_composer.py:168-191
    scale_by_hard = pcl.config["scale_by_hard_negatives_DIFFERENT_OBJECT"]
    blind = zero_loss(image_a_pred)
    if not is_empty(blind_non_matches_a.data):
        blind, h = pcl.non_match_loss_descriptor_only(
            image_a_pred, image_b_pred, blind_non_matches_a, blind_non_matches_b,
            M_descriptor=pcl.config["M_background"])
        scale = max(h, 1) if scale_by_hard else max(len(blind_non_matches_a), 1)
        blind = 1.0 / scale * blind
    z = zero_loss(image_a_pred)
    return blind, z, z.clone(), z.clone(), blind


def get_same_object_across_scene_loss(pcl_obj, image_a_pred, image_b_pred, blind_non_matches_a, blind_non_matches_b):
    # loss_composer.py:193-212 -- the reference reads an undefined global ``pcl`` at :203
    # (NameError when the blind set is non-empty) and leaves num_hard_negatives unbound
    # when it is empty (UnboundLocalError at :205-206).  Both are kept, not fixed.
    blind = zero_loss(image_a_pred)
    if not is_empty(blind_non_matches_a.data):
        raise NameError("name 'pcl' is not defined")     # loss_composer.py:203
    raise UnboundLocalError("local variable 'num_hard_negatives' referenced before assignment")


def get_loss(pcl, match_type, image_a_pred, image_b_pred, matches_a, matches_b,
             masked_non_matches_a, masked_non_matches_b,
             background_non_matches_a, background_non_matches_b,
             blind_non_matches_a, blind_non_matches_b):
    # loss_composer.py:7-67
    T = SpartanDatasetDataType
    within = (pcl, image_a_pred, image_b_pred, matches_a, matches_b, masked_non_matches_a, masked_non_matches_b,
              background_non_matches_a, background_non_matches_b, blind_non_matches_a, blind_non_matches_b)
    if (match_type == T.SINGLE_OBJECT_WITHIN_SCENE).all():
        return get_within_scene_loss(*within)
    if (match_type == T.SINGLE_OBJECT_ACROSS_SCENE).all():
        return get_same_object_across_scene_loss(pcl, image_a_pred, image_b_pred,
                                                 blind_non_matches_a, blind_non_matches_b)
    if (match_type == T.DIFFERENT_OBJECT).all():
        return get_different_object_loss(pcl, image_a_pred, image_b_pred, blind_non_matches_a, blind_non_matches_b)
    if (match_type == T.MULTI_OBJECT).all():
        return get_within_scene_loss(*within)
    if (match_type == T.SYNTHETIC_MULTI_OBJECT).all():
        return get_within_scene_loss(*within)
    raise ValueError("Should only have above scenes?")


def batched_within_scene_loss(pcl, pred_a, pred_b, idx):
    """Batch semantics fixed by SURVEY.md 8a: per-pair reference loss, mean over the B pairs.
    pred_* are [B, P, D]; idx is a dict of [B, n] int64 tensors (or None / sentinel for blind)."""
    B = pred_a.shape[0]
    outs = []
    for i in range(B):
        blind_a = idx.get("blind_a")
        blind_b = idx.get("blind_b")
        ba = empty_tensor() if blind_a is None else blind_a[i]
        bb = empty_tensor() if blind_b is None else blind_b[i]
        outs.append(get_within_scene_loss(pcl, pred_a[i:i + 1], pred_b[i:i + 1],
                                          idx["matches_a"][i], idx["matches_b"][i],
                                          idx["masked_a"][i], idx["masked_b"][i],
                                          idx["background_a"][i], idx["background_b"][i], ba, bb))
    return tuple(sum(o[k].reshape(()) for o in outs) / B for k in range(5))


# ----------------------------------------------------------------------------- numpy float64
def np_match_loss(A, B, ma, mb):
    d = A[ma].astype(np.float64) - B[mb].astype(np.float64)
    return float((d * d).sum() / len(ma))


def np_hinge_vec(A, B, na, nb, M, invert=False):
    """Returns (l_j in float64, hard count).  The count follows the fp32 values the reference
    tests with nonzero(): recompute the hinge in float32 for the count."""
    diff64 = A[na].astype(np.float64) - B[nb].astype(np.float64)
    d64 = np.sqrt((diff64 * diff64).sum(1))
    l64 = np.maximum(d64 - M, 0) ** 2 if invert else np.maximum(M - d64, 0) ** 2
    diff32 = (A[na] - B[nb]).astype(np.float32)
    d32 = np.sqrt((diff32 * diff32).sum(1, dtype=np.float32)).astype(np.float32)
    h32 = np.maximum(d32 - np.float32(M), 0) if invert else np.maximum(np.float32(M) - d32, 0)
    l32 = (h32 * h32).astype(np.float32)
    return l64, int(np.count_nonzero(l32))


def np_pixel_weight(mb, nb, W, M_pixel):
    k = len(nb) // len(mb)
    gt = np.repeat(mb, k)
    du = (gt % W - nb % W).astype(np.float64)
    dv = (gt // W - nb // W).astype(np.float64)
    return np.minimum(np.sqrt(du * du + dv * dv), M_pixel) / M_pixel


def np_within_scene_loss(A, B, idx, cfg, W):
    """A, B: [P, D] float32 arrays; idx: dict of 1-D int64 arrays (blind_* optional)."""
    match = np_match_loss(A, B, idx["matches_a"], idx["matches_b"])
    lm, hm = np_hinge_vec(A, B, idx["masked_a"], idx["masked_b"], cfg["M_masked"])
    if cfg["use_l2_pixel_loss_on_masked_non_matches"]:
        lm = lm * np_pixel_weight(idx["matches_b"], idx["masked_b"], W, cfg["M_pixel"])
    lb, hb = np_hinge_vec(A, B, idx["background_a"], idx["background_b"], cfg["M_background"])
    if cfg["use_l2_pixel_loss_on_background_non_matches"]:
        lb = lb * np_pixel_weight(idx["matches_b"], idx["background_b"], W, cfg["M_pixel"])
    Sm, Sb = lm.sum(), lb.sum()
    Sx, hx, nx = 0.0, 1, 1
    if idx.get("blind_a") is not None and not (len(idx["blind_a"]) == 1 and idx["blind_a"][0] == -1):
        lx, hx = np_hinge_vec(A, B, idx["blind_a"], idx["blind_b"], cfg["M_masked"])
        Sx = lx.sum()
        nx = len(idx["blind_a"])
    if cfg["scale_by_hard_negatives"]:
        scale = max(hm + hb, 1)
        terms = (Sm / max(hm, 1), Sb / max(hb, 1), Sx / max(hx, 1))
    else:
        nm, nb_ = max(len(idx["masked_a"]), 1), max(len(idx["background_a"]), 1)
        scale = nm + nb_
        terms = (Sm / nm, Sb / nb_, Sx / max(nx, 1))
    non_match = (Sm + Sb) / scale
    loss = cfg["match_loss_weight"] * match + cfg["non_match_loss_weight"] * non_match
    return (loss, match) + terms, (hm, hb, hx)


# ----------------------------------------------------------------------------- non-match sampling
def create_non_correspondences_flat(matches_a_flat, image_shape, num_non_matches_per_match, img_b_mask, rand_u, rand_v):
    """Restatement of create_non_correspondences (correspondence_tools/correspondence_finder.py:276-405) followed by
    create_non_matches (dataset/spartan_dataset_masked.py:841-858) and flatten_uv_tensor (:1255-1264), with the uniform
    random numbers passed in (the reference draws them with torch.rand).  The "too close" perturbation is omitted because it
    is a no-op upstream: ``ones = torch.zeros_like(...)`` at correspondence_finder.py:354 keeps need_to_be_perturbed zero.
    -> (non_matches_a, non_matches_b) int64 flat indices."""
    H, W = image_shape
    n = len(matches_a_flat) * num_non_matches_per_match

    def rand_select_pixel():      # pytorch_rand_select_pixel, correspondence_finder.py:64-75
        return torch.floor(rand_u[:n] * W).long(), torch.floor(rand_v[:n] * H).long()

    if img_b_mask is not None:
        nz = torch.nonzero(img_b_mask.view(-1, 1).squeeze(1))
        if len(nz) == 0:
            u, v = rand_select_pixel()
        else:
            idx = torch.floor(rand_u[:n] * len(nz)).long()
            sel = torch.index_select(nz, 0, idx).squeeze(1)
            u, v = sel % W, sel // W
    else:
        u, v = rand_select_pixel()
    non_matches_b = v * W + u
    non_matches_a = torch.t(matches_a_flat.repeat(num_non_matches_per_match, 1)).contiguous().view(-1)
    return non_matches_a, non_matches_b


def batch_find_pixel_correspondences(img_a_depth, img_a_pose, img_b_depth, img_b_pose, uv_a_flat, K):
    """Restatement of correspondence_finder.batch_find_pixel_correspondences (:409-619) for given candidate pixels
    (``uv_a_flat`` int64 flat indices), CPU float32 torch ops in the reference's order.  Depth images: float arrays in
    millimetres.  -> (u_a, v_a) long, (u2, v2) float  (or (None, None))."""
    import numpy
    from numpy.linalg import inv
    H, W = img_a_depth.shape

    def invert_transform(T):        # :52-62 (numpy >= 1.13 overlap semantics: R^T and -R^T t)
        Tc = numpy.copy(T)
        R = numpy.transpose(Tc[0:3, 0:3]).copy()
        Tc[0:3, 0:3] = R
        Tc[0:3, 3] = -1.0 * R.dot(T[0:3, 3])
        return Tc

    def apply_transform_torch(vec3, T4):   # :64-68
        ones_row = torch.ones_like(vec3[0, :]).unsqueeze(0)
        return T4.mm(torch.cat((vec3, ones_row), 0))[0:3]

    uv_a = (uv_a_flat % W, uv_a_flat // W)
    K_inv = inv(K)
    da = torch.from_numpy(img_a_depth.astype(numpy.float32)).view(-1, 1)
    depth_vec = (torch.index_select(da, 0, uv_a_flat) * 1.0 / 1000.0).squeeze(1)
    nonzero = torch.nonzero(depth_vec)
    if nonzero.numel() == 0:
        return None, None
    nonzero = nonzero.squeeze(1)
    depth_vec = torch.index_select(depth_vec, 0, nonzero)
    u_a = torch.index_select(uv_a[0], 0, nonzero); v_a = torch.index_select(uv_a[1], 0, nonzero)
    full_vec = torch.stack((u_a.float() * depth_vec, v_a.float() * depth_vec, depth_vec))
    p_cam = torch.from_numpy(K_inv).float().mm(full_vec)
    p_world = apply_transform_torch(p_cam, torch.from_numpy(numpy.asarray(img_a_pose)).float())
    p_cam2 = apply_transform_torch(p_world, torch.from_numpy(invert_transform(numpy.asarray(img_b_pose))).float())
    vec2 = torch.from_numpy(numpy.asarray(K)).float().mm(p_cam2)
    u2, v2, z2 = vec2[0] / vec2[2], vec2[1] / vec2[2], vec2[2]
    eps = 1e-3
    u2 = torch.where(u2 < 0.0, torch.zeros_like(u2), u2); u2 = torch.where(u2 > W * 1.0 - eps, torch.zeros_like(u2), u2)
    keep = torch.nonzero(u2)
    if keep.numel() == 0:
        return None, None
    keep = keep.squeeze(1)
    u2, v2, z2, u_a, v_a = (torch.index_select(t, 0, keep) for t in (u2, v2, z2, u_a, v_a))
    v2 = torch.where(v2 < 0.0, torch.zeros_like(v2), v2); v2 = torch.where(v2 > H * 1.0 - eps, torch.zeros_like(v2), v2)
    keep = torch.nonzero(v2)
    if keep.numel() == 0:
        return None, None
    keep = keep.squeeze(1)
    u2, v2, z2, u_a, v_a = (torch.index_select(t, 0, keep) for t in (u2, v2, z2, u_a, v_a))
    db = torch.from_numpy(img_b_depth.astype(numpy.float32)).view(-1, 1)
    uv_b_flat = v2.long() * W + u2.long()
    depth2 = (torch.index_select(db, 0, uv_b_flat) * 1.0 / 1000).squeeze(1)
    z2 = z2 - 0.003
    depth2 = torch.where(depth2 < 0, torch.zeros_like(depth2), depth2)
    depth2 = torch.where(depth2 < z2, torch.zeros_like(depth2), depth2)
    keep = torch.nonzero(depth2)
    if keep.numel() == 0:
        return None, None
    keep = keep.squeeze(1)
    return (torch.index_select(u_a, 0, keep), torch.index_select(v_a, 0, keep)), \
           (torch.index_select(u2, 0, keep), torch.index_select(v2, 0, keep))

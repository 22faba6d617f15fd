"""Device-side non-match sampling: the index tensors ``loss_composer.get_loss`` consumes, produced on
the GPU instead of by the CPU dataset workers (correspondence_finder.create_non_correspondences +
SpartanDataset.create_non_matches + flatten_uv_tensor; see csrc/sampling.cu for the line references)."""
import torch

from . import _native as N


def sample_non_matches(matches_a, img_b_mask, image_shape, num_non_matches_per_match, generator=None, rand=None):
    """matches_a: int64 [Nm] flat indices (CUDA).  img_b_mask: float32 [H, W] CUDA tensor (nonzero = selectable) or None.
    -> (non_matches_a [Nm*k], non_matches_b [Nm*k]) int64 CUDA flat pixel indices, k = num_non_matches_per_match.
    ``rand`` (optional) = (rand_u, rand_v) float32 [Nm*k] uniform numbers to use instead of drawing them."""
    H, W = image_shape
    if not matches_a.is_cuda or matches_a.dtype != torch.int64:
        raise RuntimeError("matches_a must be an int64 CUDA tensor")
    dev = matches_a.device
    n = matches_a.numel() * int(num_non_matches_per_match)
    if rand is None:
        rand_u = torch.rand(n, device=dev, generator=generator)
        rand_v = torch.rand(n, device=dev, generator=generator)
    else:
        rand_u, rand_v = rand
        N.require_cuda_f32(rand_u, "rand_u"); N.require_cuda_f32(rand_v, "rand_v")
    if img_b_mask is not None:
        N.require_cuda_f32(img_b_mask, "img_b_mask")
        if tuple(img_b_mask.shape) != (H, W):
            raise RuntimeError("mask must have shape [H, W]")
    out_a = torch.empty(n, dtype=torch.int64, device=dev)
    out_b = torch.empty(n, dtype=torch.int64, device=dev)
    nb = N.lib.ddn_sample_non_matches_scratch_bytes(H, W)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    N.check(N.lib.ddn_sample_non_matches(N.ptr(img_b_mask), H, W, N.ptr(rand_u), N.ptr(rand_v), n, N.ptr(matches_a.contiguous()),
                                         int(num_non_matches_per_match), N.ptr(out_a), N.ptr(out_b), N.ptr(scratch), nb,
                                         N.stream_ptr()))
    return out_a, out_b


def find_pixel_correspondences(depth_a, pose_a, depth_b, pose_b, candidates_a, K):
    """Device-side ``batch_find_pixel_correspondences`` (correspondence_finder.py:409-619) for candidate pixels that were
    already drawn in image A (e.g. with ``sample_non_matches(..., img_a_mask, ...)[1]``).
    depth_a / depth_b: float32 [H, W] CUDA depth images in millimetres; pose_a / pose_b: 4x4 camera-to-world (numpy or
    nested lists); K: 3x3 intrinsics; candidates_a: int64 [n] CUDA flat pixels.
    -> (matches_a [m], matches_b [m]) int64 flat pixels and (u2, v2) float32 sub-pixel positions; m is read back once."""
    import ctypes
    import numpy as np
    N.require_cuda_f32(depth_a, "depth_a"); N.require_cuda_f32(depth_b, "depth_b")
    H, W = depth_a.shape
    n = candidates_a.numel()
    dev = depth_a.device
    Kd = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    Pa = np.ascontiguousarray(np.asarray(pose_a, dtype=np.float64).reshape(16))
    Pb = np.ascontiguousarray(np.asarray(pose_b, dtype=np.float64).reshape(16))
    out_a = torch.empty(n, dtype=torch.int64, device=dev); out_b = torch.empty(n, dtype=torch.int64, device=dev)
    u2 = torch.empty(n, dtype=torch.float32, device=dev); v2 = torch.empty(n, dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    nb = N.lib.ddn_find_pixel_correspondences_scratch_bytes(n)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    N.check(N.lib.ddn_find_pixel_correspondences(N.ptr(depth_a), N.ptr(depth_b), H, W, N.ptr(candidates_a.contiguous()), n,
                                                 vp(Kd), vp(Pa), vp(Pb), N.ptr(out_a), N.ptr(out_b), N.ptr(u2), N.ptr(v2),
                                                 N.ptr(count), N.ptr(scratch), nb, N.stream_ptr()))
    m = int(count.item())
    return out_a[:m], out_b[:m], u2[:m], v2[:m]

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

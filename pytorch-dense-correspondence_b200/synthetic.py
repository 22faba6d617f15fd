"""Seeded synthetic inputs shaped like what SpartanDataset.__getitem__ hands the training loop
(dense_correspondence/dataset/spartan_dataset_masked.py:111-151, :841-858): mean/std-normalised
RGB images and flat pixel indices n = u + W*v (doc/coordinate_conventions.md), with
``non_matches_a`` being every match repeated k times consecutively (spartan_dataset_masked.py:853-854).

Pure CPU torch with an explicit generator, so the same call reproduces the same tensors in the
build container and on the GPU box.
"""
import torch


def make_pair_batch(B, H=480, W=640, num_matches=1000, num_masked=1000, num_background=1000,
                    num_blind=0, seed=1):
    """Returns a dict of CPU tensors:
    img_a/img_b [B,3,H,W] fp32; matches_a/b [B,Nm]; masked_a/b [B,Nn_m]; background_a/b [B,Nn_b];
    blind_a/b [B,Nn_x] or None -- all int64 flat indices in [0, H*W)."""
    g = torch.Generator().manual_seed(seed)
    P = H * W
    out = {
        "img_a": torch.randn(B, 3, H, W, generator=g),
        "img_b": torch.randn(B, 3, H, W, generator=g),
        "matches_a": torch.randint(0, P, (B, num_matches), generator=g),
        "matches_b": torch.randint(0, P, (B, num_matches), generator=g),
    }

    def non_matches(n):
        if n == 0:
            return None, None
        k = max(n // num_matches, 1)
        a = out["matches_a"].repeat_interleave(k, dim=1)[:, :n]
        if a.shape[1] < n:  # n not a multiple of Nm: pad with fresh samples
            a = torch.cat([a, torch.randint(0, P, (B, n - a.shape[1]), generator=g)], 1)
        b = torch.randint(0, P, (B, n), generator=g)
        return a.contiguous(), b

    out["masked_a"], out["masked_b"] = non_matches(num_masked)
    out["background_a"], out["background_b"] = non_matches(num_background)
    out["blind_a"], out["blind_b"] = non_matches(num_blind)
    return out

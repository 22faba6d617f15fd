"""BASELINE.json configs[2]: loss-kernel HBM GB/s sweep (batch 16 pairs, D in {3,8,16}, non-matches per image swept).

Algorithmic bytes per index pair: forward 16 + 8*D, backward 16 + 24*D.  Times are CUDA events around
the loss launches only (the descriptor images are resident); each configuration touches 2 x 16 x D x 1.23 MB of
descriptors, gathered at random pixels.  Prints one JSON object; run on the GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdc_b200
from pdc_b200 import loss_composer, _native as N

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
B, H, W = 16, 480, 640
P = H * W
cfg = {"M_masked": 0.5, "M_background": 0.5, "M_pixel": 50, "match_loss_weight": 1.0, "non_match_loss_weight": 1.0,
       "use_l2_pixel_loss_on_masked_non_matches": False, "use_l2_pixel_loss_on_background_non_matches": False,
       "scale_by_hard_negatives": True, "scale_by_hard_negatives_DIFFERENT_OBJECT": True, "alpha_triplet": 0.1}
out = []
g = torch.Generator().manual_seed(0)
for D in (3, 8, 16):
    A = (0.2 * torch.randn(B, D, H, W, generator=g)).cuda().requires_grad_()
    Bt = (0.2 * torch.randn(B, D, H, W, generator=g)).cuda().requires_grad_()
    pa = A.view(B, D, P).permute(0, 2, 1); pb = Bt.view(B, D, P).permute(0, 2, 1)
    pcl = pdc_b200.PixelwiseContrastiveLoss([H, W], cfg)
    for nn in (1000, 5000, 50000, 750000):
        nm = 1000
        ma = torch.randint(0, P, (B, nm), generator=g).cuda(); mb = torch.randint(0, P, (B, nm), generator=g).cuda()
        na = ma.repeat_interleave(nn // nm, dim=1); nb = torch.randint(0, P, (B, nn), generator=g).cuda()
        blind = loss_composer.empty_tensor().cuda()
        mt = torch.zeros(B, dtype=torch.int64)
        def step():
            A.grad = None; Bt.grad = None
            five = loss_composer.get_loss(pcl, mt, pa, pb, ma, mb, na, nb, na, nb, blind, blind)
            five[0].backward()
        for _ in range(3):
            step()
        N.lib.ddn_profile_reset(); N.lib.ddn_profile_enable(1)
        torch.cuda.synchronize()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        N.lib.ddn_profile_enable(0)
        pr = N.profile_read()
        f, b = pr["loss_fwd"], pr["loss_bwd"]
        out.append({"D": D, "non_matches_per_image": nn, "index_pairs_per_launch": B * (nm + 2 * nn + 1),
                    "fwd_us": 1e3 * f["ms"] / f["launches"], "fwd_GBps": f["bytes"] / (f["ms"] * 1e-3) / 1e9,
                    "bwd_us": 1e3 * b["ms"] / b["launches"], "bwd_GBps": b["bytes"] / (b["ms"] * 1e-3) / 1e9})
        print(out[-1], file=sys.stderr, flush=True)
print(json.dumps({"workload": "configs[2] loss-kernel sweep, batch 16 pairs, 640x480", "hbm_peak_GBps": peak, "rows": out}))

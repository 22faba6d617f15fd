"""BASELINE.json configs[2]: loss-kernel HBM GB/s sweep (batch 16 pairs, D in {3,8,16}, non-matches per image swept).

Two paths are timed with CUDA events around the loss launches only (ddn_profile_*, descriptor images resident):
  fused    -- the product path: the loss evaluated THROUGH the bilinear upsample from the low-resolution maps
              (csrc/loss_lowres.cu; what loss_composer.get_loss runs on Resnet34_8s outputs)
  generic  -- gather from the full-resolution [B,D,H,W] descriptor images (csrc/loss.cu; what it runs on any other tensor)
`GBps` = ALGORITHMIC bytes per index pair (forward 16 + 8*D, backward 16 + 24*D -- the bytes a gather from a
materialised descriptor image has to move) / kernel time, i.e. the same yardstick for both; the fused path moves only the 16
index bytes through HBM, so its figure is an equivalent rate, not DRAM traffic.  Prints one JSON object; run on the GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pdc_b200
from pdc_b200 import loss_composer, ops, resnet_dilated, _native as N

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
B, H, W = 16, 480, 640
h, w = H // 8, W // 8
P = H * W
cfg = dict(pdc_b200.DEFAULT_LOSS_CONFIG)
out = []
g = torch.Generator().manual_seed(0)
sweep = [int(v) for v in os.environ.get("LOSS_SWEEP", "1000,5000,50000,750000").split(",")]
for D in [int(v) for v in os.environ.get("LOSS_DIMS", "3,8,16").split(",")]:
    low_a = (0.2 * torch.randn(B, h * w, D, generator=g)).cuda().requires_grad_()
    low_b = (0.2 * torch.randn(B, h * w, D, generator=g)).cuda().requires_grad_()

    def upsampled(low):      # [B, h*w, D] -> [B, D, H, W] exactly as the network's head does
        return ops.upsample_bilinear_forward(low.detach().view(B, h, w, D).permute(0, 3, 1, 2).contiguous(), H, W)
    A = upsampled(low_a).requires_grad_(); Bt = upsampled(low_b).requires_grad_()
    pcl = pdc_b200.PixelwiseContrastiveLoss([H, W], cfg)
    for nn in sweep:
        nm = 1000
        ma = torch.randint(0, P, (B, nm), generator=g).cuda(); mb = torch.randint(0, P, (B, nm), generator=g).cuda()
        na = ma.repeat_interleave(nn // nm, dim=1); nb = torch.randint(0, P, (B, nn), generator=g).cuda()
        blind = loss_composer.empty_tensor().cuda()
        mt = torch.zeros(B, dtype=torch.int64)
        row = {"D": D, "non_matches_per_image": nn, "index_pairs_per_launch": B * (nm + 2 * nn + 1)}
        losses = {}
        for path in ("fused", "generic"):
            pa = A.view(B, D, P).permute(0, 2, 1); pb = Bt.view(B, D, P).permute(0, 2, 1)
            if path == "fused":
                resnet_dilated.attach_lowres(pa, low_a, H, W); resnet_dilated.attach_lowres(pb, low_b, H, W)

            def step():
                A.grad = None; Bt.grad = None; low_a.grad = None; low_b.grad = None
                five = loss_composer.get_loss(pcl, mt, pa, pb, ma, mb, na, nb, na, nb, blind, blind)
                five[0].backward()
                return five[0]
            for _ in range(3):
                losses[path] = float(step())
            N.lib.ddn_profile_reset(); N.lib.ddn_profile_enable(1)
            torch.cuda.synchronize()
            for _ in range(10):
                step()
            torch.cuda.synchronize()
            N.lib.ddn_profile_enable(0)
            pr = N.profile_read()
            f, b = pr["loss_fwd"], pr["loss_bwd"]
            row[path] = {"fwd_us": 1e3 * f["ms"] / f["launches"], "fwd_GBps": f["bytes"] / (f["ms"] * 1e-3) / 1e9,
                         "fwd_frac_of_hbm_peak": f["bytes"] / (f["ms"] * 1e-3) / 1e9 / peak,
                         "bwd_us": 1e3 * b["ms"] / b["launches"], "bwd_GBps": b["bytes"] / (b["ms"] * 1e-3) / 1e9,
                         "bwd_frac_of_hbm_peak": b["bytes"] / (b["ms"] * 1e-3) / 1e9 / peak}
        row["loss_fused_vs_generic_rel_diff"] = abs(losses["fused"] - losses["generic"]) / abs(losses["generic"])
        out.append(row)
        print(row, file=sys.stderr, flush=True)
print(json.dumps({"workload": "configs[2] loss-kernel sweep, batch 16 pairs, 640x480 (C3 = D 16, 5000 non-matches per image)",
                  "hbm_peak_GBps": peak, "rows": out}))

"""Where does the HOST spend its time while it enqueues a step?  (No device sync inside the timed parts.)"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pdc_b200  # noqa: E402
from pdc_b200 import synthetic, loss_composer  # noqa: E402

dev = torch.device("cuda:0")
Bp, H, W, D = 8, 480, 640, 3
dcn = pdc_b200.DenseCorrespondenceNetwork.from_config({"descriptor_dimension": D, "image_width": W, "image_height": H}, load_stored_params=False).to(dev)
dcn.train()
pcl = pdc_b200.PixelwiseContrastiveLoss(image_shape=dcn.image_shape, config=pdc_b200.DEFAULT_LOSS_CONFIG)
host = synthetic.make_pair_batch(Bp, H, W, 1000, 1000, 1000, 0, seed=1)
d = {k: v.to(dev) for k, v in host.items() if v is not None}
mt = torch.zeros(Bp, dtype=torch.int64)
blind = loss_composer.empty_tensor().to(dev)


def step(timing=None):
    t0 = time.perf_counter()
    dcn.zero_grad(set_to_none=True)
    ya, yb = dcn.forward_pair(d["img_a"], d["img_b"])
    t1 = time.perf_counter()
    five = loss_composer.get_loss(pcl, mt, dcn.process_network_output(ya, Bp), dcn.process_network_output(yb, Bp), d["matches_a"], d["matches_b"],
                                  d["masked_a"], d["masked_b"], d["background_a"], d["background_b"], blind, blind)
    t2 = time.perf_counter()
    five[0].backward()
    t3 = time.perf_counter()
    if timing is not None:
        timing.append((t1 - t0, t2 - t1, t3 - t2))


for _ in range(3):
    step()
torch.cuda.synchronize()
for label, n in (("5 steps from an idle GPU", 5), ("20 steps", 20)):
    tm = []
    t0 = time.perf_counter()
    for _ in range(n):
        step(tm)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    f = sum(t[0] for t in tm) / n * 1e3; l = sum(t[1] for t in tm) / n * 1e3; b = sum(t[2] for t in tm) / n * 1e3
    print("%s: host enqueue %.2f ms/step (forward %.2f, loss %.2f, backward %.2f); until the GPU is done %.2f ms/step"
          % (label, t_host / n * 1e3, f, l, b, t_all / n * 1e3))
    print("   per-step host ms:", ["%.1f" % (sum(t) * 1e3) for t in tm])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[:4000])

"""GPU, >= 2 devices: the data-parallel step.  Each rank runs its own pairs; after GradientAllReducer the gradient
on every rank equals the mean over ranks of the per-rank gradients, bit-for-bit across ranks."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import pdc_b200
from pdc_b200 import loss_composer, synthetic, data_parallel as DP
from oracle import loss_oracle as LO
rank, world, local = DP.init_from_env()
dev = torch.device("cuda", local)
D, B, H, W = 3, 1, 64, 96
torch.manual_seed(0)
dcn = pdc_b200.DenseCorrespondenceNetwork.from_config({"descriptor_dimension": D, "image_width": W, "image_height": H}, load_stored_params=False)
if rank == 1:
    with torch.no_grad():
        for p in dcn.parameters(): p.add_(0.5)          # broadcast must undo this
DP.broadcast_parameters(dcn)
pcl = pdc_b200.PixelwiseContrastiveLoss(dcn.image_shape, dict(LO.DEFAULT_LOSS_CONFIG))
data = synthetic.make_pair_batch(B, H, W, 30, 60, 60, 0, seed=1 + rank)
d = {k: (v.to(dev) if v is not None else None) for k, v in data.items()}
blind = loss_composer.empty_tensor().to(dev)
pa = dcn.process_network_output(dcn.forward(d["img_a"]), B); pb = dcn.process_network_output(dcn.forward(d["img_b"]), B)
five = loss_composer.get_loss(pcl, torch.tensor([0]), pa, pb, d["matches_a"], d["matches_b"], d["masked_a"], d["masked_b"],
                              d["background_a"], d["background_b"], blind, blind)
five[0].backward()
local_grads = [p.grad.detach().clone() for p in dcn.parameters()]
red = DP.GradientAllReducer(dcn.parameters())
red()
ok = True
for p, g in zip(dcn.parameters(), local_grads):
    gathered = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    mean = sum(gathered) / world
    ok = ok and torch.allclose(p.grad, mean, rtol=1e-5, atol=1e-7)
    same = [torch.empty_like(p.grad) for _ in range(world)]
    dist.all_gather(same, p.grad.contiguous())
    ok = ok and all(torch.equal(same[0], s) for s in same)
# ---- the overlapped path: all-reduce issued per gradient bucket from inside the backward, 1/world folded into the cotangent
mean_flat = dcn.fcn.flat_gradient.detach().clone()                       # the explicit path's result for the same inputs
red2 = DP.GradientAllReducer(dcn.parameters(), module=dcn.fcn, overlap=True)
dcn.zero_grad(set_to_none=True)
pa = dcn.process_network_output(dcn.forward(d["img_a"]), B); pb = dcn.process_network_output(dcn.forward(d["img_b"]), B)
five = loss_composer.get_loss(pcl, torch.tensor([0]), pa, pb, d["matches_a"], d["matches_b"], d["masked_a"], d["masked_b"],
                              d["background_a"], d["background_b"], blind, blind)
five[0].backward()
red2()                                                                    # nothing left to do
over = dcn.fcn.flat_gradient.detach().clone()
ok = ok and red2.overlapped_steps == 2                                    # one per backward (image B's, then image A's)
err = float((over.double() - mean_flat.double()).norm() / mean_flat.double().norm())
ok = ok and err < 1e-4                                                    # BN running statistics moved between the two runs
allg = [torch.empty_like(over) for _ in range(world)]
dist.all_gather(allg, over)
ok = ok and all(torch.equal(allg[0], g) for g in allg)
print("RANK %%d overlapped: rel diff vs explicit path %%.2e, steps %%d" %% (rank, err, red2.overlapped_steps), flush=True)
red2.detach()
w0 = [torch.empty_like(dcn.fcn.flat_parameters) for _ in range(world)]
dist.all_gather(w0, dcn.fcn.flat_parameters)
ok = ok and all(torch.equal(w0[0], w) for w in w0)
print("RANK %%d ok=%%s flat_path=%%s bytes=%%d" %% (rank, ok, red.used_flat_path, red.bytes_last), flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_allreduced_gradient_is_mean_of_rank_gradients(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "RANK 0 ok=True" in out.stdout and "RANK 1 ok=True" in out.stdout
    assert "flat_path=True" in out.stdout          # the exchange ran on slices of the single flat gradient array

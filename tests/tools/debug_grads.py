"""Per-parameter gradient error: ours (GPU) vs CPU oracle, next to torch-GPU oracle vs CPU oracle (noise floor)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pdc_b200
from pdc_b200 import _native as N
from oracle.resnet34_8s_oracle import seeded_oracle

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
D, B, H, W = 3, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 96
prec = {"fp32": 0, "bf16x3": 1, "bf16": 2}[sys.argv[3] if len(sys.argv) > 3 else "fp32"]
gen = torch.Generator().manual_seed(11)
x = torch.randn(B, 3, H, W, generator=gen); cot = torch.randn(B, D, H, W, generator=gen)
o = seeded_oracle(D).train()
y = o(x); (y * cot).sum().backward()
og = seeded_oracle(D).train().cuda()
yg = og(x.cuda()); (yg * cot.cuda()).sum().backward()
net = pdc_b200.Resnet34_8s(num_classes=D, precision=prec); net.load_state_dict(seeded_oracle(D).state_dict()); net.cuda().train()
yo = net(x.cuda()); (yo * cot.cuda()).sum().backward()
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))
print("fwd: ours %.2e torch-gpu %.2e" % (rel(yo, y), rel(yg, y)))
po, pg, pn = dict(o.named_parameters()), dict(og.named_parameters()), dict(net.named_parameters())
for k in po:
    print("%-45s |g| %.3e  ours %.2e  torch-gpu %.2e" % (k, float(po[k].grad.norm()), rel(pn[k].grad, po[k].grad), rel(pg[k].grad, po[k].grad)))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from pdc_b200 import ops, _native as N
case = tuple(int(v) for v in sys.argv[1].split(","))
prec = int(sys.argv[2])
n, h, w, cin, cout, k, s, p, d = case
g = torch.Generator().manual_seed(1)
x = torch.randn(n, cin, h, w, generator=g); wt = torch.randn(cout, cin, k, k, generator=g) * 0.05
xr = x.clone().requires_grad_(); wr = wt.clone().requires_grad_()
y_ref = F.conv2d(xr, wr, None, s, p, d); dy = torch.randn(y_ref.shape, generator=g); y_ref.backward(dy)
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().norm())
y = ops.conv2d_forward(nhwc(x).cuda(), wt.cuda(), s, p, d, precision=prec); torch.cuda.synchronize(); print("fwd ok", rel(y, nhwc(y_ref.detach())), flush=True)
dx, dw = ops.conv2d_backward(nhwc(x).cuda(), wt.cuda(), nhwc(dy).cuda(), s, p, d, precision=prec); torch.cuda.synchronize()
print("bwd ok dx", rel(dx, nhwc(xr.grad)), "dw", rel(dw, wr.grad), flush=True)

"""Thin functional wrappers over the single-operator C-ABI entry points (used by the unit tests and as
building blocks).  NHWC fp32 CUDA tensors in, NHWC fp32 CUDA tensors out; no autograd here."""
import torch

from . import _native as N


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def conv2d_forward(x_nhwc, w_oihw, stride=1, pad=0, dil=1, precision=N.PRECISION_FP32_SIMT):
    N.require_cuda_f32(x_nhwc, "x"); N.require_cuda_f32(w_oihw, "w")
    n, h, w, cin = x_nhwc.shape
    cout, cin2, k, k2 = w_oihw.shape
    assert cin == cin2 and k == k2
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    y = torch.empty(n, ho, wo, cout, dtype=torch.float32, device=x_nhwc.device)
    nb = N.lib.ddn_conv2d_workspace_bytes(n, h, w, cin, cout, k, stride, pad, dil, precision)
    ws = _ws(nb, x_nhwc.device)
    N.check(N.lib.ddn_conv2d_forward(N.ptr(x_nhwc), N.ptr(w_oihw), N.ptr(y), n, h, w, cin, cout, k, stride, pad, dil,
                                     precision, N.ptr(ws), ws.numel(), N.stream_ptr()))
    return y


def conv2d_backward(x_nhwc, w_oihw, dy_nhwc, stride=1, pad=0, dil=1, need_dx=True, precision=N.PRECISION_FP32_SIMT):
    N.require_cuda_f32(x_nhwc, "x"); N.require_cuda_f32(w_oihw, "w"); N.require_cuda_f32(dy_nhwc, "dy")
    n, h, w, cin = x_nhwc.shape
    cout, _, k, _ = w_oihw.shape
    dx = torch.empty_like(x_nhwc) if need_dx else None
    dw = torch.empty_like(w_oihw)
    nb = N.lib.ddn_conv2d_workspace_bytes(n, h, w, cin, cout, k, stride, pad, dil, precision)
    ws = _ws(nb, x_nhwc.device)
    N.check(N.lib.ddn_conv2d_backward(N.ptr(x_nhwc), N.ptr(w_oihw), N.ptr(dy_nhwc), N.ptr(dx), N.ptr(dw),
                                      n, h, w, cin, cout, k, stride, pad, dil, precision, N.ptr(ws), ws.numel(),
                                      N.stream_ptr()))
    return dx, dw


def batchnorm_forward(x, gamma, beta, residual=None, relu=False, training=True, running_mean=None, running_var=None,
                      momentum=0.1, eps=1e-5):
    """x [..., C] channels-last.  -> (y, save_mean, save_invstd)"""
    N.require_cuda_f32(x, "x")
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    invstd = torch.empty(C, dtype=torch.float32, device=x.device)
    nb = N.lib.ddn_batchnorm_workspace_bytes(M, C)
    ws = _ws(nb, x.device)
    N.check(N.lib.ddn_batchnorm_forward(N.ptr(x), N.ptr(gamma), N.ptr(beta), N.ptr(residual), N.ptr(y), N.ptr(mean),
                                        N.ptr(invstd), N.ptr(running_mean), N.ptr(running_var), M, C, int(relu),
                                        int(training), momentum, eps, N.ptr(ws), ws.numel(), N.stream_ptr()))
    return y, mean, invstd


def batchnorm_backward(dy, x, y, gamma, mean, invstd, relu=False, need_residual_grad=False):
    N.require_cuda_f32(dy, "dy")
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
    dres = torch.empty_like(x) if need_residual_grad else None
    ws = _ws(N.lib.ddn_batchnorm_workspace_bytes(M, C), x.device)
    N.check(N.lib.ddn_batchnorm_backward(N.ptr(dy), N.ptr(x), N.ptr(y), N.ptr(gamma), N.ptr(mean), N.ptr(invstd),
                                         N.ptr(dx), N.ptr(dgamma), N.ptr(dbeta), N.ptr(dres), M, C, int(relu),
                                         N.ptr(ws), ws.numel(), N.stream_ptr()))
    return dx, dgamma, dbeta, dres


def upsample_bilinear_forward(x, H, W):
    """x [N,C,h,w] -> [N,C,H,W], align_corners=True."""
    N.require_cuda_f32(x, "x")
    n, c, h, w = x.shape
    y = torch.empty(n, c, H, W, dtype=torch.float32, device=x.device)
    N.check(N.lib.ddn_upsample_bilinear_forward(N.ptr(x), N.ptr(y), n * c, h, w, H, W, N.stream_ptr()))
    return y


def upsample_bilinear_backward(dy, h, w):
    N.require_cuda_f32(dy, "dy")
    n, c, H, W = dy.shape
    dx = torch.empty(n, c, h, w, dtype=torch.float32, device=dy.device)
    N.check(N.lib.ddn_upsample_bilinear_backward(N.ptr(dy), N.ptr(dx), n * c, h, w, H, W, N.stream_ptr()))
    return dx


def scale_inplace(t, scale):
    N.require_cuda_f32(t, "t")
    N.check(N.lib.ddn_scale_inplace(N.ptr(t), t.numel(), float(scale), N.stream_ptr()))
    return t

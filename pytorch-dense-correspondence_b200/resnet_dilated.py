"""Resnet34_8s -- drop-in for the backbone class the reference builds with
``getattr(resnet_dilated, "Resnet34_8s")(num_classes=D)``
(dense_correspondence/network/dense_correspondence_network.py:373-375;
original: external/pytorch-segmentation-detection/pytorch_segmentation_detection/models/resnet_dilated.py:283-322
on top of .../vision/torchvision/models/resnet.py:112-265).

Same constructor, same ``forward(x, feature_alignment=False)``, same 218 state-dict keys
(``resnet34_8s.conv1.weight`` ... ``resnet34_8s.fc.bias``), same train()/eval() BatchNorm semantics --
but the module holds no torch.nn layers: all learnable tensors are views into one flat fp32 array
and the whole forward / backward runs inside libddn_b200.so (hand-written sm_100a kernels) through
one autograd.Function.  CUDA only; there is no CPU path.

Differences from the reference constructor, on purpose: no ImageNet download (``pretrained=True`` at
resnet_dilated.py:292-295 needs the network); weights start from the reference's own initialisers
(He-normal convs resnet.py:174-180, fc ~ N(0, 0.01) resnet_dilated.py:305-308) and are normally
overwritten by ``load_state_dict``.
"""
import math

import torch
import torch.nn as nn

from . import _native as N

_BN_MOMENTUM = 0.1   # nn.BatchNorm2d defaults, resnet.py:46
_BN_EPS = 1e-5

_cache_nonce = [0]
_default_precision = [N.PRECISION_BF16X3]     # fp32-equivalent results on the tensor cores


def set_default_precision(p):
    """'fp32' (CUDA-core FFMA), 'bf16x3' (tcgen05, fp32-equivalent split) or 'bf16' (tcgen05 single pass)."""
    _default_precision[0] = {"fp32": N.PRECISION_FP32_SIMT, "bf16x3": N.PRECISION_BF16X3, "bf16": N.PRECISION_BF16}[p]


class _Holder(nn.Module):
    """A name-space node of the reference module tree (it owns parameters/buffers, never computes)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("Resnet34_8s sub-modules are parameter holders; call the top-level module")


class _Backbone(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, owner, *params):
        N.require_cuda_f32(x, "input image batch")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError("expected input of shape [N,3,H,W], got %s" % (tuple(x.shape),))
        B, _, H, W = x.shape
        D = owner.num_classes
        flat, bufs = owner._ensure_flat(x.device)
        training = 1 if owner.training else 0
        keep = bool(training) and any(ctx.needs_input_grad)   # grad mode is off inside Function.forward; this is the signal
        prec = owner.precision
        owner._register_weight_cache(flat, prec)
        ws_bytes = N.lib.ddn_resnet34_8s_workspace_bytes(B, H, W, D, training, prec)
        if ws_bytes == 0:
            raise N.DdnError("bad shape for Resnet34_8s: %s" % N.lib.ddn_last_error().decode())
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        y = torch.empty(B, D, H, W, dtype=torch.float32, device=x.device)
        N.check(N.lib.ddn_resnet34_8s_forward(N.ptr(x), N.ptr(flat), N.ptr(bufs), N.ptr(y), N.ptr(ws), ws_bytes,
                                              B, H, W, D, training, _BN_MOMENTUM, _BN_EPS, prec, N.stream_ptr()))
        if training:
            torch._foreach_add_(owner._nbt, 1)
        if keep:
            ctx.owner, ctx.ws, ctx.shape, ctx.prec = owner, ws, (B, H, W, D), prec
            ctx.param_version = owner._flat_version
        ctx.keep = keep
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.keep:
            raise RuntimeError("Resnet34_8s.backward: forward ran in eval mode or without grad; nothing was saved")
        owner = ctx.owner
        B, H, W, D = ctx.shape
        if owner._flat_version != ctx.param_version:
            raise RuntimeError("parameters were re-allocated between forward and backward")
        dy = dy.contiguous()
        N.require_cuda_f32(dy, "descriptor cotangent")
        flat, _ = owner._ensure_flat(dy.device)
        owner._register_weight_cache(flat, ctx.prec)
        grads = torch.empty_like(flat)
        N.check(N.lib.ddn_resnet34_8s_backward(N.ptr(dy), N.ptr(flat), N.ptr(grads), N.ptr(ctx.ws), ctx.ws.numel(),
                                               B, H, W, D, _BN_EPS, ctx.prec, N.stream_ptr()))
        ctx.ws = None
        if owner._pad_index is not None:           # alignment padding between tensors: keep it zero (it is all-reduced / stepped too)
            grads.index_fill_(0, owner._pad_index.to(grads.device), 0.0)
        # Gradients are accumulated by THIS function into one flat array that every p.grad aliases (like a fused
        # "main_grad"): the second backward of a step (image B, then image A) is one flat add instead of 110 small
        # AccumulateGrad kernels, and the data-parallel all-reduce / a fused optimizer see a single buffer.
        ps = owner._params
        fg = owner._flat_grad
        fresh = (fg is None or ps[0].grad is None or ps[-1].grad is None or fg.device != grads.device or
                 ps[0].grad.data_ptr() != fg.data_ptr() + 4 * owner._ptab[0][2] or
                 ps[-1].grad.data_ptr() != fg.data_ptr() + 4 * owner._ptab[-1][2])
        if fresh:                      # first backward since zero_grad(set_to_none=True)
            owner._flat_grad = grads
            for p, (_, s, o, n) in zip(ps, owner._ptab):
                p.grad = grads[o:o + n].view(s)
        else:
            fg.add_(grads)
        return (None, None) + (None,) * len(ps)


class Resnet34_8s(nn.Module):
    def __init__(self, num_classes=1000, precision=None):
        super().__init__()
        if not (1 <= num_classes <= 32):
            raise ValueError("this build supports descriptor dimensions 1..32 (got %d)" % num_classes)
        self.num_classes = num_classes
        self.precision = _default_precision[0] if precision is None else precision
        self._ptab = N.param_table(num_classes)
        self._btab = N.buffer_table()
        n_params = int(N.lib.ddn_resnet34_8s_param_count(num_classes))
        n_bufs = int(N.lib.ddn_resnet34_8s_buffer_count())
        self._flat = torch.zeros(n_params, dtype=torch.float32)
        self._flat_bufs = torch.zeros(n_bufs, dtype=torch.float32)
        self._flat_version = 0
        self._flat_grad = None
        pads = []
        for (_, _, off, n) in self._ptab:
            pads += list(range(off + n, (off + n + 3) // 4 * 4))
        self._pad_index = torch.tensor(pads, dtype=torch.long) if pads else None
        self._wcache = None
        self._wcache_nonce = 0
        self._param_epoch = 0
        self._params = []
        self._nbt = []
        root = _Holder()
        self.resnet34_8s = root
        buf_by_name = {name: (shape, off, n) for name, shape, off, n in self._btab}
        for name, shape, off, n in self._ptab:
            path = name.split(".")
            node = root
            for part in path[:-1]:
                if not hasattr(node, part):
                    node.add_module(part, _Holder())
                node = getattr(node, part)
            p = nn.Parameter(self._flat[off:off + n].view(shape))
            node.register_parameter(path[-1], p)
            self._params.append(p)
            prefix = ".".join(path[:-1])
            if path[-1] == "bias" and (prefix + ".running_mean") in buf_by_name:   # a BatchNorm: add its buffers
                for bname in ("running_mean", "running_var"):
                    s, o, m = buf_by_name[prefix + "." + bname]
                    node.register_buffer(bname, self._flat_bufs[o:o + m].view(s))
                node.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
                self._nbt.append(node.num_batches_tracked)
        self._holders_with_buffers = [m for m in root.modules() if "running_mean" in m._buffers]
        self.reset_parameters()

    # ---- initialisation (resnet.py:174-180, resnet_dilated.py:305-308)
    def reset_parameters(self):
        with torch.no_grad():
            for (name, shape, _, _), p in zip(self._ptab, self._params):
                if name.startswith("fc."):
                    p.normal_(0, 0.01) if name == "fc.weight" else p.zero_()
                elif len(shape) == 4:
                    p.normal_(0, math.sqrt(2.0 / (shape[2] * shape[3] * shape[0])))
                elif name.endswith(".weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()
            for m in self._holders_with_buffers:
                m.running_mean.zero_()
                m.running_var.fill_(1.0)
                m.num_batches_tracked.zero_()

    # ---- flat storage management
    def _is_flat(self, device):
        if self._flat.device != device or self._flat_bufs.device != device:
            return False
        base, bbase = self._flat.data_ptr(), self._flat_bufs.data_ptr()
        for (_, _, off, _), p in zip(self._ptab, self._params):
            if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                return False
        for m in self._holders_with_buffers:
            if m.running_mean.device != device:
                return False
        for (name, _, off, _) in self._btab:
            node = self.resnet34_8s
            parts = name.split(".")
            for part in parts[:-1]:
                node = getattr(node, part)
            if node._buffers[parts[-1]].data_ptr() != bbase + 4 * off:
                return False
        return True

    def _ensure_flat(self, device):
        """(Re)packs parameters and BN statistics into the two flat arrays the C ABI takes; a no-op
        unless something (.cuda(), .to(), p.data = ...) re-allocated them."""
        if self._is_flat(device):
            return self._flat, self._flat_bufs
        with torch.no_grad():
            flat = torch.zeros(self._flat.numel(), dtype=torch.float32, device=device)
            for (_, shape, off, n), p in zip(self._ptab, self._params):
                flat[off:off + n].copy_(p.detach().reshape(-1).to(device=device, dtype=torch.float32))
                p.data = flat[off:off + n].view(shape)
            bufs = torch.zeros(self._flat_bufs.numel(), dtype=torch.float32, device=device)
            for (name, shape, off, n) in self._btab:
                node = self.resnet34_8s
                parts = name.split(".")
                for part in parts[:-1]:
                    node = getattr(node, part)
                old = node._buffers[parts[-1]]
                bufs[off:off + n].copy_(old.detach().reshape(-1).to(device=device, dtype=torch.float32))
                node._buffers[parts[-1]] = bufs[off:off + n].view(shape)
            self._nbt = []
            for m in self._holders_with_buffers:
                m._buffers["num_batches_tracked"] = m._buffers["num_batches_tracked"].to(device)
                self._nbt.append(m._buffers["num_batches_tracked"])
            self._flat, self._flat_bufs = flat, bufs
            self._flat_version += 1
        return self._flat, self._flat_bufs

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        dev = self._params[0].device
        if dev.type == "cuda":
            self._ensure_flat(dev)
        else:
            self._flat = self._flat.to(dev)   # CPU copy only serves state_dict round trips
        return out

    def _register_weight_cache(self, flat, prec):
        """Packed bf16 weights are cached across the calls of a step and re-packed when the parameters change (the
        tensor version counter of the flat array advances on every in-place update of it or of any view)."""
        if prec == N.PRECISION_FP32_SIMT:
            return
        if self._wcache is None or self._wcache.device != flat.device:
            self._wcache = torch.empty(N.lib.ddn_resnet34_8s_weight_cache_bytes(self.num_classes), dtype=torch.uint8,
                                       device=flat.device)
            _cache_nonce[0] += 1           # a fresh buffer may reuse the address of a dead one: never look "unchanged"
            self._wcache_nonce = _cache_nonce[0]
        # p.data = view shares storage but NOT the autograd version counter with the flat array, so the key is the sum of the
        # parameters' own counters (every optimizer / load_state_dict / init write advances one of them) plus manual bumps
        version = (self._wcache_nonce << 44) + ((self._flat_version + self._param_epoch) << 32) + (sum(p._version for p in self._params) & 0xffffffff)
        N.check(N.lib.ddn_resnet34_8s_set_weight_cache(N.ptr(self._wcache), self._wcache.numel(), N.ptr(flat), version, prec))

    def mark_parameters_changed(self):
        """Call after writing parameters through a raw pointer / ``.data`` (anything autograd's version counters miss)."""
        self._param_epoch += 1

    @property
    def flat_gradient(self):
        """The flat fp32 array all ``p.grad`` alias after a backward (None before the first one / after zero_grad)."""
        if self._flat_grad is None or self._params[0].grad is None:
            return None
        return self._flat_grad

    @property
    def flat_parameters(self):
        """The single fp32 array every parameter aliases (valid after the module is on its device)."""
        return self._ensure_flat(self._params[0].device)[0]

    def forward(self, x, feature_alignment=False):
        if feature_alignment:
            raise NotImplementedError("feature_alignment=True is not on the dense-descriptor hot path "
                                      "(resnet_dilated.py:314 is never taken by the reference)")
        return _Backbone.apply(x, self, *self._params)

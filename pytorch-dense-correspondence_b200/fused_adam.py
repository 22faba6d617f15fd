"""FusedAdam -- torch.optim.Adam semantics (as constructed in dense_correspondence/training/training.py:133-145) executed as
ONE kernel launch over the flat parameter / gradient arrays of ``Resnet34_8s`` (csrc/optim.cu).

    optimizer = FusedAdam(dcn, lr=1e-4, weight_decay=1e-4)       # instead of optim.Adam(dcn.parameters(), ...)
    loss.backward(); optimizer.step()

``param_groups[0]['lr']`` is honoured every step, so the reference's ``adjust_learning_rate`` (training.py:544-558: x0.9
every 250 iterations) works unchanged.  ``state_dict()`` / ``load_state_dict()`` round-trip the two moment arrays and the
step count.  ``grad_scale`` lets a data-parallel run fold the 1/world of the gradient all-reduce into the update.
"""
import torch

from . import _native as N


class FusedAdam(object):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        fcn = getattr(module, "fcn", module)
        if not hasattr(fcn, "flat_parameters"):
            raise ValueError("FusedAdam needs a module whose parameters alias one flat array (Resnet34_8s)")
        self.fcn = fcn
        self.param_groups = [{"params": list(fcn.parameters()), "lr": lr, "betas": betas, "eps": eps,
                              "weight_decay": weight_decay}]
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None

    def zero_grad(self, set_to_none=True):
        for p in self.param_groups[0]["params"]:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _ensure_state(self, flat):
        if self.exp_avg is None or self.exp_avg.device != flat.device or self.exp_avg.numel() != flat.numel():
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        flat = self.fcn.flat_parameters
        grad = self.fcn.flat_gradient
        if grad is None:
            raise RuntimeError("FusedAdam.step(): no gradient (call loss.backward() first)")
        N.require_cuda_f32(flat, "parameters"); N.require_cuda_f32(grad, "gradients")
        self._ensure_state(flat)
        g = self.param_groups[0]
        self.step_count += 1
        N.check(N.lib.ddn_adam_step(N.ptr(flat), N.ptr(grad), N.ptr(self.exp_avg), N.ptr(self.exp_avg_sq), flat.numel(),
                                    self.step_count, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                    float(g["weight_decay"]), float(grad_scale), N.stream_ptr()))
        self.fcn.mark_parameters_changed()      # the kernel wrote through a raw pointer: invalidate the weight-pack cache

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg, self.exp_avg_sq = sd["exp_avg"], sd["exp_avg_sq"]
        self.param_groups[0].update(sd["param_groups"][0])

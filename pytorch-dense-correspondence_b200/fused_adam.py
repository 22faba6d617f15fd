"""FusedAdam -- torch.optim.Adam semantics (as constructed in dense_correspondence/training/training.py:133-145) executed as
ONE kernel launch over the flat parameter / gradient arrays of ``Resnet34_8s`` (csrc/optim.cu).

    optimizer = FusedAdam(dcn, lr=1e-4, weight_decay=1e-4)       # instead of optim.Adam(dcn.parameters(), ...)
    loss.backward(); optimizer.step()

``param_groups[0]['lr']`` is honoured every step, so the reference's ``adjust_learning_rate`` (training.py:544-558: x0.9
every 250 iterations) works unchanged.  ``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's own format, so the
reference's ``NNNNNN.pth.opt`` files (training.py:509-511) resume here and ours resume in torch.optim.Adam.  ``grad_scale`` lets a data-parallel run fold the 1/world of the gradient all-reduce into the update.
"""
import torch

from . import _native as N


def adjust_learning_rate(optimizer, iteration, steps_between_learning_rate_decay=250, learning_rate_decay=0.9):
    """DenseCorrespondenceTraining.adjust_learning_rate (dense_correspondence/training/training.py:544-558) as a free
    function: every ``steps_between_learning_rate_decay`` iterations every param group's lr is multiplied by
    ``learning_rate_decay`` (defaults: config/dense_correspondence/training/training.yaml:16-17).  Works on ``FusedAdam`` and on
    any torch optimizer (anything with ``param_groups``)."""
    if iteration % steps_between_learning_rate_decay == 0:
        for param_group in optimizer.param_groups:
            param_group["lr"] = param_group["lr"] * learning_rate_decay


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

    # ---- optimizer checkpoints: the reference saves `optimizer.state_dict()` next to every model file
    # (training.py:509-511, NNNNNN.pth.opt) and reloads it to resume (training.py:147-150), so the state is exchanged in
    # torch.optim.Adam's own format: per-parameter `exp_avg` / `exp_avg_sq` / `step`, parameters numbered in
    # `dcn.parameters()` order.  Files written by the reference load here, files written here load into torch.optim.Adam.
    def _offsets(self, flat):
        base, es = flat.data_ptr(), flat.element_size()
        return [((q.data_ptr() - base) // es, q.numel()) for q in self.param_groups[0]["params"]]

    def state_dict(self):
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g.setdefault("amsgrad", False); g.setdefault("maximize", False)
        params = self.param_groups[0]["params"]
        g["params"] = list(range(len(params)))
        state = {}
        if self.exp_avg is not None:
            flat = self.fcn.flat_parameters
            for i, (q, (off, n)) in enumerate(zip(params, self._offsets(flat))):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[off:off + n].view(q.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(q.shape).clone()}
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        if "state" not in sd:      # the flat layout this class wrote before it spoke torch's format
            self.step_count = int(sd["step"])
            self.exp_avg, self.exp_avg_sq = sd["exp_avg"], sd["exp_avg_sq"]
            self.param_groups[0].update(sd["param_groups"][0])
            return
        groups = sd["param_groups"]
        if len(groups) != 1:
            raise ValueError("FusedAdam.load_state_dict: expected one parameter group (training.py:133-145 builds one), got %d" % len(groups))
        params = self.param_groups[0]["params"]
        ids = list(groups[0]["params"])
        if len(ids) != len(params):
            raise ValueError("FusedAdam.load_state_dict: checkpoint has %d parameters, the network has %d" % (len(ids), len(params)))
        if groups[0].get("amsgrad", False):
            raise ValueError("FusedAdam.load_state_dict: amsgrad state is not supported")
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        state = sd["state"]
        if not state:
            self.step_count, self.exp_avg, self.exp_avg_sq = 0, None, None
            return
        flat = self.fcn.flat_parameters
        self._ensure_state(flat)
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        steps = set()
        for pid, q, (off, n) in zip(ids, params, self._offsets(flat)):
            st = state.get(pid, state.get(str(pid)))
            if st is None:
                raise ValueError("FusedAdam.load_state_dict: no state for parameter %r" % (pid,))
            if tuple(st["exp_avg"].shape) != tuple(q.shape):
                raise ValueError("FusedAdam.load_state_dict: parameter %r has shape %s in the checkpoint, %s here"
                                 % (pid, tuple(st["exp_avg"].shape), tuple(q.shape)))
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) != 1:
            raise ValueError("FusedAdam.load_state_dict: parameters disagree on the step count (%s); one fused update cannot resume that" % sorted(steps))
        self.step_count = steps.pop()

"""Flat-buffer Adam: torch.optim.Adam's update (defaults of scripts/train.py:258) as ONE kernel
launch over the model's flat parameter / gradient buffers, with the data-parallel gradient
all-reduce (one NCCL message) folded in front of it."""
from __future__ import annotations

import torch

from . import ops


class FlatAdam(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.Adam(model.parameters(), lr)`` on a nerf_mlp_b200.NeRFMLP.
    Subclasses Optimizer so LR schedulers (StepLR at scripts/train.py:260) work unchanged."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, world_size=None):
        self.model = model
        super().__init__(model._param_list, dict(lr=lr, betas=betas, eps=eps))
        self._step = 0
        self._m = None
        self._v = None
        self.process_group = process_group
        self._world = world_size

    def _world_size(self):
        if self._world is not None:
            return self._world
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size(self.process_group)
        return 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model = self.model
        model._ensure_flat()
        flat_g = model.flat_grad
        if flat_g is None or any(p.grad is None for p in model._param_list):
            raise RuntimeError("FlatAdam.step: gradients are not bound to the model's flat gradient buffer "
                               "(run backward through NeRFRenderer._render_rays / NeRFMLP.forward first)")
        world = self._world_size()
        if world > 1:
            torch.distributed.all_reduce(flat_g, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
        if self._m is None or self._m.device != model.flat_params.device:
            self._m = torch.zeros_like(model.flat_params)
            self._v = torch.zeros_like(model.flat_params)
        self._step += 1
        g = self.param_groups[0]
        ops.adam_step(model.flat_params, flat_g, self._m, self._v, self._step, g["lr"], g["betas"], g["eps"],
                      grad_scale=1.0 / world)
        model.mark_dirty()
        return loss

    def state_dict(self):
        sd = super().state_dict()
        sd["flat"] = {"step": self._step, "exp_avg": self._m, "exp_avg_sq": self._v}
        return sd

    def load_state_dict(self, sd):
        flat = sd.get("flat")
        super().load_state_dict({k: v for k, v in sd.items() if k != "flat"})
        if flat is not None:
            self._step = int(flat["step"])
            dev = self.model.flat_params.device
            self._m = None if flat["exp_avg"] is None else flat["exp_avg"].to(dev).clone()
            self._v = None if flat["exp_avg_sq"] is None else flat["exp_avg_sq"].to(dev).clone()

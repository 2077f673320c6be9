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
        # the full set of torch.optim.Adam defaults: a state_dict written by a FRESH FlatAdam must load into the
        # reference's torch.optim.Adam (which replaces its param_groups wholesale and then reads group['weight_decay'] etc.)
        super().__init__(model._param_list, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                                 foreach=None, capturable=False, differentiable=False, fused=None,
                                                 decoupled_weight_decay=False))
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

    def _ensure_moments(self):
        """(Re-)create zero Adam moments on the model's device (fresh optimizer, or a checkpoint saved before step 1)."""
        flat = self.model.flat_params
        if self._m is None or self._v is None or self._m.device != flat.device:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)

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
        self._ensure_moments()
        self._step += 1
        g = self.param_groups[0]
        ops.adam_step(model.flat_params, flat_g, self._m, self._v, self._step, g["lr"], g["betas"], g["eps"],
                      grad_scale=1.0 / world)
        model.mark_dirty()
        return loss

    # ---- checkpoint interop: the reference saves torch.optim.Adam's state_dict ------------------
    # (scripts/train.py:471-475 'optimizer_state_dict') and resumes from it (:303-306).  FlatAdam
    # writes and reads exactly that format -- per-parameter {'step','exp_avg','exp_avg_sq'} in
    # state_dict order, param_groups with every torch.optim.Adam key -- so checkpoints move between the
    # reference and this package both ways (tests/test_host_cpu.py: fresh FlatAdam -> torch Adam -> step).
    def state_dict(self):
        sd = super().state_dict()
        state = {}
        if self._m is not None and self._step > 0:
            step = torch.tensor(float(self._step), dtype=torch.float32)
            for i, (m, v) in enumerate(zip(self.model._views_of(self._m), self.model._views_of(self._v))):
                state[i] = {"step": step.clone(), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        sd["state"] = state
        return sd

    def load_state_dict(self, sd):
        state = sd.get("state", {})
        flat = sd.get("flat")                      # layout written by earlier versions of this class
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.model._param_list):
            raise ValueError("FlatAdam.load_state_dict: expected one param group over the model's 24 tensors")
        g = groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("FlatAdam implements plain Adam (no weight decay / amsgrad / maximize), "
                                      "the configuration of scripts/train.py:258")
        super().load_state_dict({"state": {}, "param_groups": groups})
        self.model._ensure_flat()
        dev = self.model.flat_params.device
        if flat is not None:
            self._step = int(flat["step"])
            self._m = None if flat["exp_avg"] is None else flat["exp_avg"].to(dev).clone()
            self._v = None if flat["exp_avg_sq"] is None else flat["exp_avg_sq"].to(dev).clone()
            return
        if not state:
            self._step, self._m, self._v = 0, None, None
            return
        ids = g["params"]
        if any(i not in state for i in ids):
            raise ValueError("FlatAdam.load_state_dict: optimizer state is missing parameters")
        steps = {int(float(state[i]["step"])) for i in ids}
        if len(steps) != 1:
            raise ValueError(f"FlatAdam.load_state_dict: parameters disagree on the step count {sorted(steps)}")
        self._step = steps.pop()
        if self._m is None or self._m.device != dev:
            self._m = torch.zeros_like(self.model.flat_params)
            self._v = torch.zeros_like(self.model.flat_params)
        with torch.no_grad():
            for i, m, v, p in zip(ids, self.model._views_of(self._m), self.model._views_of(self._v), self.model._param_list):
                ea, es = state[i]["exp_avg"], state[i]["exp_avg_sq"]
                if tuple(ea.shape) != tuple(p.shape):
                    raise ValueError(f"FlatAdam.load_state_dict: state {i} has shape {tuple(ea.shape)}, parameter {tuple(p.shape)}")
                m.copy_(ea)
                v.copy_(es)

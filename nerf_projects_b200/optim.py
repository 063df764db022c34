"""Fused Adam: the optimiser create_nerf returns (ipynb:905).  Same update
rule, defaults and ``state_dict`` layout as ``torch.optim.Adam`` (so the
reference's checkpoints -- 'optimizer_state_dict', ipynb:1293-1298 -- load and
save unchanged), but the whole step is ONE kernel launch over all parameter
tensors of both networks (``nerfb_adam_step``)."""
from __future__ import annotations

import torch
from torch.optim import Optimizer

from . import ops


class FusedAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam implements the configuration the reference uses: "
                                      "weight_decay=0, amsgrad=False")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.grad_scale = 1.0   # multiplied into every gradient inside the kernel (used by data parallel)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            step_no = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item())
                if step_no is None:
                    step_no = s
                if s != step_no:
                    # parameters with different step counts: flush what we have, start a new launch
                    ops.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                                  step_no, self.grad_scale)
                    ps, gs, ms, vs, step_no = [], [], [], [], s
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            if ps:
                ops.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                              step_no, self.grad_scale)
        # parameters were updated in place by the kernel: bump version counters so caches
        # keyed on them (packed bf16 weights) are refreshed
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    torch.autograd.graph.increment_version(p)
        return loss

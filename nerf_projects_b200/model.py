"""The NeRF MLP -- drop-in for /root/reference/nerf/nerf.py (same constructor,
same ``state_dict`` keys and shapes, ``load_weights_from_keras``).

``forward(x)`` runs the fp32 CUDA linear kernels layer by layer (differentiable
through ``autograd.LinearFn``) so the module also works inside user-written
``network_query_fn`` closures.  The render path (`render.render_rays`) does
not go through ``forward``: it recognises this module and runs the fused
kernels in ``mlp.py`` on its parameters instead.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import autograd as AG


class NeRF(nn.Module):
    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.output_ch = output_ch
        self.skips = list(skips)
        self.use_viewdirs = use_viewdirs
        # nerf.py:32-55 -- identical module names so checkpoints drop in
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self._packed = None
        self._packed_key = None
        self._packed_bwd = None
        self._packed_bwd_key = None

    # nerf.py:57-111
    def forward(self, x):
        lead = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1])
        pts, views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        h = pts
        for i, lin in enumerate(self.pts_linears):
            h = AG.LinearFn.apply(h, lin.weight, lin.bias, True)
            if i in self.skips:
                h = torch.cat([pts, h], -1)
        if self.use_viewdirs:
            alpha = AG.LinearFn.apply(h, self.alpha_linear.weight, self.alpha_linear.bias, False)
            feature = AG.LinearFn.apply(h, self.feature_linear.weight, self.feature_linear.bias, False)
            h = torch.cat([feature, views], -1)
            for lin in self.views_linears:
                h = AG.LinearFn.apply(h, lin.weight, lin.bias, True)
            rgb = AG.LinearFn.apply(h, self.rgb_linear.weight, self.rgb_linear.bias, False)
            out = torch.cat([rgb, alpha], -1)
        else:
            out = AG.LinearFn.apply(h, self.output_linear.weight, self.output_linear.bias, False)
        return out.reshape(*lead, out.shape[-1])

    # nerf.py:113-146
    def load_weights_from_keras(self, weights):
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        dev = self.pts_linears[0].weight.device

        def put(lin, k):
            # same values as the reference's transposed views, stored contiguously (the kernels read row-major weights)
            lin.weight.data = torch.from_numpy(np.ascontiguousarray(np.transpose(weights[k]), dtype=np.float32)).to(dev)
            lin.bias.data = torch.from_numpy(np.ascontiguousarray(np.transpose(weights[k + 1]), dtype=np.float32)).to(dev)
        for i in range(self.D):
            put(self.pts_linears[i], 2 * i)
        put(self.feature_linear, 2 * self.D)
        put(self.views_linears[0], 2 * self.D + 2)
        put(self.rgb_linear, 2 * self.D + 4)
        put(self.alpha_linear, 2 * self.D + 6)
        self.invalidate_packed()

    def invalidate_packed(self):
        """Drop the cached bf16 weight images of the fused kernels.  They are keyed on (data_ptr, _version) of every
        parameter, which in-place writes through ``p.data`` (``p.data.copy_()``, EMA updates, legacy init idioms) do NOT
        change: call this after such a write.  ``load_state_dict``, ``.to()``/``.cuda()``, optimizer steps and
        ``load_weights_from_keras`` are covered automatically."""
        self._packed = None
        self._packed_key = None
        self._packed_bwd = None
        self._packed_bwd_key = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_packed()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed()
        return out

    # ---- helpers for the fused kernels -------------------------------
    def is_canonical(self) -> bool:
        """Architecture the tensor-core kernel is specialised for."""
        ok = (self.D == 8 and self.W == 256 and self.skips == [4] and self.input_ch == 63)
        if self.use_viewdirs:
            return ok and self.input_ch_views == 27
        return ok and 1 <= self.output_ch <= 8 and self.output_ch >= 4

"""Positional encoding -- drop-in for /root/reference/nerf/embedder.py.

``get_embedder(multires, i)`` returns ``(embed_fn, out_dim)`` like the
reference (embedder.py:82-116).  ``embed_fn`` runs the CUDA encoding kernel
(``nerfb_embed``); the fused render path never calls it -- it reads
``embed_fn.multires`` and computes the encoding in registers instead.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class Embedder:
    """Same kwargs contract as the reference class (embedder.py:8-80); only the
    configuration the reference ever uses is supported: include_input=True,
    log_sampling=True, periodic_fns=[sin, cos], input_dims=3."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if not kwargs.get("include_input", True) or not kwargs.get("log_sampling", True) \
                or kwargs.get("input_dims", 3) != 3:
            raise NotImplementedError("only include_input / log_sampling / 3-D inputs are implemented in CUDA")
        fns = kwargs.get("periodic_fns", [torch.sin, torch.cos])
        if list(fns) != [torch.sin, torch.cos]:
            raise NotImplementedError("periodic_fns must be [torch.sin, torch.cos]")
        self.num_freqs = int(kwargs["num_freqs"])
        if int(kwargs.get("max_freq_log2", self.num_freqs - 1)) != self.num_freqs - 1:
            raise NotImplementedError("max_freq_log2 must equal num_freqs-1 (frequencies 2^k)")
        self.out_dim = 3 + 6 * self.num_freqs

    def embed(self, inputs: torch.Tensor) -> torch.Tensor:
        lead = inputs.shape[:-1]
        return ops.embed(inputs, self.num_freqs).reshape(*lead, self.out_dim)


class _EmbedFn:
    """Callable with the encoder's configuration attached (used by the fused path)."""

    def __init__(self, multires: int):
        self.multires = multires
        self._e = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                           log_sampling=True, periodic_fns=[torch.sin, torch.cos])
        self.out_dim = self._e.out_dim

    def __call__(self, x):
        return self._e.embed(x)


class _Identity(nn.Identity):
    multires = -1
    out_dim = 3


def get_embedder(multires, i=0):
    """embedder.py:82-116: i == -1 -> (nn.Identity(), 3); else gamma() with
    ``multires`` frequencies, out_dim = 3 + 6*multires."""
    if i == -1:
        return _Identity(), 3
    fn = _EmbedFn(int(multires))
    return fn, fn.out_dim

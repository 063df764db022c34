"""Multi-GPU: one process per GPU (torch.distributed, NCCL over NVLink).

The path shards by rays (independent units): every rank renders a contiguous
slice of the flattened ray list with replicated weights; the only
collectives are ONE all-gather of the rendered tiles per frame and ONE
all-reduce of the flat parameter-gradient bucket per training step
(SURVEY.md 8e).  Nothing else crosses GPUs.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import rng as RNG


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous slice [lo, hi) of n rays for `rank`, plus the padded per-rank length."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def _rng_ctx(rng_seed, lo):
    """keyed_rng context positioned at the rank's first ray (draws independent of the world size)."""
    return RNG.keyed_rng(rng_seed, ray_offset=lo) if rng_seed is not None else contextlib.nullcontext()


def render_sharded(render_fn, H, W, K, rays_o, rays_d, group=None, upload_to=None, return_packed=False,
                   rng_seed=None, **kwargs):
    """Ray-sharded render of one frame: each rank renders rays[lo:hi] with
    ``render_fn`` (= nerf_projects_b200.render), then one all_gather assembles
    [rgb(3), disp, acc] for the full frame on every rank.  With ``upload_to=device`` the rays
    are (pinned) host tensors and each rank copies only its own slice to its GPU.  Returns (rgb [N,3], disp [N], acc [N]) for the N input rays, or the
    packed [N,5] tile tensor with ``return_packed=True``.  ``rng_seed`` keys perturb / noise draws by
    (seed, global ray index), so the frame does not depend on the number of ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays_o.shape[0]
    lo, hi, per = shard_bounds(n, world, rank)
    if upload_to is None:
        dev = rays_o.device
        my_o, my_d = rays_o[lo:hi], rays_d[lo:hi]
    else:
        # host (pinned) rays: every rank uploads only ITS slice to its GPU
        dev = torch.device(upload_to)
        my_o = rays_o[lo:hi].to(dev, non_blocking=True)
        my_d = rays_d[lo:hi].to(dev, non_blocking=True)
    tile = torch.zeros((per, 5), dtype=torch.float32, device=dev)
    if hi > lo:
        with _rng_ctx(rng_seed, lo):
            rgb, disp, acc, _ = render_fn(H, W, K, rays=(my_o, my_d), **kwargs)
        tile[:hi - lo, 0:3] = rgb
        tile[:hi - lo, 3] = disp
        tile[:hi - lo, 4] = acc
    if world == 1:
        full = tile
    else:
        full = torch.empty((world * per, 5), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(full, tile, group=group)
    full = full[:n]
    if return_packed:          # [N,5] = rgb(3), disp, acc in one contiguous tensor (one D2H copy for the caller)
        return full
    return full[:, 0:3], full[:, 3], full[:, 4]


class GradBucket:
    """Flat fp32 bucket for the parameter gradients of both networks
    (2 x 595 844 floats = 4.77 MB): one all-reduce per step."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def allreduce_mean(self, group=None):
        """sum over ranks of d(sum-loss)/dp divided by world (the MSE is a mean over the
        global batch); gradients are written back into p.grad."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return                               # nothing to average: the gradients stay where backward() put them
        have = [(v, p.grad) for p, v in zip(self.params, self.views) if p.grad is not None]
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
        if have:                                 # one multi-tensor copy instead of one launch per parameter
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            p.grad = v


def train_step_sharded(render_fn, loss_fn, optimizer, bucket: GradBucket, H, W, K, rays_o, rays_d, target,
                       group=None, rng_seed=None, **kwargs):
    """Data-parallel step: the global ray batch is split across ranks, each rank
    does forward/backward on its slice, gradients are averaged with one
    all-reduce, every rank applies the identical Adam update.  ``rng_seed`` (e.g.
    ``rng.step_seed(seed, global_step)``) keys the perturb / noise draws by global ray index."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays_o.shape[0]
    lo, hi, _ = shard_bounds(n, world, rank)
    with _rng_ctx(rng_seed, lo):
        rgb, disp, acc, extras = render_fn(H, W, K, rays=(rays_o[lo:hi], rays_d[lo:hi]), **kwargs)
    loss = loss_fn(rgb, target[lo:hi])
    if "rgb0" in extras:
        loss = loss + loss_fn(extras["rgb0"], target[lo:hi])
    optimizer.zero_grad()
    loss.backward()
    bucket.allreduce_mean(group)
    optimizer.step()
    return loss.detach()

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

SHARD_ALIGN = 8     # rays: 8 x 64 coarse samples = one 512-row super-tile of the fused MLP kernel, 8 x 192 fine = three


def world_rank(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_bounds(n: int, world: int, rank: int, align: int = 1) -> Tuple[int, int, int]:
    """Contiguous slice [lo, hi) of n rays for `rank`, plus the padded per-rank length (a multiple of ``align``, so a
    shard is whole super-tiles of the fused kernel and only the LAST rank has a ragged tail)."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def _rng_ctx(rng_seed, lo):
    """keyed_rng context positioned at the rank's first ray (draws independent of the world size)."""
    return RNG.keyed_rng(rng_seed, ray_offset=lo) if rng_seed is not None else contextlib.nullcontext()


_tiles: Dict[tuple, List[torch.Tensor]] = {}


def _tile_buffers(dev, per, world, slot):
    """Persistent [per,5] send tile and [world*per,5] gather buffer (two slots, so the gather of frame k can still be in
    flight while frame k+1 is rendered)."""
    key = (dev, per, world, slot)
    buf = _tiles.get(key)
    if buf is None:
        if len(_tiles) > 16:
            _tiles.clear()
        buf = [torch.empty((per, 5), dtype=torch.float32, device=dev),
               torch.empty((world * per, 5), dtype=torch.float32, device=dev) if world > 1 else None]
        _tiles[key] = buf
    return buf


class ShardedFrame:
    """Handle of one ray-sharded frame whose all-gather may still be in flight: ``wait()`` makes the current stream wait
    for it and returns the packed [N,5] tiles (rgb(3), disp, acc)."""

    def __init__(self, full, n, work):
        self._full, self._n, self._work = full, n, work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._full[:self._n]

    def maps(self):
        full = self.wait()
        return full[:, 0:3], full[:, 3], full[:, 4]


def render_sharded(render_fn, H, W, K, rays_o, rays_d, group=None, upload_to=None, return_packed=False,
                   rng_seed=None, async_gather=False, slot=0, **kwargs):
    """Ray-sharded render of one frame: each rank renders rays[lo:hi] with
    ``render_fn`` (= nerf_projects_b200.render), then one all_gather assembles
    [rgb(3), disp, acc] for the full frame on every rank.  With ``upload_to=device`` the rays
    are (pinned) host tensors and each rank copies only its own slice to its GPU.  Returns (rgb [N,3], disp [N], acc [N])
    for the N input rays, or the packed [N,5] tile tensor with ``return_packed=True``.  ``rng_seed`` keys perturb / noise
    draws by (seed, global ray index), so the frame does not depend on the number of ranks.

    The rank's maps are packed into a persistent [per,5] tile with ONE launch and gathered into a persistent buffer
    (valid until the next call with the same ``slot``).  ``async_gather=True`` returns a ``ShardedFrame`` handle at once:
    the gather runs on NCCL's stream under the next frame's render, so ranks do not meet once per frame."""
    world, rank = world_rank(group)
    n = rays_o.shape[0]
    lo, hi, per = shard_bounds(n, world, rank, SHARD_ALIGN)
    if upload_to is None:
        dev = rays_o.device
        my_o, my_d = rays_o[lo:hi], rays_d[lo:hi]
    else:
        # host (pinned) rays: every rank uploads only ITS slice to its GPU
        dev = torch.device(upload_to)
        my_o = rays_o[lo:hi].to(dev, non_blocking=True)
        my_d = rays_d[lo:hi].to(dev, non_blocking=True)
    tile, full = _tile_buffers(dev, per, world, slot)
    if hi > lo:
        with _rng_ctx(rng_seed, lo):
            rgb, disp, acc, _ = render_fn(H, W, K, rays=(my_o, my_d), **kwargs)
        torch.cat([rgb.reshape(-1, 3), disp.reshape(-1, 1), acc.reshape(-1, 1)], -1, out=tile[:hi - lo])
    work = None
    if world == 1:
        full = tile
    else:
        work = dist.all_gather_into_tensor(full, tile, group=group, async_op=True)
    frame = ShardedFrame(full, n, work)
    if async_gather:
        return frame
    if return_packed:          # [N,5] = rgb(3), disp, acc in one contiguous tensor (one D2H copy for the caller)
        return frame.wait()
    return frame.maps()


def broadcast_parameters(params: Sequence[torch.Tensor], optimizer=None, src: int = 0, group=None):
    """Make every replica identical to rank ``src``: parameters and -- when given -- the optimizer's per-parameter state
    (exp_avg / exp_avg_sq / step) and hyper-parameters.  Data parallel training all-reduces gradients only, so replicas
    that start different (per-rank RNG state before create_nerf, a checkpoint only rank 0 can see) would stay different."""
    world, _ = world_rank(group)
    if world == 1:
        return
    with torch.no_grad():
        for p in params:
            dist.broadcast(p.data, src=src, group=group)
            torch.autograd.graph.increment_version(p)          # caches keyed on the version (packed bf16 weights) refresh
    if optimizer is not None:
        obj = [optimizer.state_dict() if dist.get_rank(group) == src else None]
        dist.broadcast_object_list(obj, src=src, group=group)
        if dist.get_rank(group) != src:
            optimizer.load_state_dict(obj[0])


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

    def allreduce_sum(self, group=None):
        """Sum of the ranks' gradients, written back into p.grad.  Each rank's loss is already weighted by its share of
        the global batch (train_step_sharded), so the sum IS the gradient of the global-batch mean -- also for uneven
        or empty shards.  Ranks whose parameters have no gradient (empty shard) contribute zeros."""
        world, _ = world_rank(group)
        if world == 1:
            return                               # nothing to reduce: the gradients stay where backward() put them
        have = [(v, p.grad) for p, v in zip(self.params, self.views) if p.grad is not None]
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
        if have:                                 # one multi-tensor copy instead of one launch per parameter
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def allreduce_mean(self, group=None):
        """Mean over ranks (equal shards with per-rank mean losses)."""
        world, _ = world_rank(group)
        if world == 1:
            return
        self.allreduce_sum(group)
        self.flat.mul_(1.0 / world)


def train_step_sharded(render_fn, loss_fn, optimizer, bucket: GradBucket, H, W, K, rays_o, rays_d, target,
                       group=None, rng_seed=None, **kwargs):
    """Data-parallel step: the global ray batch is split across ranks, each rank
    does forward/backward on its slice, gradients are summed with one
    all-reduce, every rank applies the identical Adam update.  ``rng_seed`` (e.g.
    ``rng.step_seed(seed, global_step)``) keys the perturb / noise draws by global ray index.

    ``loss_fn`` is a MEAN over its rays (img2mse): the rank's loss is weighted by (rays of this rank) / (rays of the
    batch), so the reduced gradient is that of the global-batch mean whatever the split -- uneven shards (n % world != 0)
    and empty ones (n < world, the short last batch of an epoch) included; a rank with no rays skips the render and
    joins the all-reduce with zero gradients.  Returns the global loss of the step (detached, same on every rank after
    one scalar all-reduce)."""
    world, rank = world_rank(group)
    n = rays_o.shape[0]
    lo, hi, _ = shard_bounds(n, world, rank, SHARD_ALIGN if n >= world * SHARD_ALIGN else 1)
    optimizer.zero_grad()
    if hi > lo:
        with _rng_ctx(rng_seed, lo):
            rgb, disp, acc, extras = render_fn(H, W, K, rays=(rays_o[lo:hi], rays_d[lo:hi]), **kwargs)
        loss = loss_fn(rgb, target[lo:hi])
        if "rgb0" in extras:
            loss = loss + loss_fn(extras["rgb0"], target[lo:hi])
        if world > 1:
            loss = loss * ((hi - lo) / n)
        loss.backward()
        total = loss.detach().clone()
    else:
        total = torch.zeros((), dtype=torch.float32, device=bucket.flat.device)
    bucket.allreduce_sum(group)
    optimizer.step()
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total

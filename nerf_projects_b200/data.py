"""Device-side ray-batch producer for the training loop -- the `use_batching` / per-image sampling block of
train() (ipynb:1167-1255) without its host work.

The reference either precomputes rays for ALL training pixels on the host (`rays_rgb`, [N*H*W, 3, 3] fp32,
np.random.shuffle, sliding window of N_rand rows, reshuffle at epoch end) or, per step, picks one image, builds
its H x W rays with get_rays, builds a (possibly centre-cropped) coordinate grid and draws N_rand coordinates
with np.random.choice(replace=False).  At a few milliseconds per optimisation step both become the bottleneck.
Here images and poses live on the GPU, the selection is a device permutation, and ONE kernel
(`nerfb_gather_ray_batch`) generates the selected rays from (pose, intrinsics, pixel id) with get_rays' exact
arithmetic and gathers the target colours: no per-image ray tensors, no host round trip.

Semantics kept: sampling without replacement; batching mode walks a permutation of all training pixels and
reshuffles when the window passes the end (ipynb:1236-1242); per-image mode draws the image uniformly from
i_train and applies the centre crop of size 2*int(H//2*frac) x 2*int(W//2*frac) while i < precrop_iters
(ipynb:1254-1271).  The random streams are torch's, not numpy's (the reference's draws are not reproducible
across libraries either).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


def _intrinsics(K):
    K = K.detach().cpu().numpy() if isinstance(K, torch.Tensor) else np.asarray(K)
    return float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])


def precrop_pixel_ids(H: int, W: int, frac: float) -> torch.Tensor:
    """Pixel ids (row*W + col) of the centre crop of train() (ipynb:1256-1262): rows H//2-dH .. H//2+dH-1."""
    dH, dW = int(H // 2 * frac), int(W // 2 * frac)
    rows = torch.arange(H // 2 - dH, H // 2 + dH)
    cols = torch.arange(W // 2 - dW, W // 2 + dW)
    return (rows[:, None] * W + cols[None, :]).reshape(-1)


class RayBatcher:
    """images: [N,H,W,3] float (already composited on white / RGB only), poses: [N,3,4] or [N,4,4], K: 3x3.
    ``next_batch(i)`` returns (batch_rays [2,N_rand,3], target_s [N_rand,3]) on the device, ready for
    ``render(H, W, K, rays=batch_rays, ...)`` exactly like the reference's loop body."""

    def __init__(self, images, poses, H: int, W: int, K, i_train: Optional[Sequence[int]] = None, N_rand: int = 1024,
                 use_batching: bool = True, precrop_iters: int = 0, precrop_frac: float = 0.5, device="cuda",
                 seed: Optional[int] = None):
        self.H, self.W = int(H), int(W)
        self.fx, self.fy, self.cx, self.cy = _intrinsics(K)
        dev = torch.device(device)
        self.images = torch.as_tensor(images, dtype=torch.float32).to(dev).contiguous()
        self.poses = torch.as_tensor(poses, dtype=torch.float32).to(dev).contiguous()
        n_img = self.images.shape[0]
        self.i_train_host = [int(i) for i in (range(n_img) if i_train is None else i_train)]   # host copy: no device read per step
        self.i_train = torch.as_tensor(self.i_train_host, dtype=torch.int64, device=dev)
        self.N_rand = int(N_rand)
        self.use_batching = bool(use_batching)
        self.precrop_iters, self.precrop_frac = int(precrop_iters), float(precrop_frac)
        self.gen = torch.Generator(device=dev)
        # per-image mode draws the IMAGE index on the host (same seed -> same choice on every data-parallel rank) so that a
        # step has no device->host read: only the pixel permutation stays on the device
        self.host_gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)
            self.host_gen.manual_seed(seed)
        self.dev = dev
        self._crop_ids = None
        self.epoch = 0
        if self.use_batching:
            self.n_total = int(self.i_train.numel()) * self.H * self.W
            self._reshuffle()

    # ---- use_batching: permutation of all training pixels, sliding window, reshuffle at epoch end
    def _reshuffle(self):
        self.perm = torch.randperm(self.n_total, device=self.dev, generator=self.gen)
        self.i_batch = 0

    def _next_batching(self):
        idx = self.perm[self.i_batch:self.i_batch + self.N_rand]
        hw = self.H * self.W
        sel = (self.i_train[idx // hw] * hw + idx % hw).contiguous()      # position in i_train -> image id
        ro, rd, tgt = ops.gather_ray_batch(self.H, self.W, self.fx, self.fy, self.cx, self.cy, self.poses, self.images,
                                           sel, img0=-1)
        self.i_batch += self.N_rand
        if self.i_batch >= self.n_total:
            self.epoch += 1
            self._reshuffle()
        return torch.stack([ro, rd], 0), tgt

    # ---- per-image sampling with the optional centre crop
    def _next_per_image(self, i: int):
        k = int(torch.randint(0, len(self.i_train_host), (1,), generator=self.host_gen))      # host draw: no sync
        img = self.i_train_host[k]
        if i < self.precrop_iters:
            if self._crop_ids is None:
                self._crop_ids = precrop_pixel_ids(self.H, self.W, self.precrop_frac).to(self.dev)
            pool = self._crop_ids
            pick = torch.randperm(pool.numel(), device=self.dev, generator=self.gen)[:self.N_rand]
            sel = pool[pick].contiguous()
        else:
            sel = torch.randperm(self.H * self.W, device=self.dev, generator=self.gen)[:self.N_rand].contiguous()
        ro, rd, tgt = ops.gather_ray_batch(self.H, self.W, self.fx, self.fy, self.cx, self.cy, self.poses, self.images,
                                           sel, img0=img)
        return torch.stack([ro, rd], 0), tgt

    def next_batch(self, i: int = 0):
        return self._next_batching() if self.use_batching else self._next_per_image(i)

"""``train(args)`` -- the caller of the hot path in the reference (nerf.ipynb ``train()``, ipynb:993-1452),
reduced to what drives the path: dataset selection and near/far bounds (ipynb:1017-1098), ``create_nerf``,
the ray-batch producer (device-side ``RayBatcher``), render -> MSE(rgb)+MSE(rgb0) -> backward -> Adam ->
exponential lr decay (ipynb:1258-1282), checkpoints in the reference's ``.tar`` layout (ipynb:1290-1299) and the
periodic validation-trajectory and test-set renders with on-GPU metrics (ipynb:1301-1385).  Same ``args`` attribute names as the reference
(missing optional attributes take the reference yaml defaults).  Logging to CSV/JSONL and MP4 writing are out of scope.

Additions, all optional: ``data=`` passes arrays instead of reading ``args.datadir``; ``N_iters=`` replaces the
notebook's global; ``args.rng_seed`` switches perturb / noise draws to the keyed generator (rng.py); under an
initialised ``torch.distributed`` group every step is ``parallel.train_step_sharded`` (ray-sharded batch, one
gradient all-reduce).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import parallel
from . import rng as RNG
from .data import RayBatcher
from .render import create_nerf, img2mse, mse2psnr, render, render_path


def _get(args, name, default):
    return getattr(args, name, default)


def load_dataset(args):
    """(images [N,H,W,3], poses [N,3|4,4], render_poses, hwf, (i_train, i_val, i_test), near, far[, K]) -- ipynb:1017-1098."""
    from . import datasets as D
    if args.dataset_type == "llff":
        images, poses, bds, render_poses, i_test = D.load_llff_data(
            args.datadir, _get(args, "factor", 8), recenter=True, bd_factor=.75, spherify=_get(args, "spherify", False))
        hwf = poses[0, :3, -1]
        poses = poses[:, :3, :4]
        i_test = [i_test] if not isinstance(i_test, list) else i_test
        if _get(args, "llffhold", 0) > 0:
            i_test = np.arange(images.shape[0])[::args.llffhold]
        i_val = i_test
        i_train = np.array([i for i in range(int(images.shape[0])) if i not in i_test and i not in i_val])
        if args.no_ndc:
            near, far = float(np.min(bds)) * .9, float(np.max(bds)) * 1.
        else:
            near, far = 0., 1.
        return images, poses, render_poses, hwf, (i_train, i_val, i_test), near, far
    if args.dataset_type == "blender":
        images, poses, render_poses, hwf, i_split = D.load_blender_data(
            args.datadir, _get(args, "half_res", False), _get(args, "testskip", 8))
        if args.white_bkgd:
            images = images[..., :3] * images[..., -1:] + (1. - images[..., -1:])
        else:
            images = images[..., :3]
        return images, poses, render_poses, hwf, tuple(i_split), 2., 6.
    if args.dataset_type == "LINEMOD":
        images, poses, render_poses, hwf, K, i_split, near, far = D.load_LINEMOD_data(
            args.datadir, _get(args, "half_res", False), _get(args, "testskip", 8))
        if args.white_bkgd and images.shape[-1] == 4:
            images = images[..., :3] * images[..., -1:] + (1. - images[..., -1:])
        else:
            images = images[..., :3]
        return images, poses, render_poses, hwf, tuple(i_split), float(near), float(far), np.array(K, dtype=np.float64)
    if args.dataset_type == "deepvoxels":
        images, poses, render_poses, hwf, i_split = D.load_dv_data(scene=_get(args, "shape", "greek"), basedir=args.datadir,
                                                                   testskip=_get(args, "testskip", 8))
        hemi_R = float(np.mean(np.linalg.norm(poses[:, :3, -1], axis=-1)))       # ipynb:1091-1094
        return images, poses, render_poses, hwf, tuple(i_split), hemi_R - 1., hemi_R + 1.
    raise ValueError(f"Unknown dataset type {args.dataset_type!r} (supported: 'blender', 'llff', 'LINEMOD', 'deepvoxels')")


def save_checkpoint(path, global_step, render_kwargs_train, optimizer):
    """The reference's checkpoint dictionary (ipynb:1290-1299); create_nerf reloads it (ipynb:913-935)."""
    fine = render_kwargs_train.get("network_fine")
    torch.save({"global_step": global_step,
                "network_fn_state_dict": render_kwargs_train["network_fn"].state_dict(),
                "network_fine_state_dict": fine.state_dict() if fine is not None else None,
                "optimizer_state_dict": optimizer.state_dict()}, path)


def train(args, data=None, N_iters=200000 + 1, device=None, log=print):
    """Runs the optimisation loop; returns {'global_step', 'losses', 'psnrs', 'test_metrics', 'render_kwargs_test', ...}."""
    loaded = data if data is not None else load_dataset(args)
    images, poses, render_poses, hwf, (i_train, i_val, i_test), near, far = loaded[:7]
    H, W, focal = int(hwf[0]), int(hwf[1]), float(hwf[2])
    hwf = [H, W, focal]
    # intrinsics: the dataset's own matrix when it has one (LINEMOD), else a centred pinhole (ipynb:1104-1112)
    K = loaded[7] if len(loaded) > 7 and loaded[7] is not None else np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    if _get(args, "render_test", False):
        render_poses = np.array(poses[i_test])
    os.makedirs(os.path.join(args.basedir, args.expname), exist_ok=True)
    items = dict(args) if isinstance(args, dict) else vars(args)       # the reference's AttrDict, or any attribute object
    with open(os.path.join(args.basedir, args.expname, "args.txt"), "w") as f:
        for k in sorted(items):
            f.write(f"{k} = {items[k]}\n")
    try:
        from .config import save_config
        save_config(items, os.path.join(args.basedir, args.expname, "config.yaml"))
    except Exception as e:        # non-YAML-able values in a hand-built namespace: the text dump above is enough
        log(f"config.yaml not written: {e}")

    render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer = create_nerf(args, device=device)
    dev = next(render_kwargs_train["network_fn"].parameters()).device
    if dist.is_initialized() and dist.get_world_size() > 1:
        # data parallel all-reduces gradients only: make every replica (weights, Adam state, lr, resume step) rank 0's,
        # whatever each rank's RNG state was before create_nerf and whichever rank could see the checkpoint
        parallel.broadcast_parameters(grad_vars, optimizer)
        step0 = torch.tensor([start], dtype=torch.int64, device=dev)
        dist.broadcast(step0, src=0)
        start = int(step0.item())
    for kw in (render_kwargs_train, render_kwargs_test):
        kw.update(near=near, far=far)
    global_step = start
    chunk = _get(args, "chunk", 1024 * 32)
    render_poses = torch.as_tensor(np.asarray(render_poses), dtype=torch.float32).to(dev)
    images = np.asarray(images, dtype=np.float32)
    poses_t = torch.as_tensor(np.asarray(poses), dtype=torch.float32).to(dev)

    is_rank0 = not dist.is_initialized() or dist.get_rank() == 0

    def testset(i, tag):
        """Test-set render (ipynb:1312-1385).  Under torch.distributed every frame is ray-sharded over ALL ranks
        (render_path(sharded=...) default); rank 0 alone writes the frames and metrics.json."""
        savedir = os.path.join(args.basedir, args.expname, f"{tag}_{i:06d}")
        if is_rank0:
            os.makedirs(savedir, exist_ok=True)
        with torch.no_grad():
            out = render_path(poses_t[list(i_test)], hwf, K, chunk, render_kwargs_test, gt_imgs=images[list(i_test)],
                              savedir=savedir if is_rank0 else None, calculate_metrics=True, metrics_include_lpips=False,
                              render_factor=_get(args, "render_factor", 0))
        metrics = out[2] if len(out) > 2 else {}
        if is_rank0:
            with open(os.path.join(savedir, "metrics.json"), "w") as f:
                json.dump({"iteration": i, "metrics": metrics}, f, indent=2)
        return metrics

    if _get(args, "render_only", False):                       # ipynb:1142-1165
        metrics = testset(start, "renderonly_test") if _get(args, "render_test", False) else {}
        if not _get(args, "render_test", False):
            with torch.no_grad():
                render_path(render_poses, hwf, K, chunk, render_kwargs_test, render_factor=_get(args, "render_factor", 0))
        return {"global_step": start, "losses": [], "psnrs": [], "test_metrics": {start: metrics},
                "render_kwargs_test": render_kwargs_test}

    world = dist.get_world_size() if dist.is_initialized() else 1
    # data-parallel ranks must draw the SAME global batch (each then takes its slice): common seed required
    batch_seed = _get(args, "rng_seed", None) if world == 1 else int(_get(args, "rng_seed", None) or 0)
    batcher = RayBatcher(images, poses_t[:, :3, :4], H, W, K, i_train=list(i_train), N_rand=args.N_rand,
                         use_batching=not _get(args, "no_batching", False),
                         precrop_iters=_get(args, "precrop_iters", 0), precrop_frac=_get(args, "precrop_frac", .5),
                         device=dev, seed=batch_seed)
    bucket = parallel.GradBucket(grad_vars) if world > 1 else None
    seed = _get(args, "rng_seed", None)
    losses, psnrs, test_metrics = [], [], {}
    i_print, i_weights, i_testset = _get(args, "i_print", 100), _get(args, "i_weights", 10000), _get(args, "i_testset", 50000)
    i_video = _get(args, "i_video", 50000)
    t0 = time.time()
    for i in range(start + 1, N_iters):
        global_step = i
        batch_rays, target_s = batcher.next_batch(i)
        step_seed = RNG.step_seed(seed, i) if seed is not None else None
        if world > 1:
            loss = parallel.train_step_sharded(render, img2mse, optimizer, bucket, H, W, K, batch_rays[0], batch_rays[1],
                                               target_s, rng_seed=step_seed, chunk=chunk, retraw=True,
                                               **render_kwargs_train)
            img_loss = loss
        else:
            ctx = RNG.keyed_rng(step_seed) if step_seed is not None else torch.enable_grad()
            with ctx:
                rgb, disp, acc, extras = render(H, W, K, chunk=chunk, rays=batch_rays, verbose=i < 10, retraw=True,
                                                **render_kwargs_train)
            optimizer.zero_grad()
            img_loss = img2mse(rgb, target_s)
            loss = img_loss
            if "rgb0" in extras:
                loss = loss + img2mse(extras["rgb0"], target_s)
            loss.backward()
            optimizer.step()
        # exponential decay, applied after the step (ipynb:1278-1282)
        new_lrate = args.lrate * (0.1 ** (global_step / (args.lrate_decay * 1000)))
        for g in optimizer.param_groups:
            g["lr"] = new_lrate
        if i % i_weights == 0 and (not dist.is_initialized() or dist.get_rank() == 0):
            path = os.path.join(args.basedir, args.expname, "checkpoints", f"{i:06d}.tar")
            save_checkpoint(path, global_step, render_kwargs_train, optimizer)
            log("Saved checkpoints at", path)
        if i % i_video == 0 and i > 0:
            # validation trajectory (ipynb:1301-1310), ray-sharded over all ranks; rank 0 keeps the frames as arrays (MP4
            # writing needs imageio-ffmpeg)
            with torch.no_grad():
                rgbs, disps = render_path(render_poses, hwf, K, chunk, render_kwargs_test)[:2]
            if is_rank0:
                base = os.path.join(args.basedir, args.expname, "{}_spiral_{:06d}_".format(args.expname, i))
                np.save(base + "rgb.npy", (255 * np.clip(rgbs, 0, 1)).astype(np.uint8))
                np.save(base + "disp.npy", (255 * np.clip(disps / max(float(np.max(disps)), 1e-10), 0, 1)).astype(np.uint8))
                log("Done, saving", rgbs.shape, disps.shape)
        if i % i_testset == 0 and i > 0 and len(i_test) > 0:
            test_metrics[i] = testset(i, "testset")
            log(f"[TEST] Iter: {i} " + " ".join(f"{k}: {v:.4f}" for k, v in test_metrics[i].items()))
        if i % i_print == 0:                                  # the only host sync of the loop
            losses.append((i, float(loss.detach())))
            psnrs.append((i, float(mse2psnr(img_loss.detach()))))
            log(f"[TRAIN] Iter: {i} Loss: {losses[-1][1]}  PSNR: {psnrs[-1][1]}  ({(time.time() - t0) / (i - start):.4f} s/it)")
    return {"global_step": global_step, "losses": losses, "psnrs": psnrs, "test_metrics": test_metrics,
            "render_kwargs_train": render_kwargs_train, "render_kwargs_test": render_kwargs_test,
            "optimizer": optimizer}

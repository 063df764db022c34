"""Multi-GPU parity on hardware (NCCL, one process per GPU; skipped with fewer than 2 GPUs -- run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):

* a frame rendered ray-sharded over 2 ranks and gathered is BIT-equal to the single-GPU render of the same rays
  (BASELINE configs[3] shape: LLFF / NDC, and configs[1]);
* render_path(sharded) returns the same frames on every rank as the single-GPU render_path;
* the data-parallel training step on a 32 768-ray batch (configs[4]) produces the gradients of the single-GPU step on
  the whole batch (fp32 accumulation-order tolerance) with keyed perturb / noise draws.
"""
import os
import socket
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _args():
    import tempfile
    return types.SimpleNamespace(multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=128, netdepth=8,
                                 netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4,
                                 basedir=tempfile.mkdtemp(prefix="nerfb_multi_"), expname="m", ft_path=None, no_reload=True,
                                 perturb=0.0, N_samples=64, white_bkgd=True, raw_noise_std=0.0, dataset_type="blender",
                                 no_ndc=False, lindisp=False)


def _worker(rank, world, port, out):
    import contextlib, io
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import nerf_projects_b200 as nb
        from nerf_projects_b200 import parallel
        from nerf_projects_b200 import synthetic as SY
        res = {}
        torch.manual_seed(100 + rank)             # replicas start DIFFERENT on purpose
        with contextlib.redirect_stdout(io.StringIO()):
            kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(_args(), device=dev)
        parallel.broadcast_parameters(grad_vars, optim)
        ref = [p.detach().clone() for p in grad_vars]
        for p in ref:
            dist.broadcast(p, src=0)
        res["replicas_synchronised"] = all(torch.equal(a, b) for a, b in zip(ref, grad_vars))
        # ---- render: sharded == single, bitwise (C1-shaped subset and C3-shaped NDC subset, ragged ray counts)
        ok = True
        for cfg in ("C1", "C3"):
            if cfg == "C1":
                H, W = 800, 800; K, c2w = SY.blender_shaped_camera(H, W, 30.0); near, far = 2., 6.
                kw = dict(kw_test); kw.update(ndc=False, white_bkgd=True)
            else:
                H, W = 756, 1008; K, c2w = SY.llff_shaped_camera(H, W); near, far = 0., 1.
                kw = dict(kw_test); kw.update(ndc=True, white_bkgd=False, lindisp=False)
            ro, rd = nb.get_rays(H, W, K, torch.from_numpy(c2w).to(dev))
            ro, rd = ro.reshape(-1, 3)[:70001].contiguous(), rd.reshape(-1, 3)[:70001].contiguous()
            with torch.no_grad():
                tiles = parallel.render_sharded(nb.render, H, W, K, ro, rd, chunk=32768, near=near, far=far, return_packed=True,
                                                **kw).clone()
                rgb, disp, acc, _ = nb.render(H, W, K, chunk=32768, rays=(ro, rd), near=near, far=far, **kw)
            ok &= torch.equal(tiles[:, 0:3], rgb) and torch.equal(tiles[:, 3], disp) and torch.equal(tiles[:, 4], acc)
        res["render_sharded_equals_single"] = bool(ok)
        # ---- render_path: sharded frames == per-frame single-GPU frames
        Hs = Ws = 96
        Ks, _ = SY.blender_shaped_camera(Hs, Ws, 0.0)
        poses = [torch.from_numpy(SY.blender_shaped_camera(Hs, Ws, 40.0 * k)[1]).to(dev) for k in range(3)]
        kwp = dict(kw_test); kwp.update(ndc=False, near=2., far=6.)
        with torch.no_grad():
            rgbs, disps = nb.render_path(poses, (Hs, Ws, Ks[0][0]), Ks, 4096, kwp)[:2]
            rgbs1, disps1 = nb.render_path(poses, (Hs, Ws, Ks[0][0]), Ks, 4096, kwp, sharded=False)[:2]
        res["render_path_sharded_equals_single"] = bool(np.array_equal(rgbs, rgbs1) and np.array_equal(disps, disps1))
        # ---- C4: data-parallel step on a 32 768-ray batch == single-GPU step on the whole batch
        H, W = 800, 800
        K, c2w = SY.blender_shaped_camera(H, W, 30.0)
        ro, rd = nb.get_rays(H, W, K, torch.from_numpy(c2w).to(dev))
        gen = torch.Generator(device=dev).manual_seed(0)
        sel = torch.randint(0, H * W, (32768,), device=dev, generator=gen)
        b_o, b_d = ro.reshape(-1, 3)[sel].contiguous(), rd.reshape(-1, 3)[sel].contiguous()
        target = torch.rand((32768, 3), device=dev, generator=gen)
        kw_tr = dict(kw_train); kw_tr.update(ndc=False, lindisp=False, perturb=1.0, raw_noise_std=1.0, near=2., far=6.)
        bucket = parallel.GradBucket(grad_vars)
        before = [p.detach().clone() for p in grad_vars]
        seed = nb.step_seed(3, 1)
        loss = parallel.train_step_sharded(nb.render, nb.img2mse, optim, bucket, H, W, K, b_o, b_d, target, rng_seed=seed,
                                           chunk=32768, **kw_tr)
        g_sharded = bucket.flat.clone()
        after = [p.detach().clone() for p in grad_vars]
        with torch.no_grad():
            for p, b in zip(grad_vars, before):
                p.copy_(b)
        optim.zero_grad()
        with nb.keyed_rng(seed, ray_offset=0):
            rgb, _, _, ex = nb.render(H, W, K, rays=(b_o, b_d), chunk=32768, **kw_tr)
        loss1 = nb.img2mse(rgb, target) + nb.img2mse(ex["rgb0"], target)
        loss1.backward()
        g_single = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in grad_vars])
        res["grad_rel_err"] = float((g_sharded - g_single).norm() / g_single.norm())
        res["loss_abs_err"] = abs(float(loss) - float(loss1))
        res["params_moved"] = any(not torch.equal(a, b) for a, b in zip(after, before))
        gathered = [torch.empty_like(after[0]) for _ in range(world)]
        dist.all_gather(gathered, after[0])
        res["replicas_equal_after_step"] = all(torch.equal(gathered[0], t) for t in gathered)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_sharded_render_and_training_equal_single_gpu_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in (0, 1):
        r = out[rank]
        print(rank, r)
        assert r["replicas_synchronised"] and r["render_sharded_equals_single"] and r["render_path_sharded_equals_single"]
        assert r["grad_rel_err"] < 1e-4 and r["loss_abs_err"] < 1e-6, r
        assert r["params_moved"] and r["replicas_equal_after_step"]

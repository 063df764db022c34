"""Host-side logic of the multi-GPU path with world_size-2 gloo on CPU: ray
sharding bounds, the tile all-gather of render_sharded (with a stand-in
render function -- the kernels need a GPU) and the gradient bucket
all-reduce.  The CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_bounds_cover_all_rays():
    from nerf_projects_b200.parallel import shard_bounds
    for n in (0, 1, 7, 640000, 762048, 32768):
        for world in (1, 2, 3, 4, 8):
            seen = []
            per = None
            for r in range(world):
                lo, hi, per_r = shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per_r
                per = per_r
                seen += list(range(lo, hi))[:3] + list(range(max(lo, hi - 3), hi))
            total = sum(shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world))
            assert total == n and per * world >= n


def _fake_render(H, W, K, rays=None, **kw):
    ro, rd = rays
    rgb = ro * 2.0 + rd
    disp = ro[:, 0] + 1.0
    acc = rd[:, 2] * 3.0
    return [rgb, disp, acc, {}]


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nerf_projects_b200 import parallel
        g = torch.Generator().manual_seed(0)
        ro, rd = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
        rgb, disp, acc = parallel.render_sharded(_fake_render, 4, 4, None, ro, rd)
        ok_render = torch.equal(rgb, ro * 2.0 + rd) and torch.equal(disp, ro[:, 0] + 1.0) and torch.equal(acc, rd[:, 2] * 3)
        # gradient bucket: rank r holds grad = (r+1) * ones -> mean over ranks
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
        for p in params[:2]:
            p.grad = torch.full_like(p, float(rank + 1))
        bucket = parallel.GradBucket(params)          # third parameter has no grad -> treated as zero
        bucket.allreduce_mean()
        mean = sum(range(1, world + 1)) / world
        ok_grad = all(torch.allclose(p.grad, torch.full_like(p, mean)) for p in params[:2]) and \
            torch.equal(params[2].grad, torch.zeros(2, 2))
        out[rank] = int(ok_render) + 2 * int(ok_grad)
    finally:
        dist.destroy_process_group()


def _fake_render_keyed(H, W, K, rays=None, **kw):
    """Stand-in render that reports the global ray index its keyed_rng context gives every ray."""
    from nerf_projects_b200 import rng
    ro, rd = rays
    seed, off = rng.current()
    idx = off + torch.arange(ro.shape[0], dtype=torch.float32)
    return [ro * 0 + float(seed), idx, idx * 2, {}]


def _worker_keyed(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nerf_projects_b200 import parallel, rng
        ro = torch.zeros(n, 3)
        rgb, disp, acc = parallel.render_sharded(_fake_render_keyed, 4, 4, None, ro, ro, rng_seed=77)
        ok = torch.equal(disp, torch.arange(n, dtype=torch.float32)) and bool((rgb == 77).all()) and rng.current() is None
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def test_render_sharded_keys_draws_by_global_ray_index_gloo_world2():
    """each rank's keyed_rng context starts at its first ray."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_keyed, args=(2, port, 33, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_keyed_rng_context_nesting_and_chunk_advance():
    from nerf_projects_b200 import rng
    assert rng.current() is None
    with rng.keyed_rng(5, ray_offset=100):
        assert rng.current() == (5, 100)
        with rng.advanced(32):
            assert rng.current() == (5, 132)
        assert rng.current() == (5, 100)
    assert rng.current() is None
    with rng.advanced(3):                    # no active context: stays off
        assert rng.current() is None
    assert rng.step_seed(1, 0) != rng.step_seed(1, 1) and 0 <= rng.step_seed(1, 0) < 2 ** 64


@pytest.mark.parametrize("n", [10, 33])
def test_render_sharded_and_grad_bucket_gloo_world2(n):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    assert dict(out) == {0: 3, 1: 3}


def test_shard_bounds_alignment():
    """Shards are whole multiples of SHARD_ALIGN rays (whole super-tiles of the fused kernel); only the tail is ragged."""
    from nerf_projects_b200.parallel import SHARD_ALIGN, shard_bounds
    for n in (640000, 762048, 32768, 4097, 5):
        for world in (2, 3, 8):
            spans = [shard_bounds(n, world, r, SHARD_ALIGN) for r in range(world)]
            assert all(per % SHARD_ALIGN == 0 for _, _, per in spans)
            assert sum(hi - lo for lo, hi, _ in spans) == n
            assert all((hi - lo) % SHARD_ALIGN == 0 for lo, hi, _ in spans if hi < n)


class _TinyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.lin = torch.nn.Linear(6, 3)


def _model_render(model):
    def render(H, W, K, rays=None, **kw):
        ro, rd = rays
        rgb = torch.tanh(model.lin(torch.cat([ro, rd], -1)))
        return [rgb, rgb[:, 0], rgb[:, 1], {"rgb0": rgb * 0.5}]
    return render


def _worker_train(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nerf_projects_b200 import parallel
        mse = lambda x, y: torch.mean((x - y) ** 2)
        g = torch.Generator().manual_seed(1)
        ro, rd, tgt = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g)
        # replicas that start DIFFERENT (a per-rank seed): broadcast_parameters makes them rank 0's
        model = _TinyModel()
        with torch.no_grad():
            model.lin.weight.add_(float(rank))
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        parallel.broadcast_parameters(list(model.parameters()), opt)
        ref = _TinyModel()
        ok_bcast = all(torch.equal(a, b) for a, b in zip(model.parameters(), ref.parameters()))
        # single-process reference step on the GLOBAL batch
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
        r_rgb = _model_render(ref)(0, 0, None, rays=(ro, rd))
        r_loss = mse(r_rgb[0], tgt) + mse(r_rgb[3]["rgb0"], tgt)
        ropt.zero_grad(); r_loss.backward(); ropt.step()
        bucket = parallel.GradBucket(list(model.parameters()))
        loss = parallel.train_step_sharded(_model_render(model), mse, opt, bucket, 0, 0, None, ro, rd, tgt)
        ok_loss = abs(float(loss) - float(r_loss)) < 1e-6
        ok_step = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(model.parameters(), ref.parameters()))
        out[rank] = int(ok_bcast) + 2 * int(ok_loss) + 4 * int(ok_step)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 9, 16, 37])
def test_train_step_sharded_equals_global_batch_step_gloo_world2(n):
    """the data-parallel step must equal the single-process step on the global batch for uneven shards
    (n % world != 0; 9 and 37 rays also leave the aligned split ragged) and when a rank gets NO rays (n = 1: the short
    last batch of an epoch) -- that rank joins the all-reduce with zero gradients instead of raising / hanging.
    Replicas that start different are synchronised by broadcast_parameters."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_train, args=(2, port, n, out), nprocs=2, join=True)
    assert dict(out) == {0: 7, 1: 7}


def _worker_async(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nerf_projects_b200 import parallel
        g = torch.Generator().manual_seed(0)
        frames = [(torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)) for _ in range(3)]
        handles = [parallel.render_sharded(_fake_render, 4, 4, None, ro, rd, async_gather=True, slot=i % 2)
                   for i, (ro, rd) in enumerate(frames[:2])]
        ok = True
        for (ro, rd), h in zip(frames[:2], handles):          # two gathers in flight, two buffer slots
            rgb, disp, acc = h.maps()
            ok &= torch.equal(rgb, ro * 2.0 + rd) and torch.equal(disp, ro[:, 0] + 1.0) and torch.equal(acc, rd[:, 2] * 3)
        packed = parallel.render_sharded(_fake_render, 4, 4, None, *frames[2], return_packed=True)
        ok &= packed.shape == (n, 5) and torch.equal(packed[:, 0:3], frames[2][0] * 2.0 + frames[2][1])
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def test_render_sharded_async_gather_gloo_world2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_async, args=(2, port, 21, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}

import sys, torch, time
sys.path.insert(0, "/root/repo")
import nerf_projects_b200 as nb
from nerf_projects_b200 import ops, parallel, mlp
from nerf_projects_b200 import synthetic as SY
import bench
dev = torch.device("cuda")
kw_train, kw_test, start, gv, optim = nb.create_nerf(bench.make_args(), device=dev)
K, c2w = SY.blender_shaped_camera(800, 800, 30.0)
ro, rd = nb.get_rays(800, 800, K, torch.from_numpy(c2w).cuda())
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
kw = dict(kw_train); kw.update(ndc=False, lindisp=False, perturb=1.0, raw_noise_std=1.0)
gen = torch.Generator(device=dev).manual_seed(0)
sel = torch.randint(0, 640000, (4096,), device=dev, generator=gen)
bo, bd = ro[sel].contiguous(), rd[sel].contiguous()
target = torch.rand((4096, 3), device=dev, generator=gen)
bucket = parallel.GradBucket(gv)
cnt = [0]
def step():
    cnt[0] += 1
    parallel.train_step_sharded(nb.render, nb.img2mse, optim, bucket, 800, 800, K, bo, bd, target,
                                rng_seed=nb.step_seed(0, cnt[0]), chunk=32768, near=2., far=6., **kw)
for _ in range(3): step()
torch.cuda.synchronize()
for rep in range(3):
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print(f"train step: {e0.elapsed_time(e1)/10:.3f} ms (wall {(time.perf_counter()-t0)*100:.3f} ms), launches/step {(ops.launches()-l0)/10}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

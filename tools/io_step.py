"""The ray / compositing I/O kernels of one 32768-ray chunk (camera -> rays -> packed rays -> depths -> [MLP] -> compositing ->
hierarchical depths -> [MLP] -> compositing), one 32768-ray training step (adds the compositing backward and the keyed draws) and
the image metrics of one frame: the command behind profiles/r01_ncu_full_io_kernels.txt
(ncu --set full -k regex:"raw2outputs|hierarchical|coarse_z|pack_rays|get_rays|keyed_rng|ssim|gather_ray|adam" ...)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_projects_b200 as nb  # noqa: E402
from nerf_projects_b200 import synthetic as SY  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda")
kw_train, kw_test, start, gv, optim = nb.create_nerf(bench.make_args(), device=dev)
kw_test["network_fn"].load_state_dict(SY.random_init_state_dict(0, use_viewdirs=True, output_ch=5))
kw_test["network_fine"].load_state_dict(SY.random_init_state_dict(1, use_viewdirs=True, output_ch=5))
H, W = 128, 256                                   # 32768 rays = one chunk
K, c2w = SY.blender_shaped_camera(H, W, 30.0)
pose = torch.from_numpy(c2w).to(dev)
with torch.no_grad():
    for _ in range(2):
        rgb, disp, acc, _ = nb.render(H, W, K, chunk=32768, c2w=pose[:3, :4], near=2., far=6., **kw_test)
# training step on the same 32768 rays through the device ray batcher (one image)
images = rgb.detach().reshape(1, H, W, 3).clamp(0, 1)
batcher = nb.RayBatcher(images, pose[None, :3, :4], H, W, K, N_rand=32768, use_batching=True, device=dev, seed=0)
kw = dict(kw_train); kw.update(ndc=False, lindisp=False, perturb=1.0, raw_noise_std=1.0)
for i in range(2):
    batch_rays, target = batcher.next_batch(i)
    with nb.keyed_rng(nb.step_seed(0, i)):
        out = nb.render(H, W, K, chunk=32768, rays=batch_rays, near=2., far=6., **kw)
    loss = nb.img2mse(out[0], target) + nb.img2mse(out[3]["rgb0"], target)
    optim.zero_grad(); loss.backward(); optim.step()
a, b = torch.rand(800, 800, 3, device=dev), torch.rand(800, 800, 3, device=dev)
print(nb.calculate_metrics(a, b, include_lpips=False), float(loss.detach()))
torch.cuda.synchronize()

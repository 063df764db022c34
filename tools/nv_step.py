"""One training step of the network WITHOUT view directions (64 + 128 samples, 4096 rays, keyed perturb / noise draws) followed by
the fused image metrics of an 800x800 frame -- the command behind profiles/r01_launches_v5_nv_train_step.csv (ncu launch list)."""
import os
import sys
import tempfile
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_projects_b200 as nb  # noqa: E402

dev = "cuda"
args = types.SimpleNamespace(multires=10, i_embed=0, use_viewdirs=False, multires_views=4, N_importance=128, netdepth=8,
                             netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4,
                             basedir=tempfile.mkdtemp(prefix="nerfb_"), expname="t", ft_path=None, no_reload=True, perturb=1.0,
                             N_samples=64, white_bkgd=True, raw_noise_std=1.0, dataset_type="blender", no_ndc=False, lindisp=False)
torch.manual_seed(0)
kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(args)
R = 4096
gen = torch.Generator().manual_seed(0)
ro = (torch.randn(R, 3, generator=gen) * 0.1 + torch.tensor([0., 0., 4.])).to(dev)
rd = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen) * 0.2 + torch.tensor([0., 0., -1.]), dim=-1).to(dev)
target = torch.rand(R, 3, generator=gen).to(dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    with nb.keyed_rng(nb.step_seed(1, i)):
        rgb, disp, acc, extras = nb.render(800, 800, [[1111.0, 0, 400.0], [0, 1111.0, 400.0], [0, 0, 1]], rays=(ro, rd),
                                           near=2., far=6., retraw=True, **kw_train)
    loss = nb.img2mse(rgb, target) + nb.img2mse(extras["rgb0"], target)
    optim.zero_grad(); loss.backward(); optim.step()
    ev[i + 1].record()
a, b = torch.rand(800, 800, 3, device=dev), torch.rand(800, 800, 3, device=dev)
m = nb.calculate_metrics(a, b, include_lpips=False)
torch.cuda.synchronize()
print("step ms:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)], "loss", float(loss.detach()), "metrics", m)

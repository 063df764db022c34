"""Diagnostic (GPU box only): the end-to-end training loop of tests/test_gpu_render.py::test_end_to_end_training_converges_bf16
as a script, optionally after polluting the caching allocator with NaN / huge values, to tell numerical chaos from reads of
uninitialised memory.  Lives under tests/ (not collected by pytest) because it builds its teacher scene with the oracle.
    INIT_SEED=k python tests/diag_converge.py [none|nan|big] [steps] [seed]"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))    # _cases
import nerf_projects_b200 as nb  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402
from _cases import T, case_params  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "none"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
DEV = "cuda"


def pollute():
    if mode == "none":
        return
    bufs = []
    for _ in range(24):
        b = torch.empty(1 << 28, dtype=torch.float32, device=DEV)      # 1 GiB each
        b.fill_(float("nan") if mode == "nan" else 3.0e38)
        bufs.append(b)
    del bufs                                                         # back to the caching allocator, contents stay


def make_args(**over):
    a = dict(multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=128, netdepth=8, netwidth=256,
             netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4, basedir=tempfile.mkdtemp(prefix="nerfb_"),
             expname="t", ft_path=None, no_reload=True, perturb=0.0, N_samples=64, white_bkgd=False, raw_noise_std=0.0,
             dataset_type="blender", no_ndc=False, lindisp=False)
    a.update(over)
    return types.SimpleNamespace(**a)


H = W = 48
K, _ = O.blender_camera(H, W, 0.0)
kw_train, teacher_kw, _, _, _ = nb.create_nerf(make_args())
pc, pf = case_params("render_c1_sharp")
teacher_kw["network_fn"].load_state_dict(O.to_torch(pc))
teacher_kw["network_fine"].load_state_dict(O.to_torch(pf))
tk = dict(teacher_kw); tk.update(ndc=False, near=2., far=6., white_bkgd=True, lindisp=False)
poses = torch.stack([T(O.blender_camera(H, W, 45.0 * k)[1]).float() for k in range(8)], 0)
with torch.no_grad():
    images = torch.stack([nb.render(H, W, K, chunk=4096, c2w=p[:3, :4].to(DEV), **tk)[0] for p in poses], 0)
pollute()
torch.manual_seed(int(os.environ.get("INIT_SEED", "0")))        # student initialisation (nn.Linear default init, CPU generator)
kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(make_args())
kw = dict(kw_train); kw.update(ndc=False, lindisp=False, white_bkgd=True, perturb=1.0, raw_noise_std=0.0)
batcher = nb.RayBatcher(images, poses[:, :3, :4], H, W, K, N_rand=1024, use_batching=True, device=DEV, seed=seed)
torch.manual_seed(seed)
losses = []
cs = lambda t: f"{float(t.double().sum()):.10g}"
print("checksums: images", cs(images), "params", cs(torch.cat([p.flatten() for p in grad_vars])), "perm", cs(batcher.perm[:4096]))
for i in range(steps):
    batch_rays, target = batcher.next_batch(i)
    if i == 0:
        print("batch0: rays", cs(batch_rays), "target", cs(target), "cuda rng probe", cs(torch.rand(4, device=DEV)))
        torch.manual_seed(seed)
    rgb, disp, acc, extras = nb.render(H, W, K, chunk=32768, rays=batch_rays, near=2., far=6., retraw=True, **kw)
    loss = nb.img2mse(rgb, target) + nb.img2mse(extras["rgb0"], target)
    optim.zero_grad(); loss.backward(); optim.step()
    losses.append(float(loss.detach()))
print(f"variant={os.environ.get('NERFB_VARIANT', '')!r} init={os.environ.get('INIT_SEED', '0')} pollute={mode} seed={seed}: loss {np.mean(losses[:10]):.4f} -> "
      f"{np.mean(losses[-10:]):.4f}; every 50: {[round(l, 4) for l in losses[::50]]}; finite={np.isfinite(losses).all()}")

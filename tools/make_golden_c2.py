#!/usr/bin/env python
"""BASELINE.json configs[2] (C2) pinned by the reference: ONE training objective on a 4096-ray batch, coarse 64 + fine
128, MSE(rgb) + MSE(rgb0), autograd gradients -- computed by the UNMODIFIED reference on CPU (build container only;
same harness as tools/make_golden.py, loop body of ipynb:1258-1274).

    python tools/make_golden_c2.py       # writes tests/golden/train_grads_c2.npz  (~25 GB of autograd state, ~1 min)

Networks: the calibrated synthetic scene of the ``render_c1_sharp`` fixture (its density heads are read back from that
fixture).  Stored: rays, targets, loss, the rendered maps of the batch, and per parameter tensor the gradient's norm,
sum and either all entries (small tensors) or 2048 sampled entries.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from make_golden import load_reference, ref_models, save, subset_rays  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402  (input generation only)
from _cases import case_params  # noqa: E402

N_RAYS = 4096


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    R = load_reference()
    K1, c1 = O.blender_camera(800, 800, 30.0)
    ro, rd, idx = subset_rays(800, 800, K1, c1, N_RAYS, 33)
    rng = np.random.Generator(np.random.PCG64(34))
    target = torch.from_numpy(rng.uniform(0, 1, size=(N_RAYS, 3)).astype(np.float32))
    pcs, pfs = case_params("render_c1_sharp")
    kw_train, kw_test, grad_vars = ref_models(R, pcs, pfs, True, 128)
    kw = dict(kw_test); kw.update(white_bkgd=True, ndc=False, lindisp=False)
    rgb, disp, acc, extras = R.render(800, 800, K1, chunk=32768, rays=(ro, rd), near=2.0, far=6.0, retraw=True, **kw)
    loss = R.ref_helpers.img2mse(rgb, target) + R.ref_helpers.img2mse(extras["rgb0"], target)
    loss.backward()
    g = dict(rays_o=ro, rays_d=rd, idx=idx, target=target, loss=loss.detach(), rgb_map=rgb.detach(), acc_map=acc.detach(),
             disp_map=disp.detach(), rgb0=extras["rgb0"].detach(), acc0=extras["acc0"].detach(),
             sigma_last=extras["raw"][:, -1, 3].detach())
    srng = np.random.Generator(np.random.PCG64(35))
    for tag, net in (("c", kw["network_fn"]), ("f", kw["network_fine"])):
        for name, prm in net.named_parameters():
            gr = prm.grad if prm.grad is not None else torch.zeros_like(prm)
            flat = gr.reshape(-1)
            g[f"{tag}.{name}.norm"] = flat.double().norm().float()
            g[f"{tag}.{name}.sum"] = flat.double().sum().float()
            if flat.numel() <= 4096:
                g[f"{tag}.{name}.full"] = gr
            else:
                sel = np.sort(srng.choice(flat.numel(), size=2048, replace=False))
                g[f"{tag}.{name}.idx"] = sel.astype(np.int32)
                g[f"{tag}.{name}.val"] = flat[torch.from_numpy(sel)]
    save("train_grads_c2", **g)


if __name__ == "__main__":
    main()

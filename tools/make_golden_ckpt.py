#!/usr/bin/env python
"""Generate the checkpoint / Keras-weights fixtures by executing the UNMODIFIED reference (build container only).

    python tools/make_golden_ckpt.py      # writes tests/golden/ref_ckpt_w64.tar, ref_ckpt_w64.npz, keras_weights.npz

1. ``ref_ckpt_w64.tar`` -- a checkpoint WRITTEN BY THE REFERENCE: its ``create_nerf`` (ipynb:876-961) builds the two
   networks and ``torch.optim.Adam``; two iterations of the loop body of ``train()`` (ipynb:1258-1282: render, MSE(rgb) +
   MSE(rgb0), backward, step, lr decay) run on CPU; then the dictionary of ipynb:1290-1299 is ``torch.save``d.  The
   networks are 64 wide (netwidth=64) so the file stays ~1 MB -- module names, state_dict layout, optimizer state layout
   and the reload code path are independent of the width.  ``ref_ckpt_w64.npz`` holds the inputs (rays, targets, initial
   parameters' seeds) and what the reference produces when it RESUMES from that file and takes one more step: loss and
   every parameter after step 3.
2. ``keras_weights.npz`` -- ``NeRF.load_weights_from_keras`` (nerf.py:113-146) of the reference applied to a seeded
   Keras-layout weight list: per-tensor digests of the resulting state_dict and the network's outputs on 32 inputs.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from make_golden import load_reference, save, subset_rays  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402  (input generation only)

GOLD = os.path.join(ROOT, "tests", "golden")
W64 = dict(netdepth=8, netwidth=64, netdepth_fine=8, netwidth_fine=64)
LRATE, LRATE_DECAY = 5e-4, 250


def ckpt_args(basedir, no_reload):
    return types.SimpleNamespace(multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=128,
                                 netchunk=65536, lrate=LRATE, basedir=basedir, expname="ck", ft_path=None,
                                 no_reload=no_reload, perturb=0.0, N_samples=64, white_bkgd=True, raw_noise_std=0.0,
                                 dataset_type="blender", no_ndc=False, lindisp=False, **W64)


def ckpt_init_params(seed):
    """64-wide networks, numpy PCG64 init (torch-version independent), density head scaled up so the rays are not empty"""
    p = O.make_params(seed, W=64, use_viewdirs=True, output_ch=5)
    p["alpha_linear.weight"] = (p["alpha_linear.weight"] * 8.0).astype(np.float32)
    p["alpha_linear.bias"] = (p["alpha_linear.bias"] + 0.3).astype(np.float32)
    return p


def train_iteration(R, K, rays, target, kw_train, optimizer, global_step):
    """the loop body of the reference's train() (ipynb:1258-1282), its own functions"""
    rgb, disp, acc, extras = R.render(800, 800, K, chunk=32768, rays=rays, verbose=False, retraw=True, **kw_train)
    optimizer.zero_grad()
    img_loss = R.ref_helpers.img2mse(rgb, target)
    loss = img_loss
    if "rgb0" in extras:
        loss = loss + R.ref_helpers.img2mse(extras["rgb0"], target)
    loss.backward()
    optimizer.step()
    new_lrate = LRATE * (0.1 ** (global_step / (LRATE_DECAY * 1000)))
    for g in optimizer.param_groups:
        g["lr"] = new_lrate
    return float(loss.detach())


def make_checkpoint(R):
    K1, c1 = O.blender_camera(800, 800, 30.0)
    ro, rd, _ = subset_rays(800, 800, K1, c1, 48, 21)
    rng = np.random.Generator(np.random.PCG64(77))
    target = torch.from_numpy(rng.uniform(0, 1, size=(48, 3)).astype(np.float32))
    base = tempfile.mkdtemp(prefix="golden_ckpt_")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            kw_train, kw_test, start, grad_vars, optimizer = R.create_nerf(ckpt_args(base, True))
        kw_train["network_fn"].load_state_dict(O.to_torch(ckpt_init_params(31)))
        kw_train["network_fine"].load_state_dict(O.to_torch(ckpt_init_params(32)))
        kw_train.update(near=2.0, far=6.0)
        losses = []
        for i in (1, 2):
            losses.append(train_iteration(R, K1, (ro, rd), target, kw_train, optimizer, i))
        path = os.path.join(base, "ck", "checkpoints", "{:06d}.tar".format(2))
        torch.save({                                                    # ipynb:1292-1298, verbatim keys
            'global_step': 2,
            'network_fn_state_dict': kw_train['network_fn'].state_dict(),
            'network_fine_state_dict': kw_train['network_fine'].state_dict(),
            'optimizer_state_dict': optimizer.state_dict(),
        }, path)
        # the reference RESUMES from its own file (create_nerf with no_reload=False) and takes step 3
        with contextlib.redirect_stdout(io.StringIO()):
            kw2, _, start2, grad_vars2, optimizer2 = R.create_nerf(ckpt_args(base, False))
        assert start2 == 2
        kw2.update(near=2.0, far=6.0)
        loss3 = train_iteration(R, K1, (ro, rd), target, kw2, optimizer2, 3)
        out = dict(rays_o=ro, rays_d=rd, target=target, losses=np.array(losses + [loss3], np.float64),
                   lr_after_step3=np.float64(optimizer2.param_groups[0]["lr"]))
        for tag, net in (("c", kw2["network_fn"]), ("f", kw2["network_fine"])):
            for name, prm in net.named_parameters():
                out[f"after3.{tag}.{name}"] = prm.detach()
        shutil.copy(path, os.path.join(GOLD, "ref_ckpt_w64.tar"))
        print(f"ref_ckpt_w64.tar: {os.path.getsize(path) / 1024:.1f} KiB, losses {losses + [loss3]}")
        save("ref_ckpt_w64", **out)
    finally:
        shutil.rmtree(base, ignore_errors=True)


def make_keras(R):
    """Keras layout: kernel [in, out], bias [out]; order pts_linears 0..7, feature ('bottleneck'), views, rgb, alpha"""
    rng = np.random.Generator(np.random.PCG64(4242))
    dims = [(63, 256)] + [(256, 256)] * 4 + [(319, 256)] + [(256, 256)] * 2 + [(256, 256), (283, 128), (128, 3), (256, 1)]
    weights = []
    for fi, fo in dims:
        weights.append((rng.uniform(-1, 1, size=(fi, fo)) / np.sqrt(fi)).astype(np.float32))
        weights.append(rng.uniform(-0.05, 0.05, size=(fo,)).astype(np.float32))
    m = R.ref_nerf.NeRF(8, 256, 63, 27, 5, [4], True)
    m.load_weights_from_keras(weights)
    x = torch.from_numpy(rng.uniform(-1, 1, size=(32, 90)).astype(np.float32))
    with torch.no_grad():
        y = m(x)
    out = dict(seed=4242, x=x, y=y)
    for name, t in m.state_dict().items():
        t = t.double()
        out[f"sd.{name}.shape"] = np.array(t.shape)
        out[f"sd.{name}.sum"] = t.sum()
        out[f"sd.{name}.abssum"] = t.abs().sum()
        out[f"sd.{name}.first"] = t.reshape(-1)[:8].float()
        out[f"sd.{name}.last"] = t.reshape(-1)[-8:].float()
    save("keras_weights", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    R = load_reference()
    make_checkpoint(R)
    make_keras(R)

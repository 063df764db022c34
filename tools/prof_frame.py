"""Small driver for ncu: renders `--chunks` chunks of 32768 C1-shaped rays (coarse 64 + fine 128, view directions) through
render() after one warm-up pass, i.e. per chunk: fused MLP (+compositing) coarse, resample, fused MLP (+compositing) fine,
finish.  `--train` runs 4096-ray optimisation steps instead (forward-with-save, chain, weight gradients, Adam).

    ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_kernel -s 4 -c 2 -o gpurun_out/prof \\
        python tools/prof_frame.py --chunks 2
"""
import argparse
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import nerf_projects_b200 as nb  # noqa: E402
from nerf_projects_b200 import synthetic as SY  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    with contextlib.redirect_stdout(io.StringIO()):
        kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(bench.make_args(), device=dev)
    kw_test["network_fn"].load_state_dict(SY.random_init_state_dict(0, use_viewdirs=True, output_ch=5))
    kw_test["network_fine"].load_state_dict(SY.random_init_state_dict(1, use_viewdirs=True, output_ch=5))
    K, c2w = SY.blender_shaped_camera(800, 800, 30.0)
    ro, rd = nb.get_rays(800, 800, K, torch.from_numpy(c2w).to(dev))
    n = 32768 * args.chunks
    ro, rd = ro.reshape(-1, 3)[:n].contiguous(), rd.reshape(-1, 3)[:n].contiguous()
    if args.train:
        kw = dict(kw_train); kw.update(ndc=False, lindisp=False, perturb=1.0, raw_noise_std=1.0)
        tgt = torch.rand((4096, 3), device=dev)
        for i in range(1 + args.steps):
            with nb.keyed_rng(nb.step_seed(0, i)):
                rgb, _, _, ex = nb.render(800, 800, K, rays=(ro[:4096], rd[:4096]), near=2., far=6., **kw)
            loss = nb.img2mse(rgb, tgt) + nb.img2mse(ex["rgb0"], tgt)
            optim.zero_grad(); loss.backward(); optim.step()
    else:
        with torch.no_grad():
            for _ in range(2):
                nb.render(800, 800, K, chunk=32768, rays=(ro, rd), near=2., far=6., **kw_test)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()

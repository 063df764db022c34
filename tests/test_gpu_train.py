"""train() (the reference's training driver, ipynb:993-1452) end to end on the GPU from an on-disk Blender-format
dataset: loader -> device ray batcher -> fused render -> tensor-core backward -> fused Adam -> checkpoint in the
reference's layout -> resume through create_nerf -> test-set render with on-GPU PSNR/SSIM."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from _cases import T
from test_gpu_render import build

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _write_blender_scene(root, H, W, n_train=8, n_test=2):
    """Teacher renders (calibrated synthetic scene) written as a Blender-synthetic directory tree."""
    import cv2
    import nerf_projects_b200 as nb
    teacher_kw, _, _, _ = build("render_c1_sharp")
    tk = dict(teacher_kw); tk.update(ndc=False, near=2., far=6., white_bkgd=True)
    K, _ = O.blender_camera(H, W, 0.0)
    cax = 2.0 * np.arctan(0.5 * W / K[0][0])
    k = 0
    for split, cnt in (("train", n_train), ("val", 1), ("test", n_test)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for _ in range(cnt):
            c2w = np.eye(4, dtype=np.float32)
            c2w[:3, :4] = O.blender_camera(H, W, 37.0 * k)[1][:3, :4]
            with torch.no_grad():
                rgb = nb.render(H, W, K, chunk=8192, c2w=T(c2w[:3, :4]).to(DEV), **tk)[0]
            rgba = np.concatenate([nb.to8b(rgb.cpu().numpy()), np.full((H, W, 1), 255, np.uint8)], -1)
            cv2.imwrite(os.path.join(root, split, f"r_{k}.png"), cv2.cvtColor(rgba, cv2.COLOR_RGBA2BGRA))
            frames.append({"file_path": f"./{split}/r_{k}", "transform_matrix": c2w.tolist()})
            k += 1
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": float(cax), "frames": frames}, f)


def test_train_driver_blender_dataset_checkpoint_resume_and_metrics(tmp_path):
    import nerf_projects_b200 as nb
    H = W = 40
    data = str(tmp_path / "scene")
    _write_blender_scene(data, H, W)
    # the reference's dict-with-attribute-access args (config.py); a plain attribute namespace works too
    args = nb.Args(
        datadir=data, dataset_type="blender", half_res=False, testskip=1, white_bkgd=True, basedir=str(tmp_path / "logs"),
        expname="exp", multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=128, N_samples=64,
        netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4, lrate_decay=500,
        ft_path=None, no_reload=False, perturb=1.0, raw_noise_std=0.0, no_ndc=False, lindisp=False, N_rand=1024,
        no_batching=False, precrop_iters=0, precrop_frac=0.5, chunk=32768, i_print=1, i_weights=100, i_testset=200, i_video=200,
        rng_seed=7)
    torch.manual_seed(0)
    out = nb.train(args, N_iters=201, log=lambda *a: None)
    assert out["global_step"] == 200 and len(out["losses"]) == 200
    ls = [l for _, l in out["losses"]]
    first, last = float(np.mean(ls[:10])), float(np.mean(ls[-10:]))
    assert np.isfinite(ls).all() and last < 0.6 * first, (first, last)
    # lr schedule applied after the step (ipynb:1278-1282)
    assert abs(out["optimizer"].param_groups[0]["lr"] - 5e-4 * 0.1 ** (200 / 500000)) < 1e-12
    # checkpoints: the reference's dictionary layout and file naming
    ck = sorted(os.listdir(os.path.join(args.basedir, "exp", "checkpoints")))
    assert ck == ["000100.tar", "000200.tar"]
    c = torch.load(os.path.join(args.basedir, "exp", "checkpoints", ck[-1]), map_location="cpu")
    assert sorted(c) == ["global_step", "network_fine_state_dict", "network_fn_state_dict", "optimizer_state_dict"]
    assert c["global_step"] == 200 and "pts_linears.0.weight" in c["network_fn_state_dict"]
    # test-set render with on-GPU metrics, frames saved
    m = out["test_metrics"][200]
    assert set(m) >= {"avg_psnr", "avg_ssim", "avg_mse", "std_psnr"} and m["avg_psnr"] > 5 and 0 < m["avg_ssim"] <= 1
    assert len([f for f in os.listdir(os.path.join(args.basedir, "exp", "testset_000200")) if f[:3].isdigit()]) == 2
    # validation trajectory of the 40 spherical render poses, saved as uint8 arrays
    spiral = np.load(os.path.join(args.basedir, "exp", "exp_spiral_000200_rgb.npy"))
    assert spiral.shape == (40, H, W, 3) and spiral.dtype == np.uint8 and spiral.std() > 1
    assert os.path.exists(os.path.join(args.basedir, "exp", "config.yaml")) and os.path.exists(os.path.join(args.basedir, "exp", "args.txt"))
    # resume: create_nerf picks up 000200.tar, the loop continues at 201 (here: per-image sampling with centre crop)
    args.no_batching, args.precrop_iters = True, 205
    out2 = nb.train(args, N_iters=211, log=lambda *a: None)
    assert out2["global_step"] == 210 and out2["losses"][0][0] == 201
    sd = out2["render_kwargs_train"]["network_fn"].state_dict()
    assert not torch.equal(sd["pts_linears.0.weight"].cpu(), c["network_fn_state_dict"]["pts_linears.0.weight"])
    assert np.mean([l for _, l in out2["losses"]]) < 0.8 * first    # continued from the trained weights, not from scratch
    # render_only + render_test short-circuit (ipynb:1142-1165)
    args.render_only, args.render_test = True, True
    out3 = nb.train(args, log=lambda *a: None)
    assert out3["global_step"] == 200 and out3["test_metrics"][200]["avg_psnr"] > 5     # latest checkpoint = 000200.tar

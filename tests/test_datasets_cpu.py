"""nerf_projects_b200.datasets (load_blender_data / load_llff_data / pose_spherical) against the outputs of the
unmodified reference loaders on the same seeded directory trees (tests/golden/datasets.npz, written by
tools/make_golden_datasets.py).  Host-side formats: everything bit-exact."""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import datasets_synth as DS
from _cases import gold


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def test_load_blender_data_vs_reference(tmp_path):
    from nerf_projects_b200 import datasets as D
    g = gold("datasets")
    root = DS.write_blender(str(tmp_path / "blender"))
    for tag, kw in (("b_full", dict(half_res=False, testskip=1)), ("b_half", dict(half_res=True, testskip=2)),
                    ("b_skip0", dict(half_res=False, testskip=0))):
        imgs, poses, rp, hwf, i_split = D.load_blender_data(root, **kw)
        assert imgs.dtype == g[tag + ".imgs"].dtype and np.array_equal(imgs, g[tag + ".imgs"]), tag
        assert poses.dtype == np.float32 and np.array_equal(poses, g[tag + ".poses"])
        assert np.array_equal(rp.numpy(), g[tag + ".render_poses"]) and rp.shape == (40, 4, 4)
        assert np.array_equal(np.array(hwf, np.float64), g[tag + ".hwf"])
        for got, key in zip(i_split, ("i_train", "i_val", "i_test")):
            assert np.array_equal(got, g[f"{tag}.{key}"])


def test_pose_spherical_vs_reference():
    from nerf_projects_b200 import datasets as D
    g = gold("datasets")
    got = np.stack([D.pose_spherical(a, p, r).numpy() for a, p, r in
                    ((0., -30., 4.), (37.5, -30., 4.), (-180., 12., 2.5), (91., 0., 1.))])
    assert np.array_equal(got, g["pose_spherical"])


def test_load_llff_data_vs_reference(tmp_path):
    from nerf_projects_b200 import datasets as D
    g = gold("datasets")
    root = DS.write_llff(str(tmp_path / "llff"))
    for tag, kw in (("l_default", dict(factor=4)), ("l_spherify", dict(factor=4, spherify=True)),
                    ("l_norecenter", dict(factor=4, recenter=False, bd_factor=None))):
        with _quiet():
            images, poses, bds, rp, i_test = D.load_llff_data(root, **kw)
        assert images.dtype == np.float32 and np.array_equal(images, g[tag + ".images"]), tag
        assert poses.dtype == np.float32 and np.array_equal(poses, g[tag + ".poses"]), tag
        assert np.array_equal(bds, g[tag + ".bds"]) and np.array_equal(rp, g[tag + ".render_poses"]), tag
        assert int(i_test) == int(g[tag + ".i_test"])
    # hwf column: H, W of the minified images and focal / factor
    assert tuple(poses[0, :, 4]) == (24.0, 32.0, 815.0 / 4)


def test_llff_minify_when_directory_missing(tmp_path):
    """A missing images_<factor> directory is produced (OpenCV area resize) with the right geometry."""
    from nerf_projects_b200 import datasets as D
    root = DS.write_llff(str(tmp_path / "llff"), H=8, W=12, n=3, factor=4)
    with _quiet():
        images, poses, *_ = D.load_llff_data(root, factor=2)
    assert os.path.isdir(os.path.join(root, "images_2")) and images.shape == (3, 16, 24, 3)
    assert tuple(poses[0, :, 4]) == (16.0, 24.0, 815.0 / 2)
    with _quiet():
        assert D._load_data(root, factor=None, load_imgs=False)[0].shape == (3, 5, 3)


def test_train_driver_dataset_dispatch(tmp_path):
    """trainer.load_dataset = the dataset block of train() (ipynb:1017-1098): Blender white-background compositing and
    near/far = 2/6; LLFF holdout selection, NDC bounds 0/1 or bounds from `bds` with no_ndc; unknown types rejected."""
    import types
    from nerf_projects_b200 import trainer
    from nerf_projects_b200 import datasets as D
    b = DS.write_blender(str(tmp_path / "blender"))
    raw = D.load_blender_data(b, False, 2)[0]
    args = types.SimpleNamespace(dataset_type="blender", datadir=b, half_res=False, testskip=2, white_bkgd=True)
    images, poses, render_poses, hwf, (i_train, i_val, i_test), near, far = trainer.load_dataset(args)
    assert (near, far) == (2., 6.) and images.shape == (5 + 2 + 3, 12, 16, 3) and len(i_train) == 5 and len(i_test) == 3
    assert np.array_equal(images, raw[..., :3] * raw[..., -1:] + (1. - raw[..., -1:]))
    args.white_bkgd = False
    assert np.array_equal(trainer.load_dataset(args)[0], raw[..., :3])
    ll = DS.write_llff(str(tmp_path / "llff"))
    args = types.SimpleNamespace(dataset_type="llff", datadir=ll, factor=4, spherify=False, llffhold=3, no_ndc=False)
    with _quiet():
        images, poses, render_poses, hwf, (i_train, i_val, i_test), near, far = trainer.load_dataset(args)
    assert (near, far) == (0., 1.) and list(i_test) == [0, 3, 6] and list(i_train) == [1, 2, 4, 5]
    assert poses.shape == (7, 3, 4) and tuple(hwf) == (24.0, 32.0, 815.0 / 4) and render_poses.shape == (120, 3, 5)
    args.no_ndc, args.llffhold = True, 0
    with _quiet():
        out = trainer.load_dataset(args)
        bds = D.load_llff_data(ll, 4)[2]
    assert abs(out[5] - float(bds.min()) * .9) < 1e-6 and abs(out[6] - float(bds.max())) < 1e-6 and len(out[4][2]) == 1
    lm = DS.write_linemod(str(tmp_path / "linemod"))
    with _quiet():
        out = trainer.load_dataset(types.SimpleNamespace(dataset_type="LINEMOD", datadir=lm, half_res=False, testskip=1,
                                                         white_bkgd=True))
    assert out[0].shape == (12, 10, 14, 3) and (out[5], out[6]) == (0.0, 3.0) and out[7].shape == (3, 3) and out[7][0, 0] == 574.4
    dv = DS.write_deepvoxels(str(tmp_path / "dv"))
    with _quiet():
        out = trainer.load_dataset(types.SimpleNamespace(dataset_type="deepvoxels", datadir=dv, shape="cube", testskip=8))
    assert len(out) == 7 and abs((out[6] - out[5]) - 2.0) < 1e-6 and out[3][0] == 512
    with pytest.raises(ValueError):
        trainer.load_dataset(types.SimpleNamespace(dataset_type="nope", datadir="."))


def test_load_linemod_and_deepvoxels_vs_reference(tmp_path):
    from nerf_projects_b200 import datasets as D
    g = gold("datasets")
    lm = DS.write_linemod(str(tmp_path / "linemod"))
    for tag, kw in (("m_full", dict(half_res=False, testskip=1)), ("m_half", dict(half_res=True, testskip=2))):
        with _quiet():
            imgs, poses, rp, hwf, K, i_split, near, far = D.load_LINEMOD_data(lm, **kw)
        assert imgs.dtype == g[tag + ".imgs"].dtype and np.array_equal(imgs, g[tag + ".imgs"]), tag
        assert np.array_equal(poses, g[tag + ".poses"]) and np.array_equal(np.array(hwf, np.float64), g[tag + ".hwf"])
        assert np.array_equal(np.array(K, np.float64), g[tag + ".K"]) and np.array_equal(i_split[2], g[tag + ".i_test"])
        assert float(near) == float(g[tag + ".near"]) and float(far) == float(g[tag + ".far"]) and rp.shape == (40, 4, 4)
    dv = DS.write_deepvoxels(str(tmp_path / "dv"))
    for tag, skip in (("d_skip8", 8), ("d_skip3", 3)):
        with _quiet():
            imgs, poses, rp, hwf, i_split = D.load_dv_data(scene="cube", basedir=dv, testskip=skip)
        assert imgs.dtype == np.float32 and np.array_equal(imgs, g[tag + ".imgs"]), tag
        assert poses.dtype == np.float32 and np.array_equal(poses, g[tag + ".poses"])
        assert np.array_equal(rp, g[tag + ".render_poses"]) and np.array_equal(np.array(hwf, np.float64), g[tag + ".hwf"])
        assert np.array_equal(i_split[1], g[tag + ".i_val"]) and np.array_equal(i_split[2], g[tag + ".i_test"])

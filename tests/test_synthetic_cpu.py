"""The package's synthetic benchmark inputs equal the oracle's independently written generators (bit for bit), so the
bench's GPU arm and its CPU-baseline arm see the same weights and cameras without the product importing `oracle/`."""
import numpy as np
import torch

from nerf_projects_b200 import synthetic as S
from oracle import nerf_oracle as O


def test_random_init_matches_oracle_generator():
    for seed, arch in ((0, dict(use_viewdirs=True, output_ch=5)), (1, dict(use_viewdirs=True, output_ch=5)),
                       (100, dict(use_viewdirs=False, input_ch_views=0, output_ch=4))):
        a, b = S.random_init_state_dict(seed, **arch), O.to_torch(O.make_params(seed, **arch))
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_cameras_match_oracle_generators():
    for H, W, th in ((800, 800, 30.0), (100, 100, 75.0), (48, 64, 123.0)):
        Ka, ca = S.blender_shaped_camera(H, W, th)
        Kb, cb = O.blender_camera(H, W, th)
        assert np.array_equal(Ka, Kb) and np.array_equal(ca, cb)
    Ka, ca = S.llff_shaped_camera()
    Kb, cb = O.llff_camera()
    assert np.array_equal(Ka, Kb) and np.array_equal(ca, cb)
    assert np.array_equal(S.spherical_pose(40.0, -30.0, 4.0), O.pose_spherical(40.0, -30.0, 4.0))

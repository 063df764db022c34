"""Synthetic benchmark inputs: seeded random-init NeRF weights and Blender- / LLFF-shaped cameras.

bench.py and the tools build their inputs from this module, so that the product arm never touches `oracle/`
(the oracle keeps its own, independently written generators; `tests/test_synthetic_cpu.py` checks that both
produce identical values).  No GPU needed."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def nerf_layer_shapes(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=(4,), use_viewdirs=True):
    """(name, out_features, in_features) of every nn.Linear of the reference NeRF module in state_dict order
    (nerf.py:32-55)."""
    shapes = [(f"pts_linears.{i}", W, input_ch if i == 0 else (W + input_ch if (i - 1) in skips else W))
              for i in range(D)]
    shapes.append(("views_linears.0", W // 2, input_ch_views + W))
    if use_viewdirs:
        shapes += [("feature_linear", W, W), ("alpha_linear", 1, W), ("rgb_linear", 3, W // 2)]
    else:
        shapes.append(("output_linear", output_ch, W))
    return shapes


def random_init_state_dict(seed: int, **arch) -> Dict[str, torch.Tensor]:
    """nn.Linear's default init (weight and bias ~ U(+-1/sqrt(fan_in))) drawn from numpy's PCG64, so the values
    are independent of the torch build.  Loadable with NeRF.load_state_dict."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for name, fan_out, fan_in in nerf_layer_shapes(**arch):
        b = 1.0 / math.sqrt(fan_in)
        sd[name + ".weight"] = torch.from_numpy(rng.uniform(-b, b, size=(fan_out, fan_in)).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy(rng.uniform(-b, b, size=(fan_out,)).astype(np.float32))
    return sd


def spherical_pose(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """Camera-to-world of load_blender.py:10-34 (translate along z, tilt by phi, spin by theta, Blender axes)."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    th, ph = theta_deg / 180.0 * np.pi, phi_deg / 180.0 * np.pi
    trans = np.eye(4, dtype=np.float32)
    trans[2, 3] = radius
    rot_phi = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]],
                       dtype=np.float32)
    rot_theta = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]],
                         dtype=np.float32)
    axes = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return (axes @ (rot_theta @ (rot_phi @ trans))).astype(np.float32)


def blender_shaped_camera(H: int, W: int, theta_deg: float = 30.0):
    """(K [3,3] float64, c2w [3,4] float32): lego's camera_angle_x (load_blender.py:72-73), pose_spherical(theta, -30, 4)."""
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float64)
    return K, spherical_pose(theta_deg, -30.0, 4.0)[:3, :4]


def llff_shaped_camera(H: int = 756, W: int = 1008, focal: float = 815.0):
    """(K, c2w) of a forward-facing camera looking down -z with a small offset and tilt (fern at factor 4)."""
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float64)
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float32)
    t = np.array([[0.1], [-0.05], [0.2]], dtype=np.float32)
    return K, np.concatenate([R, t], axis=1).astype(np.float32)

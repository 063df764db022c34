"""nerf_projects_b200 -- B200-native (sm_100a) implementation of the
original-NeRF render/train hot path of isaacchunn/nerf-projects (`nerf/`).

Drop-in surface (same names and signatures as the reference's notebook cells
and helper modules):

    render, render_rays, batchify_rays, render_path, run_network, batchify,
    raw2outputs, create_nerf, NeRF, Embedder, get_embedder, get_rays,
    get_rays_np, ndc_rays, sample_pdf, img2mse, mse2psnr, to8b

plus RayBatcher, the device-side version of the ray-batch sampling block of train().

All numerical work runs in hand-written CUDA (libnerf_b200.so, C ABI in
include/nerf_b200.h).  There is no CPU fallback: importing works anywhere, but
calling an op without the built library or without a CUDA device raises.
"""
from ._lib import NerfB200Error, LIB_PATH, load as load_library  # noqa: F401
from .embedder import Embedder, get_embedder  # noqa: F401
from .model import NeRF  # noqa: F401
from .mlp import get_precision, precision, release_workspaces, set_precision  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .render import (batchify, batchify_rays, create_nerf, get_rays, get_rays_np, img2mse, mse2psnr,  # noqa: F401
                     ndc_rays, raw2outputs, render, render_path, render_rays, run_network, sample_pdf, to8b)
from .data import RayBatcher, precrop_pixel_ids  # noqa: F401
from .metrics import calculate_lpips, calculate_metrics, calculate_ssim  # noqa: F401
from .rng import keyed_rng, step_seed  # noqa: F401
from .datasets import load_blender_data, load_dv_data, load_LINEMOD_data, load_llff_data, pose_spherical  # noqa: F401
from .trainer import train  # noqa: F401
from .config import (Args, AttrDict, create_default_config, default_config, load_config, load_or_create_config,  # noqa: F401
                     load_yaml, save_config, save_yaml, validate_config)
from . import ops  # noqa: F401

__all__ = ["render", "render_rays", "batchify_rays", "render_path", "run_network", "batchify", "raw2outputs",
           "create_nerf", "NeRF", "Embedder", "get_embedder", "get_rays", "get_rays_np", "ndc_rays", "sample_pdf",
           "img2mse", "mse2psnr", "to8b", "set_precision", "get_precision", "precision", "release_workspaces", "FusedAdam", "NerfB200Error", "RayBatcher", "precrop_pixel_ids",
           "keyed_rng", "step_seed", "calculate_ssim", "calculate_metrics", "calculate_lpips",
           "load_blender_data", "load_llff_data", "load_LINEMOD_data", "load_dv_data", "pose_spherical", "train",
           "Args", "default_config", "load_config", "save_config", "AttrDict", "load_yaml", "save_yaml",
           "create_default_config", "validate_config", "load_or_create_config"]

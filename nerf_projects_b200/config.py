"""Experiment configuration in the reference's format (nerf/utils.py:8-208, nerf/yaml/*, the notebook's AttrDict cell):
a flat YAML file of the ``args.*`` attributes that ``create_nerf`` and ``train`` read.  ``load_config`` returns a dict with
attribute access, reference defaults filled in for missing keys, so the reference's yaml files drive ``train()`` unchanged:

    args = load_config("nerf/yaml/lego_blender200k"); out = train(args)

Values are kept as YAML parses them (``raw_noise_std: 1e0`` stays the string '1e0', which ``render_rays`` accepts like the
reference does).
"""
from __future__ import annotations

import os
from typing import Any, Dict


class Args(dict):
    """dict with attribute access (the notebook's AttrDict): ``args.N_rand`` == ``args['N_rand']``; ``args.items()`` works."""
    __slots__ = ()

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(f"No such attribute: {name}") from e

    def __setattr__(self, name, value):
        self[name] = value

    @classmethod
    def wrap(cls, obj):
        if isinstance(obj, dict):
            return cls({k: cls.wrap(v) for k, v in obj.items()})
        if isinstance(obj, list):
            return [cls.wrap(v) for v in obj]
        return obj


def default_config() -> Dict[str, Any]:
    """The reference's default values (utils.py:60-131), grouped as its yaml files are."""
    experiment = dict(expname="nerf_experiment", basedir="./logs/", datadir="./data/llff/fern", dataset_type="llff")
    network = dict(netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1024 * 64, i_embed=0,
                   multires=10, multires_views=4, use_viewdirs=False)
    optimisation = dict(N_rand=32 * 32 * 4, lrate=5e-4, lrate_decay=250, chunk=1024 * 32, no_batching=False,
                        no_reload=False, ft_path=None, precrop_iters=0, precrop_frac=0.5)
    sampling = dict(N_samples=64, N_importance=0, perturb=1.0, raw_noise_std=0.0, lindisp=False)
    rendering = dict(render_only=False, render_test=False, render_factor=0, white_bkgd=False)
    data = dict(testskip=8, shape="greek", half_res=False, factor=8, no_ndc=False, spherify=False, llffhold=8)
    cadence = dict(i_print=100, i_img=500, i_weights=10000, i_testset=50000, i_video=50000)
    out: Dict[str, Any] = {}
    for group in (experiment, network, optimisation, sampling, rendering, data, cadence):
        out.update(group)
    return out


_REQUIRED = ("expname", "basedir", "datadir", "dataset_type")


def _check_config(cfg: Dict[str, Any]) -> bool:
    """load_config's own check (raises): identification keys present, counts positive, dataset type known."""
    for k in _REQUIRED:
        if k not in cfg:
            raise ValueError(f"Missing required configuration key: {k}")
    for k in ("netdepth", "netwidth", "N_samples", "N_rand"):
        if k in cfg and (not isinstance(cfg[k], int) or cfg[k] <= 0):
            raise ValueError(f"{k} must be a positive integer, got {cfg[k]!r}")
    if cfg["dataset_type"] not in ("llff", "blender", "deepvoxels", "LINEMOD"):
        raise ValueError(f"Unknown dataset_type {cfg['dataset_type']!r}")
    return True


def load_config(source, defaults: bool = True) -> Args:
    """YAML file path (or a dict) -> Args; keys missing from the file take the reference defaults."""
    if isinstance(source, dict):
        loaded = dict(source)
    else:
        import yaml
        if not os.path.exists(source):
            raise FileNotFoundError(f"Configuration file not found: {source}")
        with open(source, "r", encoding="utf-8") as f:
            loaded = yaml.safe_load(f)
        if loaded is None:
            raise ValueError(f"Configuration file is empty: {source}")
    cfg = default_config() if defaults else {}
    cfg.update(loaded)
    _check_config(cfg)
    return Args.wrap(cfg)


def save_config(cfg: Dict[str, Any], path: str) -> None:
    import yaml
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        yaml.dump(dict(cfg), f, default_flow_style=False, indent=2)


# ---- the reference's own names (nerf/utils.py, notebook AttrDict cell), same behaviour: plain dicts in and out,
# ---- validate_config reports by return value and message instead of raising

class AttrDict(Args):
    __slots__ = ()

    @classmethod
    def from_obj(cls, obj):
        return cls.wrap(obj)


_REFERENCE_REQUIRED = ("expname", "basedir", "datadir", "netdepth", "netwidth", "netdepth_fine", "netwidth_fine", "N_rand",
                       "lrate", "lrate_decay", "chunk", "netchunk", "N_samples", "N_importance", "perturb", "use_viewdirs",
                       "i_embed", "multires", "multires_views", "raw_noise_std", "dataset_type")


def validate_config(config: Dict[str, Any]) -> bool:
    """utils.py:134-181: True when every required key is present and netdepth / netwidth / lrate / N_samples / chunk > 0."""
    missing = [k for k in _REFERENCE_REQUIRED if k not in config]
    if missing:
        print(f"Missing required configuration keys: {missing}")
        return False
    for key, what in (("netdepth", "Network depth (netdepth)"), ("netwidth", "Network width (netwidth)"),
                      ("lrate", "Learning rate (lrate)"), ("N_samples", "Number of samples (N_samples)"),
                      ("chunk", "Chunk size (chunk)")):
        if config[key] <= 0:
            print(f"{what} must be positive")
            return False
    return True


def load_yaml(config_path: str) -> Dict[str, Any]:
    """utils.py:8-39: the file's own keys as a plain dict (no defaults filled in)."""
    return _plain_yaml(config_path)


def _plain_yaml(path):
    import yaml
    if not os.path.exists(path):
        raise FileNotFoundError(f"Configuration file not found: {path}")
    with open(path, "r", encoding="utf-8") as f:
        cfg = yaml.safe_load(f)
    if cfg is None:
        raise ValueError(f"Configuration file is empty: {path}")
    return cfg


def save_yaml(config: Dict[str, Any], config_path: str) -> None:
    save_config(config, config_path)


def create_default_config() -> Dict[str, Any]:
    return default_config()


def load_or_create_config(config_path: str) -> Dict[str, Any]:
    """utils.py:183-208: load the file, or write the defaults there first; validation is reported, not enforced."""
    if os.path.exists(config_path):
        print(f"Loading configuration from {config_path}")
        config = load_yaml(config_path)
    else:
        print(f"Configuration file {config_path} not found, creating default config")
        config = create_default_config()
        save_yaml(config, config_path)
        print(f"Default configuration saved to {config_path}")
    if not validate_config(config):
        print("Warning: Configuration validation failed")
    else:
        print("Configuration validation passed! Arguments are valid and correctly set.")
    return config

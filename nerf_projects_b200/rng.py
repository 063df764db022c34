"""Random draws keyed by (seed, global ray index) -- SURVEY.md 8(e), csrc/rng.cu.

By default ``render_rays`` draws ``t_rand`` / ``u`` / the density noise with ``torch.rand`` /
``torch.randn`` exactly where the reference does (ipynb:428-444, nerf_helpers.py:404,
ipynb:312-325), so a run depends on chunking and on the number of GPUs.  Inside

    with keyed_rng(seed, ray_offset=lo):
        render(...)            # or render_rays / batchify_rays

every draw is instead a pure function of (seed, global ray index, sample, stream): the same
rays give bit-identical draws whatever the chunk size or world size.  ``batchify_rays`` advances
the ray index by the chunk start; ``parallel.render_sharded`` / ``train_step_sharded`` enter the
context with ``ray_offset`` = the rank's first ray when given ``rng_seed``.  The reference's
signatures are untouched (the state is thread-local, not an argument).
"""
from __future__ import annotations

import contextlib
import threading

import torch

from . import _lib as L
from . import ops

T_RAND, U, NOISE_COARSE, NOISE_FINE = 0, 1, 2, 3

_state = threading.local()


def current():
    """(seed, ray_offset) of the innermost keyed_rng context, or None."""
    return getattr(_state, "key", None)


@contextlib.contextmanager
def keyed_rng(seed: int, ray_offset: int = 0):
    prev = current()
    _state.key = (int(seed) & 0xFFFFFFFFFFFFFFFF, int(ray_offset))
    try:
        yield
    finally:
        _state.key = prev


@contextlib.contextmanager
def advanced(rays: int):
    """Shift the ray index of the active context by ``rays`` (used per chunk by batchify_rays)."""
    key = current()
    if key is None:
        yield
        return
    _state.key = (key[0], key[1] + int(rays))
    try:
        yield
    finally:
        _state.key = key


def step_seed(seed: int, step: int) -> int:
    """Distinct 64-bit key per optimisation step (splitmix64 of seed + step)."""
    z = (int(seed) + 0x9E3779B97F4A7C15 * (int(step) + 1)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def draw(R: int, S: int, stream_id: int, device, normal: bool = False, scale: float = 1.0, key=None):
    """[R,S] fp32 draws of the rays key[1] .. key[1]+R-1."""
    seed, off = key if key is not None else current()
    out = torch.empty((R, S), dtype=torch.float32, device=device)
    L.require_cuda(out)
    L.check(L.load().nerfb_keyed_rng(seed, off, int(R), int(S), int(stream_id), int(bool(normal)), float(scale),
                                     L.ptr(out), L.stream()), "nerfb_keyed_rng")
    ops._count()
    return out


def render_rays_draws(R: int, N_samples: int, N_importance: int, perturb: bool, noise_std: float, device):
    """(t_rand, u, noise_coarse, noise_fine) of one render_rays call in one launch; None where not needed."""
    seed, off = current()
    mk = lambda s: torch.empty((R, s), dtype=torch.float32, device=device)
    t_rand = mk(N_samples) if perturb else None
    u = mk(N_importance) if perturb and N_importance > 0 else None
    nc = mk(N_samples) if noise_std > 0.0 else None
    nf = mk(N_samples + N_importance) if noise_std > 0.0 and N_importance > 0 else None
    if t_rand is None and nc is None:
        return None, None, None, None
    L.require_cuda(*(t for t in (t_rand, u, nc, nf) if t is not None))
    L.check(L.load().nerfb_render_rays_rng(seed, off, int(R), int(N_samples), int(N_importance), float(noise_std),
                                           L.ptr(t_rand), L.ptr(u), L.ptr(nc), L.ptr(nf), L.stream()),
            "nerfb_render_rays_rng")
    ops._count()
    return t_rand, u, nc, nf

"""Shared description of the golden cases written by tools/make_golden.py
(inputs are re-derived from seeds; outputs come from the reference)."""
import os

import numpy as np
import torch

from oracle import nerf_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _pc():
    return O.make_params(0, use_viewdirs=True, output_ch=5)


def _pf():
    return O.make_params(1, use_viewdirs=True, output_ch=5)


# tag -> description of the render; "scene" cases take their (calibrated) density head from the fixture
RENDER_CASES = {
    "render_c0": dict(pc=lambda: O.make_params(100, use_viewdirs=False, input_ch_views=0, output_ch=4), pf=None,
                      use_viewdirs=False, N_samples=64, N_importance=0, white_bkgd=False, ndc=False, lindisp=False),
    "render_c0_scene": dict(pc=lambda: O.make_params(101, use_viewdirs=False, input_ch_views=0, output_ch=4),
                            pf=None, use_viewdirs=False, N_samples=64, N_importance=0, white_bkgd=True, ndc=False,
                            lindisp=False),
    "render_c1": dict(pc=_pc, pf=_pf, use_viewdirs=True, N_samples=64, N_importance=128, white_bkgd=True,
                      ndc=False, lindisp=False),
    "render_c1_sharp": dict(pc=_pc, pf=_pf, use_viewdirs=True, N_samples=64, N_importance=128, white_bkgd=True,
                            ndc=False, lindisp=False),
    "render_lindisp": dict(pc=_pc, pf=_pf, use_viewdirs=True, N_samples=64, N_importance=128, white_bkgd=False,
                           ndc=False, lindisp=True),
    "render_llff": dict(pc=_pc, pf=_pf, use_viewdirs=True, N_samples=64, N_importance=128, white_bkgd=False,
                        ndc=True, lindisp=False),
    "render_nov_40_24": dict(pc=lambda: O.make_params(7, use_viewdirs=False, input_ch_views=0, output_ch=5),
                             pf="same", use_viewdirs=False, N_samples=40, N_importance=24, white_bkgd=True,
                             ndc=False, lindisp=False),
}


def _apply_overrides(p, g, prefix):
    p = dict(p)
    for k, v in g.items():
        if k.startswith(prefix + "."):
            p[k[len(prefix) + 1:]] = np.ascontiguousarray(v)
    return p


def case_params(tag, fixture=None):
    """Parameters of a golden case: seeded init; "scene" cases additionally get the band-limiting of
    O.smooth_spectrum and the calibrated density head stored in the fixture."""
    c = RENDER_CASES[tag]
    g = fixture if fixture is not None else gold(tag)
    scene = any(k.startswith("override.") for k in g)
    sm = O.smooth_spectrum if scene else (lambda p: p)
    pc = _apply_overrides(sm(c["pc"]()), g, "override.c")
    if c["pf"] == "same":
        pf = pc
    elif c["pf"] is None:
        pf = None
    else:
        pf = _apply_overrides(sm(c["pf"]()), g, "override.f")
    return pc, pf


def maps_close(got, ref, tol, min_frac=1.0, cap=None, what=""):
    """|got-ref| <= tol on at least ``min_frac`` of the rays and <= cap on all of
    them.  Returns the number of rays outside ``tol`` (the reported outlier set)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref).reshape(got.shape[0], -1).max(-1)
    bad = err > tol
    assert (~bad).mean() >= min_frac, f"{what}: only {(~bad).mean():.3f} of rays within {tol} (max err {err.max():.3e})"
    if cap is not None:
        assert err.max() <= cap, f"{what}: max err {err.max():.3e} > cap {cap}"
    return int(bad.sum())


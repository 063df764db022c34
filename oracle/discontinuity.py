"""The discontinuity set of ``render_rays`` -- TEST INFRASTRUCTURE ONLY (like everything under ``oracle/``).

``render_rays`` (ipynb:359-492) has two places where its output is a DIScontinuous function of the MLP outputs, so that
a correct implementation whose MLP results differ from the reference's in the last bits (another valid fp32 summation
order) or at bf16 rounding level can legitimately land on the other side.  Parity tests compute the set of affected rays
from the ORACLE (never from the implementation under test), bound its size, and require EVERY other ray to be within
the north-star tolerance.

D1  The last sample's interval is 1e10 (ipynb:300), so alpha_last = 1 - exp(-relu(sigma_last) * 1e10 * |d|) is a step
    function of sigma_last at 0.  fp32 mode reproduces sigma to ~1e-6: nothing is excluded.  bf16 mode: rays whose
    reference |sigma_last| <= TAU_BF16 * max(1, max|sigma|) (the final pass; the coarse pass likewise for rgb0/acc0).
D2  ``sample_pdf`` (nerf_helpers.py:396-439) is discontinuous in the coarse weights: (i) an interpolation denominator
    below 1e-5 is replaced by 1 (:434), and an empty bin has pdf 1e-5 / sum(w + 1e-5), i.e. it sits ON that threshold for
    every nearly opaque ray; (ii) across a run of empty bins the inverse cdf rises by the whole width of the run within
    ~1e-5 of u, so a sample whose u lies next to such a step lands on the near or the far side of the empty region
    depending on the last bits of the cdf.  Whether that matters depends on the scene (moving one of 192 samples between
    two well-sampled places is usually harmless), so the set is MEASURED: the oracle re-evaluates the fine pass on
    alternative depths -- the other branch of (i) for every denominator within D2_DELTA of 1e-5, and u shifted by
    +-U_SHIFT for (ii) -- and a ray is in D2 when any of its maps (rgb, acc, depth / far, z_std) moves by more than
    SENSITIVITY * tol under any alternative.
"""
from __future__ import annotations

import torch

from . import nerf_oracle as O

TAU_BF16 = 1.5e-2    # the calibrated scenes' density heads amplify the bf16 rounding of h7: |sigma| errors up to ~1 % of max|sigma|
# distance of a denominator from 1e-5 inside which the branch may differ, and the shift of u (= of the cdf) that the
# implementation's coarse weights may cause: 4 ulp(1.0) in fp32 mode; bf16-rounding level in bf16 mode
D2_DELTA = {"fp32": 4 * 5.96e-8, "bf16": 3e-7}
U_SHIFT = {"fp32": 4 * 5.96e-8, "bf16": 2e-3}
SENSITIVITY = 0.25


def sigma_last_set(raw, precision):
    """D1 for one pass: raw [R,S,4|5] of the oracle."""
    R = raw.shape[0]
    if precision != "bf16":
        return torch.zeros(R, dtype=torch.bool)
    return raw[:, -1, 3].abs() <= TAU_BF16 * max(1.0, float(raw[..., 3].abs().max()))


def _resample(z_coarse, weights0, N_importance, u_shift=0.0, flip_delta=None):
    """The oracle's sample_pdf arithmetic (deterministic u) on u + u_shift, optionally with the OTHER branch of
    ``denom < 1e-5 -> 1`` taken for the samples whose denominator is within flip_delta of 1e-5."""
    bins = 0.5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    w = weights0[..., 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.0, 1.0, steps=N_importance, dtype=cdf.dtype)
    u = torch.clamp(u + u_shift, 0.0, 1.0).expand(list(cdf.shape[:-1]) + [N_importance]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    denom = c_hi - c_lo
    use_one = denom < 1e-5
    if flip_delta is not None:
        use_one = torch.where((denom - 1e-5).abs() <= flip_delta, ~use_one, use_one)
    t = (u - c_lo) / torch.where(use_one, torch.ones_like(denom), denom)
    return b_lo + t * (b_hi - b_lo)


def excluded_rays(ref, rays, p_coarse, p_fine, *, N_samples, N_importance, white_bkgd, tol, precision, depth_scale=1.0,
                  **net):
    """ref: the oracle's ``render_rays(rays, ..., retraw=True)`` result for packed ``rays`` (perturb = 0).
    -> (mask [R] of D1 u D2, mask of the coarse-pass D1 alone, info dict)."""
    R = rays.shape[0]
    d1 = sigma_last_set(ref["raw"], precision)
    d1c = torch.zeros(R, dtype=torch.bool)
    d2 = torch.zeros(R, dtype=torch.bool)
    evaluated = 0
    if N_importance > 0:
        d1c = sigma_last_set(ref["raw0"], precision)
        zc, w0 = ref["z_coarse"], ref["weights0"]
        span = float((zc[:, -1] - zc[:, 0]).abs().max())
        alts = [_resample(zc, w0, N_importance, 0.0, D2_DELTA[precision]),
                _resample(zc, w0, N_importance, +U_SHIFT[precision]),
                _resample(zc, w0, N_importance, -U_SHIFT[precision])]
        base = U_SHIFT[precision] if precision == "bf16" else 0.0
        for k, z_alt in enumerate(alts):
            # fp32: only rays where a sample actually jumped; bf16: a 2e-3 shift moves every sample a little -- all rays
            moved = (z_alt - ref["z_samples"]).abs().max(-1)[0] > (1e-6 * span if k == 0 or base == 0.0 else 0.0)
            rows = moved.nonzero().flatten()
            if rows.numel() == 0:
                continue
            evaluated += int(rows.numel())
            rr = rays[rows]
            z, _ = torch.sort(torch.cat([zc[rows], z_alt[rows]], -1), -1)
            pts = rr[:, None, 0:3] + rr[:, None, 3:6] * z[..., :, None]
            vd = rr[:, -3:] if rr.shape[-1] > 8 else None
            with torch.no_grad():
                raw = O.run_network(pts, vd, p_fine if p_fine is not None else p_coarse, **net)
                rgb, disp, acc, w, depth = O.raw2outputs(raw, z, rr[:, 3:6], 0.0, white_bkgd)
            zstd = torch.std(z_alt[rows], dim=-1, unbiased=False)
            move = torch.stack([(rgb - ref["rgb_map"][rows]).abs().max(-1)[0], (acc - ref["acc_map"][rows]).abs(),
                                (depth - ref["depth_map"][rows]).abs() / depth_scale,
                                (zstd - ref["z_std"][rows]).abs()], -1).max(-1)[0]
            d2[rows[move > SENSITIVITY * tol]] = True
    info = {"rays": R, "D1_sigma_last": int((d1 | d1c).sum()), "D2_resample_sensitive": int(d2.sum()),
            "fine_passes_reevaluated": evaluated}
    return (d1 | d1c | d2), d1c, info


MAX_OUTLIER_FRACTION = 0.005


def gate(err, tol, explained, what, cap_frac=MAX_OUTLIER_FRACTION):
    """The parity gate: ``err`` [R] = per-ray max-abs error against the reference.  EVERY ray outside the
    discontinuity set ``explained`` must be within ``tol``; rays beyond ``tol`` must all lie inside the set AND be few
    (at most ``cap_frac`` = 0.5 % of the rays; one ray is always allowed, for the 32-160-ray fixtures).  Returns a report string."""
    err = torch.as_tensor(err, dtype=torch.float64).reshape(-1)
    explained = torch.as_tensor(explained, dtype=torch.bool).reshape(-1)
    bad = err > tol
    n, nbad = err.numel(), int(bad.sum())
    unexplained = bad & ~explained
    inside = float(err[~bad].max()) if nbad < n else float("nan")
    rep = (f"{what}: max err {float(err.max()):.3e} (within-tol rays: {inside:.3e}, tol {tol:g}); {nbad} of {n} rays beyond tol, "
           f"{int(unexplained.sum())} of them outside the discontinuity set ({int(explained.sum())} rays in the set)")
    assert not bool(unexplained.any()), rep
    assert nbad <= max(1, int(cap_frac * n)), rep + f" -- more than {cap_frac:.1%} of the rays"
    return rep

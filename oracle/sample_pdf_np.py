"""Bit-level numpy restatement of the reference ``sample_pdf``
(/root/reference/nerf/nerf_helpers.py:372-439) as executed by torch CPU fp32.

TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header).

Why a second restatement: the CUDA kernel has to reproduce the *indices*
``searchsorted(cdf, u, right=True)`` bit-exactly, and whether the last sample
lands in bin M-2 or M-1 depends on the last bit of ``sum(w)`` and of the CDF.
This file spells out, with nothing but IEEE fp32/fp64 scalar semantics, the
exact arithmetic the ATen CPU kernels perform, so the kernel in
``nerf_projects_b200/csrc/sample_pdf.cu`` can follow it operation by
operation:

  * ``w = weights + 1e-5``                       fp32 add            (:396)
  * ``S = torch.sum(w, -1)``                     8-lane vectorised fp32 sum:
        vectors of 8 lanes are dealt round-robin to 4 accumulators for the
        first 4*floor(nvec/4) vectors, leftover vectors go to accumulator 0,
        then acc0 += acc1; acc0 += acc2; acc0 += acc3 (lane-wise); then a
        scalar tail over the n%8 trailing elements (sequential, starting from
        0), then the 8 lanes of acc0 are added to it in lane order.   (:397)
  * ``pdf = w / S``                              IEEE fp32 division  (:397)
  * ``cdf = cumsum(pdf)``                        fp64 running sum, each prefix
        rounded to fp32 on output; cdf = [0, ...]                    (:398-399)
  * ``u = linspace(0,1,N)``                      step = fl32(1/(N-1));
        i <  N/2 : fl32(step * i)
        i >= N/2 : fl32(fma(-step, N-1-i, 1))   (single rounding)    (:404)
  * ``inds = #{j : cdf[j] <= u}``                searchsorted right  (:423)
  * below = max(inds-1,0), above = min(inds, M-1)                    (:424-425)
  * denom = cdf[above]-cdf[below]; denom<1e-5 -> 1; t=(u-cdf[below])/denom;
    sample = bins[below] + t*(bins[above]-bins[below]), every op rounded to
    fp32 separately (no FMA contraction)                             (:433-437)

Pinned by tests/test_oracle_vs_golden.py against outputs of the real
reference function (indices captured by wrapping ``torch.searchsorted``).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def torch_sum_lastdim_f32(w: np.ndarray) -> np.ndarray:
    """fp32 sum over the last axis in ATen's CPU vectorised order."""
    w = np.ascontiguousarray(w, dtype=f32)
    n = w.shape[-1]
    lead = w.shape[:-1]
    nvec = n // 8
    if nvec == 0:
        acc = np.zeros(lead, f32)
        for j in range(n):
            acc = (acc + w[..., j]).astype(f32)
        return acc
    vecs = w[..., :nvec * 8].reshape(lead + (nvec, 8))
    full = (nvec // 4) * 4
    accs = [None] * 4
    if full > 0:
        for a in range(4):
            accs[a] = vecs[..., a, :].copy()
        for v in range(4, full):
            accs[v % 4] = (accs[v % 4] + vecs[..., v, :]).astype(f32)
        start = full
    else:
        accs[0] = vecs[..., 0, :].copy()
        start = 1
    for v in range(start, nvec):
        accs[0] = (accs[0] + vecs[..., v, :]).astype(f32)
    if full > 0:
        for a in range(1, 4):
            accs[0] = (accs[0] + accs[a]).astype(f32)
    tail = np.zeros(lead, f32)
    for j in range(nvec * 8, n):
        tail = (tail + w[..., j]).astype(f32)
    for lane in range(8):
        tail = (tail + accs[0][..., lane]).astype(f32)
    return tail


def torch_linspace01_f32(N: int) -> np.ndarray:
    """torch.linspace(0., 1., N) on CPU, fp32."""
    if N == 1:
        return np.zeros(1, f32)
    step = f32(f32(1.0) / f32(N - 1))
    out = np.empty(N, f32)
    half = N // 2
    for i in range(N):
        if i < half:
            out[i] = f32(step * f32(i))
        else:
            # single-rounded fma(-step, N-1-i, 1): exact in fp64 (24x24-bit
            # product plus 1.0 fits in 53 bits), then one rounding to fp32
            out[i] = f32(np.float64(1.0) - np.float64(step) * np.float64(N - 1 - i))
    return out


def sample_pdf_np(bins: np.ndarray, weights: np.ndarray, N_samples: int, u: np.ndarray | None = None):
    """Returns (samples f32 [R,N], inds int64 [R,N], cdf f32 [R,M])."""
    bins = np.ascontiguousarray(bins, f32)
    w = (np.ascontiguousarray(weights, f32) + f32(1e-5)).astype(f32)
    S = torch_sum_lastdim_f32(w)
    pdf = (w / S[..., None]).astype(f32)
    cdf_tail = np.cumsum(pdf.astype(np.float64), axis=-1).astype(f32)
    cdf = np.concatenate([np.zeros_like(cdf_tail[..., :1]), cdf_tail], -1)
    R, M = cdf.shape
    if u is None:
        u = np.broadcast_to(torch_linspace01_f32(N_samples), (R, N_samples))
    u = np.ascontiguousarray(u, f32)
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, M - 1)
    c_lo = np.take_along_axis(cdf, below, -1)
    c_hi = np.take_along_axis(cdf, above, -1)
    b_lo = np.take_along_axis(bins, below, -1)
    b_hi = np.take_along_axis(bins, above, -1)
    denom = (c_hi - c_lo).astype(f32)
    denom = np.where(denom < f32(1e-5), f32(1.0), denom).astype(f32)
    t = ((u - c_lo).astype(f32) / denom).astype(f32)
    samples = (b_lo + (t * (b_hi - b_lo).astype(f32)).astype(f32)).astype(f32)
    return samples, inds, cdf

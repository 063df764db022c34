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
        first 4*floor(nvec/4) vectors (each accumulator is a 4-level cascade
        flushed every 16 rows, which only matters for n >= 512), leftover
        vectors go to accumulator 0, then acc0 += acc1; acc0 += acc2;
        acc0 += acc3 (lane-wise); then a scalar tail over the n%8 trailing
        elements (sequential, starting from 0), then the 8 lanes of acc0 are
        added to it in lane order.  n < 8 takes the scalar 4-slot path. (:397)
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


def _ceil_log2(x: int) -> int:
    return 0 if x <= 1 else int(x - 1).bit_length()


def _multi_row_sum(items: np.ndarray, k: int, size_ilp: int) -> np.ndarray:
    """ATen ``multi_row_sum`` (aten/src/ATen/native/cpu/SumKernel.cpp) for ILP
    slot ``k``: rows i = 0..size_ilp-1 contribute items[..., 4*i+k, :] to a
    4-level cascade whose level 0 is flushed upwards every ``level_step`` rows."""
    levels = 4
    level_power = max(4, _ceil_log2(size_ilp) // levels)
    level_step = 1 << level_power
    level_mask = level_step - 1
    lead = items.shape[:-2] + items.shape[-1:]
    acc = [np.zeros(lead, f32) for _ in range(levels)]
    i = 0
    while i + level_step <= size_ilp:
        for _ in range(level_step):
            acc[0] = (acc[0] + items[..., 4 * i + k, :]).astype(f32)
            i += 1
        for j in range(1, levels):
            acc[j] = (acc[j] + acc[j - 1]).astype(f32)
            acc[j - 1] = np.zeros(lead, f32)
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size_ilp:
        acc[0] = (acc[0] + items[..., 4 * i + k, :]).astype(f32)
        i += 1
    for j in range(1, levels):
        acc[0] = (acc[0] + acc[j]).astype(f32)
    return acc[0]


def _row_sum(items: np.ndarray) -> np.ndarray:
    """ATen ``row_sum`` with ilp_factor 4 over items[..., size, lanes]."""
    size = items.shape[-2]
    size_ilp = size // 4
    part = [_multi_row_sum(items, k, size_ilp) for k in range(4)]
    for i in range(size_ilp * 4, size):
        part[0] = (part[0] + items[..., i, :]).astype(f32)
    for k in range(1, 4):
        part[0] = (part[0] + part[k]).astype(f32)
    return part[0]


def torch_sum_lastdim_f32(w: np.ndarray) -> np.ndarray:
    """fp32 ``torch.sum(w, -1)`` of a contiguous tensor in ATen's CPU order:
    n >= 8 -> 8-lane vectorised ``row_sum`` over the n//8 whole vectors, then a
    sequential scalar tail starting from 0, then the 8 lanes added in order;
    n < 8 -> scalar ``row_sum`` (4 ILP slots)."""
    w = np.ascontiguousarray(w, dtype=f32)
    n = w.shape[-1]
    lead = w.shape[:-1]
    if n < 8:
        return _row_sum(w[..., :, None])[..., 0]
    nvec = n // 8
    vec_acc = _row_sum(w[..., :nvec * 8].reshape(lead + (nvec, 8)))
    tail = np.zeros(lead, f32)
    for j in range(nvec * 8, n):
        tail = (tail + w[..., j]).astype(f32)
    for lane in range(8):
        tail = (tail + vec_acc[..., lane]).astype(f32)
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

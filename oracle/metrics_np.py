"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the reference's image metrics.

calculate_ssim  nerf_helpers.py:21-111   (separable Gaussian window, zero padding, sigma clamps)
calculate_metrics (mse / psnr part)  nerf_helpers.py:179-189, img2mse :8, mse2psnr :12-14

Pinned against the outputs of the unmodified reference in tests/golden/metrics.npz
(tools/make_golden_metrics.py) by tests/test_oracle_vs_golden.py.
"""
import numpy as np


def _blur(z, filt, axis):
    """1-D correlation with zero padding along `axis` (F.conv2d with padding=hw; the window is symmetric)."""
    hw = len(filt) // 2
    pad = [(0, 0)] * z.ndim
    pad[axis] = (hw, hw)
    zp = np.pad(z, pad)
    out = np.zeros_like(z)
    n = z.shape[axis]
    for k, w in enumerate(filt):
        sl = [slice(None)] * z.ndim
        sl[axis] = slice(k, k + n)
        out = out + np.float32(w) * zp[tuple(sl)]
    return out.astype(np.float32)


def calculate_ssim(img1, img2, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, return_map=False):
    a = np.clip(np.asarray(img1, np.float32), 0, max_val)          # nerf_helpers.py:55-56
    b = np.clip(np.asarray(img2, np.float32), 0, max_val)
    if a.ndim != 3:
        raise ValueError(f"Expected 3D tensor [H, W, C], got {a.shape}")
    a, b = a.transpose(2, 0, 1), b.transpose(2, 0, 1)              # [C,H,W]  (:62-63)
    hw = filter_size // 2
    shift = (2 * hw - filter_size + 1) / 2
    f_i = ((np.arange(filter_size, dtype=np.float32) - hw + shift) / np.float32(filter_sigma)) ** 2   # :70-74
    filt = np.exp(np.float32(-0.5) * f_i).astype(np.float32)
    filt = filt / filt.sum(dtype=np.float32)
    blur = lambda z: _blur(_blur(z, filt, 2), filt, 1)             # along W first, then along H (:77-86)
    mu0, mu1 = blur(a), blur(b)
    mu00, mu11, mu01 = mu0 * mu0, mu1 * mu1, mu0 * mu1
    s00 = np.maximum(blur(a * a) - mu00, 0)                        # :95-96
    s11 = np.maximum(blur(b * b) - mu11, 0)
    s01 = blur(a * b) - mu01
    s01 = np.sign(s01) * np.minimum(np.sqrt(s00 * s11), np.abs(s01))   # :97-99
    c1, c2 = np.float32((k1 * max_val) ** 2), np.float32((k2 * max_val) ** 2)
    ssim_map = ((2 * mu01 + c1) * (2 * s01 + c2)) / ((mu00 + mu11 + c1) * (s00 + s11 + c2))
    if return_map:
        return ssim_map[None].astype(np.float32)                   # [1,C,H,W]
    return float(ssim_map.astype(np.float32).mean(dtype=np.float64))


def calculate_metrics(img1, img2):
    a = np.clip(np.asarray(img1, np.float32), 0, 1)
    b = np.clip(np.asarray(img2, np.float32), 0, 1)
    mse = float(np.mean((a - b).astype(np.float32) ** 2, dtype=np.float64))
    psnr = float(-10.0 * np.log(np.float32(mse)) / np.log(np.float32(10.0)))
    return {"mse": mse, "psnr": psnr, "ssim": calculate_ssim(img1, img2)}


def synthetic_pair(H, W, seed, noise=0.08, lo=-0.1, hi=1.15):
    """Seeded image pair: a smooth pattern and a noisy, slightly shifted copy, deliberately leaving [0,1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    y, x = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32), np.linspace(0, 1, W, dtype=np.float32), indexing="ij")
    base = np.stack([0.5 + 0.5 * np.sin(9 * x + 3 * y), x * y + 0.2 * np.cos(21 * y), 0.5 + 0.6 * np.sin(14 * (x - y))], -1)
    a = np.clip(base, lo, hi).astype(np.float32)
    b = (np.roll(base, 1, axis=1) * 0.95 + rng.normal(0, noise, size=base.shape)).astype(np.float32)
    b[: H // 5, : W // 4] = a[: H // 5, : W // 4]      # an identical region (sigma01 == sqrt(s00 s11) branch)
    b[-H // 6:, -W // 6:] = 0.0                        # a flat region (zero variance)
    return a, np.clip(b, lo, hi).astype(np.float32)

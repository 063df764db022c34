"""Image-quality metrics of the reference's evaluation loop (nerf_helpers.py:21-203), same names and
signatures: ``calculate_ssim``, ``calculate_lpips``, ``calculate_metrics``.

MSE / PSNR / SSIM of a frame come from ONE fused CUDA pass over the two images (csrc/metrics.cu) -- in
``render_path`` the rendered frame is already on the GPU, so the metrics need no extra device->host copy.
LPIPS is a pretrained VGG from the optional third-party ``lpips`` package (weights are not in this image);
it is called exactly like the reference does when importable, otherwise ``calculate_metrics`` reports
``lpips: None`` as the reference does on failure (nerf_helpers.py:196-201).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops


def _as_cuda_image(img, device=None):
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img)).float()
    if not img.is_cuda:
        img = img.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    return L.f32c(img)


def image_metrics(img1, img2, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, return_map=False):
    """(out3, ssim_map): out3 = device tensor [mse, psnr, ssim] (no host sync), map [1,C,H,W] or None."""
    dev = img1.device if torch.is_tensor(img1) and img1.is_cuda else (
        img2.device if torch.is_tensor(img2) and img2.is_cuda else None)
    a, b = _as_cuda_image(img1, dev), _as_cuda_image(img2, dev)
    if a.dim() != 3:
        raise ValueError(f"Expected 3D tensor [H, W, C], got {tuple(a.shape)}")
    if a.shape != b.shape:
        raise ValueError(f"image shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    H, W, C = a.shape
    lib = L.load()
    ws = torch.empty(lib.nerfb_image_metrics_workspace_bytes(H, W) // 8, dtype=torch.float64, device=a.device)
    out = torch.empty(3, dtype=torch.float32, device=a.device)
    smap = torch.empty((1, C, H, W), dtype=torch.float32, device=a.device) if return_map else None
    with torch.cuda.device(a.device):
        L.check(lib.nerfb_image_metrics(H, W, C, L.ptr(a), L.ptr(b), float(max_val), int(filter_size),
                                        float(filter_sigma), float(k1), float(k2), L.ptr(smap), L.ptr(ws), L.ptr(out),
                                        L.stream()), "nerfb_image_metrics")
    ops._count(2)
    return out, smap


def calculate_ssim(img1, img2, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, return_map=False):
    """nerf_helpers.py:21-111.  float (mean SSIM), or the [1,C,H,W] map when ``return_map``."""
    out, smap = image_metrics(img1, img2, max_val, filter_size, filter_sigma, k1, k2, return_map)
    return smap if return_map else out[2].item()


def calculate_lpips(img1, img2, net="vgg", device="cuda", normalize=True):
    """nerf_helpers.py:113-155 (third-party pretrained network; needs the optional ``lpips`` package)."""
    import lpips
    a = _as_cuda_image(img1, device).permute(2, 0, 1).contiguous()
    b = _as_cuda_image(img2, device).permute(2, 0, 1).contiguous()
    if not hasattr(calculate_lpips, "loss_fn") or calculate_lpips.net != net:
        calculate_lpips.loss_fn = lpips.LPIPS(net=net).eval().to(device)
        calculate_lpips.net = net
    with torch.no_grad():
        return calculate_lpips.loss_fn(a, b, normalize=normalize).item()


def calculate_metrics(img1, img2, include_lpips=True, lpips_net="vgg", device="cuda"):
    """nerf_helpers.py:157-203 -> {'mse', 'psnr', 'ssim'[, 'lpips']} (python floats)."""
    out, _ = image_metrics(img1, img2)
    mse, psnr, ssim = out.tolist()
    metrics = {"mse": mse, "psnr": psnr, "ssim": ssim}
    if include_lpips:
        try:
            metrics["lpips"] = calculate_lpips(img1, img2, net=lpips_net, device=device)
        except Exception as e:  # same behaviour as the reference: warn, report None
            print(f"Warning: Could not calculate LPIPS: {e}")
            metrics["lpips"] = None
    return metrics

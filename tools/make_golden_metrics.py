#!/usr/bin/env python
"""Generate tests/golden/metrics.npz by running the UNMODIFIED reference's calculate_ssim / calculate_metrics
(/root/reference/nerf/nerf_helpers.py:21-203) on seeded image pairs.  Build container only (needs /root/reference);
inputs are re-derived from seeds by oracle.metrics_np.synthetic_pair, only outputs are stored.

    python tools/make_golden_metrics.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/nerf")

from oracle import metrics_np as MO  # noqa: E402  (input generation only)

CASES = {  # tag -> (H, W, seed, calculate_ssim kwargs)
    "a": (37, 53, 1, {}),
    "b": (64, 48, 2, dict(max_val=1.0, filter_size=7, filter_sigma=1.0)),
    "c": (16, 33, 3, dict(max_val=2.0, filter_size=11, filter_sigma=2.5, k1=0.02, k2=0.05)),
    "d": (5, 7, 4, {}),                      # image smaller than the window
    "e": (100, 100, 5, {}),
}


def main():
    import nerf_helpers as H                 # the reference
    torch.set_num_threads(4)
    out = {}
    for tag, (h, w, seed, kw) in CASES.items():
        a, b = MO.synthetic_pair(h, w, seed)
        out[f"{tag}.ssim"] = np.float64(H.calculate_ssim(a, b, **kw))
        out[f"{tag}.map"] = H.calculate_ssim(torch.from_numpy(a), torch.from_numpy(b), return_map=True, **kw).numpy()
        m = H.calculate_metrics(a, b, include_lpips=False)
        out[f"{tag}.mse"], out[f"{tag}.psnr"], out[f"{tag}.ssim_default"] = (np.float64(m["mse"]), np.float64(m["psnr"]),
                                                                           np.float64(m["ssim"]))
    path = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

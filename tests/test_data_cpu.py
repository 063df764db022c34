"""Host-side logic of the device ray-batch producer (nerf_projects_b200/data.py) that needs no GPU."""
import numpy as np
import torch

from nerf_projects_b200.data import precrop_pixel_ids


def test_precrop_pixel_ids_match_reference_grid():
    """ipynb:1256-1262: coords = meshgrid(linspace(H//2-dH, H//2+dH-1, 2dH), linspace(W//2-dW, W//2+dW-1, 2dW))."""
    for H, W, frac in ((800, 800, 0.5), (378, 504, 0.5), (100, 60, 0.3), (7, 9, 0.9)):
        dH, dW = int(H // 2 * frac), int(W // 2 * frac)
        ref = torch.stack(torch.meshgrid(torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH),
                                         torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW), indexing="ij"), -1)
        ref = ref.reshape(-1, 2).long()
        ids = precrop_pixel_ids(H, W, frac)
        assert ids.numel() == 4 * dH * dW
        assert torch.equal(ids, ref[:, 0] * W + ref[:, 1])

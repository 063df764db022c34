"""CUDA-event timing of the small kernels beside the MLP (GPU box only): keyed random draws of one training step, the fused
SSIM/MSE/PSNR pass on an 800x800 frame, against their algorithmic bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_projects_b200 as nb  # noqa: E402
from nerf_projects_b200 import metrics, ops, rng  # noqa: E402

dev = "cuda"


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


for R in (4096, 32768):
    with nb.keyed_rng(1):
        us = timeit(lambda: rng.render_rays_draws(R, 64, 128, True, 1.0, dev))
    nbytes = R * (64 + 128 + 64 + 192) * 4
    us_t = timeit(lambda: (torch.rand((R, 64), device=dev), torch.rand((R, 128), device=dev),
                           torch.randn((R, 64), device=dev) * 1.0, torch.randn((R, 192), device=dev) * 1.0))
    print(f"keyed draws R={R}: {us:7.1f} us  ({nbytes / us / 1e3:7.1f} GB/s written)   torch.rand/randn x4: {us_t:7.1f} us")
for H, W in ((800, 800), (756, 1008)):
    a, b = torch.rand(H, W, 3, device=dev), torch.rand(H, W, 3, device=dev)
    us = timeit(lambda: metrics.image_metrics(a, b))
    print(f"image_metrics {H}x{W}: {us:7.1f} us  ({H * W * 3 * 8 / us / 1e3:7.1f} GB/s algorithmic read)")

# compositing: one 32768-ray chunk, coarse (64) and fine (192) sample counts; algorithmic bytes = raw + z in (+ weights out),
# backward = raw + z + noise in, d_raw out
R = 32768
for S in (64, 192):
    raw = torch.randn(R, S, 4, device=dev)
    z = (torch.rand(R, S, device=dev) * 4 + 2).sort(-1)[0]
    d = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=-1)
    noise = torch.randn(R, S, device=dev)
    g = torch.randn(R, 3, device=dev)
    us_f = timeit(lambda: ops.raw2outputs_fwd(raw, z, d, None, True, want_weights=(S == 64)), 30)
    us_b = timeit(lambda: ops.raw2outputs_bwd(raw, z, d, noise, True, g, None, None, None, None), 30)
    bf = R * S * (16 + 4 + (4 if S == 64 else 0))
    bb = R * S * (16 + 4 + 4 + 16)
    print(f"raw2outputs S={S}: fwd {us_f:6.1f} us ({bf / us_f / 1e6:5.2f} TB/s)   bwd {us_b:6.1f} us ({bb / us_b / 1e6:5.2f} TB/s)")

# hierarchical resampling of one 32768-ray chunk: deterministic (merge path) and random u (sort of the new samples + merge)
R, S, N = 32768, 64, 128
z = (torch.rand(R, S, device=dev) * 4 + 2).sort(-1)[0]
w = torch.rand(R, S, device=dev) ** 5
u = torch.rand(R, N, device=dev)
us_d = timeit(lambda: ops.hierarchical_z(z, w, N, True, None), 30)
us_r = timeit(lambda: ops.hierarchical_z(z, w, N, False, u), 30)
print(f"hierarchical_z 32768 x (64+128): deterministic {us_d:6.1f} us   random u {us_r:6.1f} us")

"""Warm CUDA-event timing of the small kernels of one inference chunk (GPU box only): the coarse network pass with in-kernel
compositing produces the segment records and partial weights; `resample_records` (coarse maps + sample_pdf + merge) and
`composite_finish` are then timed alone on those real inputs.

    python tools/time_resample.py [rays]            (ncu: add --once to launch each kernel a single time)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_projects_b200 as nb  # noqa: E402
from nerf_projects_b200 import mlp, ops  # noqa: E402
from nerf_projects_b200 import synthetic as SY  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32560
once = "--once" in sys.argv
S, N = 64, 128
dev = torch.device("cuda")
net = nb.NeRF(8, 256, 63, 27, 5, [4], True).to(dev)
net.load_state_dict(SY.random_init_state_dict(0, use_viewdirs=True, output_ch=5))
K, c2w = SY.blender_shaped_camera(800, 800, 30.0)
ro, rd = nb.get_rays(800, 800, K, torch.from_numpy(c2w).to(dev))
ro, rd = ro.reshape(-1, 3)[:R].contiguous(), rd.reshape(-1, 3)[:R].contiguous()
vd = torch.nn.functional.normalize(rd, dim=-1)
rays = torch.cat([ro, rd, torch.full((R, 1), 2.0, device=dev), torch.full((R, 1), 6.0, device=dev), vd], -1).contiguous()
packed = mlp.packed_weights(net)
rec, wpart = ops.mlp_fwd_composite_bf16(R, S, rays, None, False, True, 5, packed, want_wpart=True)
torch.cuda.synchronize()


def timeit(fn, reps=50):
    if once:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


us = timeit(lambda: ops.resample_records(R, S, N, rays, None, False, wpart, rec, True, True, None))
nbytes = R * (S * 4 + (S // 32) * 32 + 44 + (S + N) * 4 + 28)
print(f"resample_records R={R}: {us:7.1f} us   ({nbytes / 1e6:.1f} MB algorithmic, {nbytes / max(us, 1e-9) / 1e6:5.2f} TB/s)")
u = torch.rand(R, N, device=dev)
us = timeit(lambda: ops.resample_records(R, S, N, rays, None, False, wpart, rec, True, False, u))
print(f"resample_records R={R} random u: {us:7.1f} us")
out = ops.resample_records(R, S, N, rays, None, False, wpart, rec, True, True, None)
rec_f, _ = ops.mlp_fwd_composite_bf16(R, S + N, rays, out[4], False, True, 5, packed, want_wpart=False)
us = timeit(lambda: ops.composite_finish(rec_f, True))
print(f"composite_finish R={R}: {us:7.1f} us")

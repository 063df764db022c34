"""Diagnostic (GPU box only): runs the stages of one training forward+backward twice with identical inputs and seeds --
once on fresh memory, once after the caching allocator was polluted with NaN / huge floats -- and reports the first
stage whose output changes (a read of uninitialised memory)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_projects_b200 as nb  # noqa: E402
from nerf_projects_b200 import mlp, ops  # noqa: E402
from nerf_projects_b200 import synthetic as SY  # noqa: E402

DEV = "cuda"
R, S, N_imp = 1024, 64, 128
net_c = nb.NeRF(8, 256, 63, 27, 5, [4], True).to(DEV)
net_f = nb.NeRF(8, 256, 63, 27, 5, [4], True).to(DEV)
net_c.load_state_dict(SY.random_init_state_dict(0, use_viewdirs=True, output_ch=5))
net_f.load_state_dict(SY.random_init_state_dict(1, use_viewdirs=True, output_ch=5))
gen = torch.Generator().manual_seed(0)
rays = torch.cat([torch.randn(R, 3, generator=gen) * 0.1 + torch.tensor([0., 0., 4.]), torch.nn.functional.normalize(
    torch.randn(R, 3, generator=gen) * 0.2 + torch.tensor([0., 0., -1.]), dim=-1), torch.full((R, 1), 2.0),
    torch.full((R, 1), 6.0), torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=-1)], -1).to(DEV)
t_rand = torch.rand(R, S, generator=gen).to(DEV)
u = torch.rand(R, N_imp, generator=gen).to(DEV)
target = torch.rand(R, 3, generator=gen).to(DEV)


def pollute(val):
    bufs = [torch.full((1 << 28,), val, dtype=torch.float32, device=DEV) for _ in range(20)]
    del bufs


def run():
    out = {}
    z_c = ops.coarse_z(rays, S, False, t_rand); out["z_c"] = z_c
    raw_c, saved_c = mlp.forward_save_bf16(net_c, R, S, rays, z_c); out["raw_c"] = raw_c
    out["acts_c"], out["masks_c"] = saved_c
    rgb0, disp0, acc0, w_c, depth0 = ops.raw2outputs_fwd(raw_c.view(R, S, -1), z_c, rays[:, 3:6], None, True, want_weights=True)
    out.update(rgb0=rgb0, disp0=disp0, acc0=acc0, w_c=w_c)
    z_f, z_s, z_std, _ = ops.hierarchical_z(z_c, w_c, N_imp, False, u); out.update(z_f=z_f, z_s=z_s, z_std=z_std)
    raw_f, saved_f = mlp.forward_save_bf16(net_f, R, S + N_imp, rays, z_f); out["raw_f"] = raw_f
    out["acts_f"], out["masks_f"] = saved_f
    rgb, disp, acc, w_f, depth = ops.raw2outputs_fwd(raw_f.view(R, S + N_imp, -1), z_f, rays[:, 3:6], None, True)
    out.update(rgb=rgb, disp=disp, acc=acc)
    d_rgb = 2 * (rgb - target) / rgb.numel()
    g_raw_f = ops.raw2outputs_bwd(raw_f.view(R, S + N_imp, -1), z_f, rays[:, 3:6], None, True, d_rgb, None, None, None, None)
    out["g_raw_f"] = g_raw_f
    grads = mlp.backward_from_saved(net_f, R * (S + N_imp), saved_f, g_raw_f)
    for k, v in grads.items():
        out["grad." + k] = v
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in out.items() if v is not None}


a = run()
for val, name in ((float("nan"), "nan"), (3.0e38, "big"), (-1.0, "neg")):
    pollute(val)
    b = run()
    diffs = []
    for k in a:
        x, y = a[k], b[k]
        same = torch.equal(x, y) if x.dtype == torch.uint8 else torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))
        if not same:
            d = (x.float() - y.float()).abs()
            diffs.append(f"{k}: max|d|={float(torch.nan_to_num(d, nan=9e9).max()):.3e} n={int((d != 0).sum() + torch.isnan(d).sum())}")
    print(f"pollute={name}: " + ("; ".join(diffs) if diffs else "all stages bit-identical"))

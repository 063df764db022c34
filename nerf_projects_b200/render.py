"""Drop-in replacements for the render cells of /root/reference/nerf/nerf.ipynb
(same names, signatures, defaults and return conventions):

    batchify, run_network, raw2outputs, render_rays, batchify_rays, render,
    render_path, create_nerf                      (ipynb:224-961)
    get_rays, get_rays_np, ndc_rays, sample_pdf, img2mse, mse2psnr, to8b
                                                  (nerf_helpers.py)

All numerical work runs in the CUDA kernels of libnerf_b200.so.  When
``render_rays`` receives the networks/closure built by this module's
``create_nerf`` it takes the fused path (autograd.RenderRaysFn: no positions,
encodings or activations are materialised in HBM); with arbitrary callables it
follows the reference's op-by-op structure using the same kernels as
differentiable ops.
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from . import autograd as AG
from . import mlp, ops
from . import metrics as M
from . import rng as RNG
from .embedder import get_embedder
from .model import NeRF

DEBUG = False

# ----------------------------------------------------------------- helpers

img2mse = lambda x, y: torch.mean((x - y) ** 2)          # nerf_helpers.py:8


def mse2psnr(x: torch.Tensor) -> torch.Tensor:            # nerf_helpers.py:12-14
    return -10.0 * torch.log(x) / torch.log(torch.tensor(10.0, device=x.device, dtype=x.dtype))


to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)  # nerf_helpers.py:17


def _device_of(kwargs):
    return next(kwargs["network_fn"].parameters()).device


def _as_cuda(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t), dtype=torch.float32)
    return t.to(device=device, dtype=torch.float32)


def get_rays(H, W, K, c2w):
    """nerf_helpers.py:222-296.  Returns (rays_o, rays_d), each [H, W, 3]."""
    if not torch.is_tensor(c2w):
        c2w = torch.as_tensor(np.asarray(c2w), dtype=torch.float32)
    if not c2w.is_cuda:
        c2w = c2w.cuda()
    with L.on_device(c2w.device):
        return ops.get_rays(int(H), int(W), K[0][0], K[1][1], K[0][2], K[1][2], c2w)


def get_rays_np(H, W, K, c2w):
    """nerf_helpers.py:301-308 (numpy in, numpy out)."""
    ro, rd = get_rays(H, W, K, torch.as_tensor(np.asarray(c2w), dtype=torch.float32))
    return ro.cpu().numpy(), rd.cpu().numpy()


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """nerf_helpers.py:311-369."""
    return ops.ndc_rays(H, W, focal, near, rays_o, rays_d)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """nerf_helpers.py:372-439.  Bit-exact indices/samples w.r.t. the
    reference's CPU fp32 result for det=True; for det=False the uniform draws
    come from torch.rand on the tensors' device exactly like the reference."""
    lead = bins.shape[:-1]
    b2 = L.f32c(bins).reshape(-1, bins.shape[-1])
    w2 = weights.reshape(-1, weights.shape[-1]) if weights.is_contiguous() or weights.dim() == 2 else \
        weights.reshape(-1, weights.shape[-1])
    if w2.dtype != torch.float32:
        w2 = w2.float()
    if w2.stride(-1) != 1:
        w2 = w2.contiguous()
    u = None
    if pytest:
        np.random.seed(0)
        shape = list(lead) + [N_samples]
        un = np.broadcast_to(np.linspace(0.0, 1.0, N_samples), shape) if det else np.random.rand(*shape)
        u = torch.tensor(np.ascontiguousarray(un), dtype=torch.float32, device=bins.device).reshape(-1, N_samples)
        det = False
    elif not det:
        u = torch.rand(list(lead) + [N_samples], device=bins.device, dtype=torch.float32).reshape(-1, N_samples)
    out = ops.sample_pdf(b2, w2, int(N_samples), det=det, u=u)
    return out.reshape(*lead, N_samples)


def batchify(fn, chunk):
    """ipynb:224-244."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """ipynb:790-855.  With this package's NeRF module and embedders the whole
    thing is one fused MLP evaluation over the sample positions; otherwise the
    reference's embed -> cat -> batchify(fn) sequence is followed."""
    fused = isinstance(fn, NeRF) and hasattr(embed_fn, "multires") and \
        (viewdirs is None or hasattr(embeddirs_fn, "multires")) and not torch.is_grad_enabled()
    if fused and _fusable(fn, embed_fn, embeddirs_fn if viewdirs is not None else None):
        lead = inputs.shape[:-1]
        S = inputs.shape[-2] if inputs.dim() >= 2 else 1
        pts = L.f32c(inputs).reshape(-1, 3)
        R = pts.shape[0] // S
        vd = L.f32c(viewdirs).reshape(-1, 3) if (viewdirs is not None and fn.use_viewdirs) else None
        raw = mlp.forward(fn, R, S, None, None, pts=pts, viewdirs=vd)
        return raw.reshape(*lead, raw.shape[-1])
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded = torch.cat([embedded, embeddirs_fn(input_dirs_flat)], -1)
    embedded = embedded.to(next(fn.parameters()).device)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


def _fusable(net, embed_fn, embeddirs_fn) -> bool:
    if not isinstance(net, NeRF):
        return False
    if ops.enc_dim(getattr(embed_fn, "multires", -99)) != net.input_ch:
        return False
    if net.use_viewdirs:
        if embeddirs_fn is None or ops.enc_dim(getattr(embeddirs_fn, "multires", -99)) != net.input_ch_views:
            return False
    return True


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """ipynb:254-349 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    try:
        noise_std = float(raw_noise_std)
    except (TypeError, ValueError):
        noise_std = 0.0
    noise = None
    if noise_std > 0.0:
        noise = torch.randn(raw[..., 3].shape, device=raw.device, dtype=torch.float32) * noise_std
        if pytest:
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * noise_std, device=raw.device,
                                 dtype=torch.float32)
    return AG.Raw2OutputsFn.apply(raw, z_vals, rays_d, noise, bool(white_bkgd))


# ------------------------------------------------------------ render_rays

def _noise_std(v) -> float:
    try:
        return float(v)
    except (TypeError, ValueError):
        return 0.0


def _torch_draws(N_rays, N_samples, N_importance, perturb_on, std, dev, pytest):
    """The reference's own draws (torch global generator; np.random.seed(0) under `pytest`)."""
    t_rand = u = noise_c = noise_f = None
    if perturb_on:
        t_rand = torch.rand((N_rays, N_samples), device=dev)
        if pytest:
            np.random.seed(0)
            t_rand = torch.tensor(np.random.rand(N_rays, N_samples), device=dev, dtype=torch.float32)

    def mk_noise(S):
        if std <= 0.0:
            return None
        if pytest:
            np.random.seed(0)
            return torch.tensor(np.random.rand(N_rays, S) * std, device=dev, dtype=torch.float32)
        return torch.randn((N_rays, S), device=dev) * std
    noise_c = mk_noise(N_samples)
    if N_importance > 0:
        if perturb_on:
            if pytest:
                np.random.seed(0)
                u = torch.tensor(np.random.rand(N_rays, N_importance), device=dev, dtype=torch.float32)
            else:
                u = torch.rand((N_rays, N_importance), device=dev)
        noise_f = mk_noise(N_samples + N_importance)
    return t_rand, u, noise_c, noise_f


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False):
    """ipynb:359-492.  Returns the reference's dict (rgb_map, disp_map, acc_map,
    [raw], [rgb0, disp0, acc0, z_std]) plus additive keys depth_map / depth0."""
    ray_batch = L.f32c(ray_batch)
    L.require_cuda(ray_batch)
    N_rays = ray_batch.shape[0]
    spec = getattr(network_query_fn, "_nerfb", None)
    fused = spec is not None and _fusable(network_fn, spec["embed_fn"], spec["embeddirs_fn"]) and \
        (network_fine is None or _fusable(network_fine, spec["embed_fn"], spec["embeddirs_fn"])) and \
        (ray_batch.shape[-1] > 8) == bool(network_fn.use_viewdirs)
    if not fused:
        return _render_rays_generic(ray_batch, network_fn, network_query_fn, N_samples, retraw, lindisp, perturb,
                                    N_importance, network_fine, white_bkgd, raw_noise_std, pytest)
    dev = ray_batch.device
    perturb_on = bool(perturb > 0.)
    std = _noise_std(raw_noise_std)
    if RNG.current() is not None and not pytest:
        # draws keyed by (seed, global ray index): independent of chunking / sharding (rng.py)
        t_rand, u, noise_c, noise_f = RNG.render_rays_draws(N_rays, N_samples, N_importance, perturb_on, std, dev)
    else:
        t_rand, u, noise_c, noise_f = _torch_draws(N_rays, N_samples, N_importance, perturb_on, std, dev, pytest)
    cfg = dict(network_fn=network_fn, network_fine=network_fine, N_samples=int(N_samples),
               N_importance=int(N_importance), lindisp=bool(lindisp), white_bkgd=bool(white_bkgd))
    params = list(network_fn.parameters())
    if network_fine is not None:
        params += list(network_fine.parameters())
    cfg["train"] = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if not cfg["train"] and not retraw and noise_c is None and noise_f is None and _composite_in_kernel(cfg):
        return _render_rays_inference(ray_batch, cfg, t_rand, u)
    (rgb, disp, acc, depth, weights, raw, z_vals, rgb0, disp0, acc0, depth0, z_std, z_samples) = \
        AG.RenderRaysFn.apply(cfg, ray_batch, t_rand, u, noise_c, noise_f, *params)
    ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc}
    if retraw:
        ret["raw"] = raw
    if N_importance > 0:
        ret["rgb0"], ret["disp0"], ret["acc0"], ret["z_std"] = rgb0, disp0, acc0, z_std
        ret["depth0"] = depth0
    ret["depth_map"] = depth
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def _composite_in_kernel(cfg) -> bool:
    """Inference in bf16 mode with the canonical networks and whole 32-sample segments per ray: the compositing runs
    inside the fused MLP kernel (nerf_mlp_kernel<FWD_COMP>) and the per-sample raw outputs never reach HBM."""
    S, N = cfg["N_samples"], cfg["N_importance"]
    nets = [cfg["network_fn"]] + ([cfg["network_fine"]] if cfg["network_fine"] is not None else [])
    if mlp.get_precision() != "bf16" or not all(n.is_canonical() and (n.use_viewdirs or n.output_ch in (4, 5)) for n in nets):
        return False
    if S % 32 != 0 or S > 1024 or (N > 0 and ((S + N) % 32 != 0 or S + N > 1024 or S < 3)):
        return False
    return os.environ.get("NERFB200_COMPOSITE_IN_KERNEL", "1") != "0"


def _render_rays_inference(rays, cfg, t_rand, u):
    """render_rays (ipynb:359-492) for inference: coarse network pass with the compositing of every 32-sample segment
    inside the kernel -> coarse maps + hierarchical resampling from the segment records (coarse weights rebuilt on-chip)
    -> fine network pass + compositing -> finish of the rays from the records (1 byte per sample).  Coarse depths are
    computed inside the kernels (written to HBM only when perturb > 0 supplies jitter); raw is never written."""
    net_c, net_f = cfg["network_fn"], cfg["network_fine"]
    R, S, N = rays.shape[0], cfg["N_samples"], cfg["N_importance"]
    white, lindisp = cfg["white_bkgd"], cfg["lindisp"]
    z_c = ops.coarse_z(rays, S, lindisp, t_rand) if t_rand is not None else None
    vd = bool(net_c.use_viewdirs)
    rec, wpart = ops.mlp_fwd_composite_bf16(R, S, rays, z_c, lindisp, vd, net_c.output_ch, mlp.packed_weights(net_c),
                                            want_wpart=N > 0)
    if N == 0:
        rgb, disp, acc, depth = ops.composite_finish(rec, white)
        return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth}
    det = (t_rand is None and u is None)
    rgb0, disp0, acc0, depth0, z_f, _, z_std = ops.resample_records(R, S, N, rays, z_c, lindisp, wpart, rec, white, det, u)
    run = net_f if net_f is not None else net_c
    rec_f, _ = ops.mlp_fwd_composite_bf16(R, S + N, rays, z_f, False, bool(run.use_viewdirs), run.output_ch,
                                          mlp.packed_weights(run), want_wpart=False)
    rgb, disp, acc, depth = ops.composite_finish(rec_f, white)
    return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "rgb0": rgb0, "disp0": disp0, "acc0": acc0, "z_std": z_std,
            "depth0": depth0, "depth_map": depth}


def _render_rays_generic(ray_batch, network_fn, network_query_fn, N_samples, retraw, lindisp, perturb,
                         N_importance, network_fine, white_bkgd, raw_noise_std, pytest):
    """Reference structure with arbitrary ``network_query_fn`` (ipynb:412-486)."""
    N_rays = ray_batch.shape[0]
    dev = ray_batch.device
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    keyed = RNG.current() is not None and not pytest
    if keyed:
        t_rand, u, noise_c, noise_f = RNG.render_rays_draws(N_rays, N_samples, N_importance, bool(perturb > 0.),
                                                            _noise_std(raw_noise_std), dev)
    else:
        # the reference's order of draws: t_rand, coarse noise (inside raw2outputs), u, fine noise
        t_rand = _torch_draws(N_rays, N_samples, 0, bool(perturb > 0.), 0.0, dev, pytest)[0]

    def composite(raw, z, noise):
        if keyed:
            return AG.Raw2OutputsFn.apply(raw, z, rays_d, noise, bool(white_bkgd))
        return raw2outputs(raw, z, rays_d, raw_noise_std, white_bkgd, pytest=pytest)
    z_vals = ops.coarse_z(ray_batch, N_samples, lindisp, t_rand)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    raw = network_query_fn(pts, viewdirs, network_fn)
    rgb_map, disp_map, acc_map, weights, depth_map = composite(raw, z_vals, noise_c if keyed else None)
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0, depth_0 = rgb_map, disp_map, acc_map, depth_map
        if not keyed:
            u = None
            if perturb > 0.:
                u = torch.rand((N_rays, N_importance), device=dev)
                if pytest:
                    np.random.seed(0)
                    u = torch.tensor(np.random.rand(N_rays, N_importance), device=dev, dtype=torch.float32)
        z_vals, z_samples, z_std, _ = ops.hierarchical_z(z_vals, weights.detach(), N_importance, det=(u is None), u=u)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        run_fn = network_fn if network_fine is None else network_fine
        raw = network_query_fn(pts, viewdirs, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map = composite(raw, z_vals, noise_f if keyed else None)
    ret = {"rgb_map": rgb_map, "disp_map": disp_map, "acc_map": acc_map}
    if retraw:
        ret["raw"] = raw
    if N_importance > 0:
        ret.update(rgb0=rgb_map_0, disp0=disp_map_0, acc0=acc_map_0, z_std=z_std, depth0=depth_0)
    ret["depth_map"] = depth_map
    return ret


def _chunk_period(kwargs) -> int:
    """Rays after which the fused bf16 kernel's assignment of samples to tile slots repeats: the smallest r with
    r * S = 0 (mod 512 rows per super-tile) for both passes.  Cutting chunks at multiples of it makes a ray's result
    independent of `chunk` bit for bit (the two row tiles of a CTA accumulate the K chunks of the skip / view layers in
    different orders -- the kernel's only position dependence)."""
    import math
    S = int(kwargs.get("N_samples", 0) or 0)
    N = int(kwargs.get("N_importance", 0) or 0)
    if S <= 0:
        return 1
    p = 512 // math.gcd(S, 512)
    if N > 0:
        q = 512 // math.gcd(S + N, 512)
        p = p * q // math.gcd(p, q)
    return p


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """ipynb:514-548.  `chunk` is a memory cap: it is rounded down to a multiple of the tile-slot period of the fused
    kernel (8 rays for 64 + 128 samples) so that results do not depend on it, bit for bit."""
    all_ret = {}
    period = _chunk_period(kwargs)
    if chunk > period:
        chunk = chunk // period * period
    for i in range(0, rays_flat.shape[0], chunk):
        with RNG.advanced(i):      # no-op unless a keyed_rng context is active
            ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """ipynb:558-640 -> [rgb_map, disp_map, acc_map, extras].  Runs on the device of ``network_fn`` whatever the
    current CUDA device is (the reference's ``.to(model_device)`` convention)."""
    dev = _device_of(kwargs)
    with L.on_device(dev):
        return _render(dev, H, W, K, chunk, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam, kwargs)


def _render(dev, H, W, K, chunk, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam, kwargs):
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, _as_cuda(c2w, dev))
    else:
        rays_o, rays_d = rays
        rays_o, rays_d = _as_cuda(rays_o, dev), _as_cuda(rays_d, dev)
    view_d = rays_d
    if use_viewdirs and c2w_staticcam is not None:
        rays_o, rays_d = get_rays(H, W, K, _as_cuda(c2w_staticcam, dev))
    sh = rays_d.shape
    focal = K[0][0]
    if torch.is_tensor(focal):
        focal = focal.item()
    packed = ops.pack_rays(rays_o, rays_d, view_d, use_viewdirs, ndc, H, W, focal, near, far)
    all_ret = batchify_rays(packed, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ["rgb_map", "disp_map", "acc_map"]
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


def _render_frame_sharded(H, W, K, chunk, c2w, render_kwargs, slot):
    """One frame of render_path ray-sharded over all ranks of the default process group (parallel.render_sharded):
    every rank renders its slice, one all-gather; every rank gets the full [H,W,3] / [H,W] maps."""
    from . import parallel
    dev = _device_of(render_kwargs)
    rays_o, rays_d = get_rays(H, W, K, _as_cuda(c2w, dev))
    frame = parallel.render_sharded(render, H, W, K, rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), chunk=chunk, slot=slot,
                                    async_gather=True, **render_kwargs)
    rgb, disp, acc = frame.maps()
    # own storage: the gather buffer of this slot is reused two frames later, the D2H copy may still be reading
    return rgb.reshape(H, W, 3).contiguous(), disp.reshape(H, W).contiguous(), acc.reshape(H, W).contiguous()


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                calculate_metrics=False, metrics_include_lpips=True, metrics_device="cuda", sharded=None):
    """ipynb:650-758.  The device->host copy of frame k overlaps the render of
    frame k+1 (pinned buffers + a copy stream).  With ``calculate_metrics`` and
    ``gt_imgs``, MSE / PSNR / SSIM of every frame are computed on the GPU from the
    frame where it lies (metrics.image_metrics, one fused pass, no host sync inside
    the loop) and averaged like the reference: ``avg_/std_{psnr,ssim,mse[,lpips]}``.
    LPIPS needs the optional ``lpips`` package (skipped with a warning otherwise).

    ``sharded`` (default: automatically, when a torch.distributed group with more than one rank is initialised): every
    frame is ray-sharded over all ranks -- all ranks must call render_path with the same poses; every rank returns the
    full frames, only rank 0 writes files to ``savedir``."""
    import torch.distributed as dist
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    dev = _device_of(render_kwargs)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    sharded = multi if sharded is None else (bool(sharded) and multi)
    if sharded and dist.get_rank() != 0:
        savedir = None
    n = len(render_poses)
    rgbs = np.empty((n, H, W, 3), np.float32)
    disps = np.empty((n, H, W), np.float32)
    copy_stream = torch.cuda.Stream(device=dev)
    pin = [(torch.empty((H, W, 3), dtype=torch.float32).pin_memory(), torch.empty((H, W), dtype=torch.float32).pin_memory())
           for _ in range(2)]
    pending = []
    frame_metrics = []          # device tensors [mse, psnr, ssim], read back once after the loop
    lpips_vals = []

    def drain(upto):
        while len(pending) > upto:
            i, ev, slot = pending.pop(0)
            ev.synchronize()
            rgbs[i] = pin[slot][0].numpy()
            disps[i] = pin[slot][1].numpy()
            if savedir is not None:
                try:
                    import imageio
                    imageio.imwrite(os.path.join(savedir, "{:03d}.png".format(i)), to8b(rgbs[i]))
                except ImportError:
                    np.save(os.path.join(savedir, "{:03d}.npy".format(i)), to8b(rgbs[i]))
    t = time.time()
    for i, c2w in enumerate(render_poses):
        t = time.time()
        if sharded:
            rgb, disp, acc = _render_frame_sharded(H, W, K, chunk, c2w[:3, :4], render_kwargs, slot=i % 2)
        else:
            rgb, disp, acc, _ = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)
        if calculate_metrics and gt_imgs is not None and render_factor == 0 and i < len(gt_imgs):
            gt = torch.as_tensor(np.asarray(gt_imgs[i]), dtype=torch.float32, device=dev)[..., :3]
            frame_metrics.append(M.image_metrics(rgb, gt)[0])
            if metrics_include_lpips:
                try:
                    lpips_vals.append(M.calculate_lpips(rgb, gt, device=metrics_device))
                except Exception as e:
                    if not lpips_vals and i == 0:
                        print(f"Warning: Could not calculate LPIPS: {e}")
        drain(1)
        slot = i % 2
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            pin[slot][0].copy_(rgb, non_blocking=True)
            pin[slot][1].copy_(disp, non_blocking=True)
            rgb.record_stream(copy_stream)
            disp.record_stream(copy_stream)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending.append((i, ev, slot))
    drain(0)
    if calculate_metrics and gt_imgs is not None:
        avg = {}
        if frame_metrics:
            m = torch.stack(frame_metrics).double().cpu().numpy()      # [frames, 3] = mse, psnr, ssim
            for j, key in enumerate(("mse", "psnr", "ssim")):
                avg[f"avg_{key}"] = float(np.mean(m[:, j]))
                avg[f"std_{key}"] = float(np.std(m[:, j]))
        if lpips_vals:
            avg["avg_lpips"], avg["std_lpips"] = float(np.mean(lpips_vals)), float(np.std(lpips_vals))
        return rgbs, disps, avg
    return rgbs, disps


# ------------------------------------------------------------ create_nerf

class _QueryFn:
    """network_query_fn of create_nerf (ipynb:899-902) with its configuration
    attached so render_rays can take the fused path."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self._nerfb = dict(embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=netchunk)

    def __call__(self, inputs, viewdirs, network_fn):
        s = self._nerfb
        return run_network(inputs, viewdirs, network_fn, embed_fn=s["embed_fn"], embeddirs_fn=s["embeddirs_fn"],
                           netchunk=s["netchunk"])


def create_nerf(args, device: Optional[torch.device] = None):
    """ipynb:876-961.  Same `args` attributes, same 5-tuple.  ``device``
    defaults to the current CUDA device (the notebook's global ``device``)."""
    from .optim import FusedAdam
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views = 0
    embeddirs_fn = None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                 input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)
        grad_vars += list(model_fine.parameters())
    network_query_fn = _QueryFn(embed_fn, embeddirs_fn, args.netchunk)
    optimizer = FusedAdam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    basedir, expname = args.basedir, args.expname
    checkpoint_dir = os.path.join(basedir, expname, "checkpoints")
    os.makedirs(checkpoint_dir, exist_ok=True)
    if args.ft_path is not None and args.ft_path != "None":
        ckpts = [args.ft_path]
    else:
        ckpts = [os.path.join(checkpoint_dir, f) for f in sorted(os.listdir(checkpoint_dir)) if "tar" in f]
    print("Found ckpts", ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        print("Reloading from", ckpt_path)
        ckpt = torch.load(ckpt_path, map_location=device)
        start = ckpt["global_step"]
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        model.load_state_dict(ckpt["network_fn_state_dict"])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt["network_fine_state_dict"])

    render_kwargs_train = {
        "network_query_fn": network_query_fn,
        "perturb": args.perturb,
        "N_importance": args.N_importance,
        "network_fine": model_fine,
        "N_samples": args.N_samples,
        "network_fn": model,
        "use_viewdirs": args.use_viewdirs,
        "white_bkgd": args.white_bkgd,
        "raw_noise_std": args.raw_noise_std,
    }
    if args.dataset_type != "llff" or args.no_ndc:
        print("Not ndc!")
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = args.lindisp
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test["perturb"] = False
    render_kwargs_test["raw_noise_std"] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer

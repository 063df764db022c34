"""torch.autograd glue: the CUDA kernels are exposed as differentiable ops so
the reference's training loop (``loss.backward(); optimizer.step()``,
ipynb:1274-1275) keeps working unchanged."""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function

from . import _lib as L
from . import mlp, ops


class LinearFn(Function):
    """y = act(x W^T + b) with the fp32 kernels (used by NeRF.forward)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool):
        L.require_cuda(x, weight, bias)
        x = L.f32c(x)
        y = torch.empty((x.shape[0], weight.shape[0]), dtype=torch.float32, device=x.device)
        ops.linear_fwd(x, weight, bias, y, relu)
        ctx.relu = relu
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = L.f32c(dy)
        if ctx.relu:
            dy = dy * (y > 0)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.linear_bwd_data(dy, weight, dx)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw = torch.zeros_like(weight)
            db = torch.zeros(weight.shape[0], dtype=torch.float32, device=x.device)
            ops.linear_bwd_weight(dy, x, dw, db)
        return dx, dw, db, None


class Raw2OutputsFn(Function):
    """raw2outputs (ipynb:254-349), differentiable w.r.t. raw."""

    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise, white_bkgd: bool):
        raw, z_vals, rays_d = L.f32c(raw), L.f32c(z_vals), L.f32c(rays_d)
        ctx.set_materialize_grads(False)      # unused outputs arrive as None in backward, not as zero-filled tensors
        rgb, disp, acc, weights, depth = ops.raw2outputs_fwd(raw, z_vals, rays_d, noise, white_bkgd)
        ctx.white = white_bkgd
        ctx.save_for_backward(raw, z_vals, rays_d, noise)
        return rgb, disp, acc, weights, depth

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_weights, d_depth):
        raw, z_vals, rays_d, noise = ctx.saved_tensors
        d_raw = ops.raw2outputs_bwd(raw, z_vals, rays_d, noise, ctx.white, d_rgb, d_disp, d_acc, d_weights, d_depth)
        return d_raw, None, None, None, None


def _backward_coarse(ctx, net_c, R, S, rays, z_c, g_raw_c):
    if ctx.acts_c is not None:
        grads = mlp.backward_from_saved(net_c, R * S, ctx.acts_c, g_raw_c)
        ctx.acts_c = None
        return grads
    return mlp.backward(net_c, R, S, rays, z_c, None, None, g_raw_c.view(R * S, -1))


class RenderRaysFn(Function):
    """The fused render_rays (ipynb:359-492): coarse pass, compositing,
    hierarchical sampling, fine pass, compositing -- differentiable w.r.t. the
    parameters of both networks.  Forward stores rays, depths, the 4-channel raw
    outputs and -- bf16 mode, when they fit mlp._save_budget_bytes() -- the bf16 layer
    inputs of each pass for the tensor-core backward; otherwise the backward
    recomputes the MLP activations chunk by chunk.

    apply(cfg, rays, t_rand, u, noise_c, noise_f, n_coarse_params, *params)
    returns (rgb, disp, acc, depth, <empty: final weights are not materialised>, raw, z_vals,
             rgb0, disp0, acc0, depth0, z_std, z_samples)
    """

    @staticmethod
    def forward(ctx, cfg, rays, t_rand, u, noise_c, noise_f, *params):
        net_c, net_f = cfg["network_fn"], cfg["network_fine"]
        R = rays.shape[0]
        S, N_imp = cfg["N_samples"], cfg["N_importance"]
        white = cfg["white_bkgd"]
        rays_d = rays[:, 3:6]
        # outputs the loss does not use (disp, acc, depth, raw under retraw=True, ...) arrive in backward as None instead
        # of freshly zero-filled tensors: a dozen fills and one [R, S, 4] add fewer per training step
        ctx.set_materialize_grads(False)
        z_c = ops.coarse_z(rays, S, cfg["lindisp"], t_rand)
        # training: keep the layer inputs of each pass for the backward when they fit (else recompute there)
        # (cfg["train"] is decided by the caller, where grad mode is visible: inside Function.forward grad mode is
        #  always off and ctx.needs_input_grad is set even under torch.no_grad())
        n_cp = len(list(net_c.parameters()))
        train = bool(cfg.get("train", False))
        train_c = train and any(ctx.needs_input_grad[6:6 + n_cp])
        train_f = train and any(ctx.needs_input_grad[6 + (n_cp if net_f is not None else 0):])
        ctx.acts_c = ctx.acts_f = None
        if train_c and mlp.can_save_forward(net_c, R * S):
            raw_c, ctx.acts_c = mlp.forward_save_bf16(net_c, R, S, rays, z_c)
            raw_c = raw_c.view(R, S, -1)
        else:
            raw_c = mlp.forward(net_c, R, S, rays, z_c).view(R, S, -1)
        ctx.cfg = cfg
        ctx.R = R
        empty = rays.new_empty(0)
        if N_imp > 0:
            # coarse compositing + z_mid / sample_pdf / sort / z_std in one launch: the coarse weights stay on-chip
            det = (t_rand is None and u is None)
            rgb, disp, acc, depth, _, z_f, z_samples, z_std = ops.composite_resample(raw_c, z_c, rays_d, noise_c, white,
                                                                                     N_imp, det, u)
        else:
            # (the per-sample weights of the final pass are not part of render_rays' result: not written to HBM)
            rgb, disp, acc, _, depth = ops.raw2outputs_fwd(raw_c, z_c, rays_d, noise_c, white, want_weights=False)
        if N_imp > 0:
            run = net_f if net_f is not None else net_c
            if train_f and mlp.can_save_forward(run, R * (S + N_imp)):
                raw_f, ctx.acts_f = mlp.forward_save_bf16(run, R, S + N_imp, rays, z_f)
                raw_f = raw_f.view(R, S + N_imp, -1)
            else:
                raw_f = mlp.forward(run, R, S + N_imp, rays, z_f).view(R, S + N_imp, -1)
            rgb_f, disp_f, acc_f, _, depth_f = ops.raw2outputs_fwd(raw_f, z_f, rays_d, noise_f, white, want_weights=False)
            ctx.save_for_backward(rays, z_c, raw_c, z_f, raw_f, noise_c if noise_c is not None else empty,
                                  noise_f if noise_f is not None else empty)
            outs = (rgb_f, disp_f, acc_f, depth_f, empty, raw_f, z_f, rgb, disp, acc, depth, z_std, z_samples)
            ctx.mark_non_differentiable(z_f, z_std, z_samples)
        else:
            ctx.save_for_backward(rays, z_c, raw_c, empty, empty, noise_c if noise_c is not None else empty, empty)
            outs = (rgb, disp, acc, depth, empty, raw_c, z_c, empty, empty, empty, empty, empty, empty)
            ctx.mark_non_differentiable(z_c)
        return outs

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_depth, d_w, d_raw, d_z, d_rgb0, d_disp0, d_acc0, d_depth0, d_zstd,
                 d_zs):
        cfg = ctx.cfg
        net_c, net_f = cfg["network_fn"], cfg["network_fine"]
        rays, z_c, raw_c, z_f, raw_f, noise_c, noise_f = ctx.saved_tensors
        noise_c = noise_c if noise_c.numel() else None
        noise_f = noise_f if noise_f.numel() else None
        R, S, N_imp, white = ctx.R, cfg["N_samples"], cfg["N_importance"], cfg["white_bkgd"]
        rays_d = rays[:, 3:6]
        named_c = [n for n, _ in net_c.named_parameters()]
        grads_c = grads_f = None
        if N_imp > 0:
            g_raw_f = ops.raw2outputs_bwd(raw_f, z_f, rays_d, noise_f, white, d_rgb, d_disp, d_acc, None, d_depth)
            if d_raw is not None:
                g_raw_f = g_raw_f + d_raw
            run = net_f if net_f is not None else net_c
            if ctx.acts_f is not None:
                grads_f = mlp.backward_from_saved(run, R * (S + N_imp), ctx.acts_f, g_raw_f)
                ctx.acts_f = None
            else:
                grads_f = mlp.backward(run, R, S + N_imp, rays, z_f, None, None, g_raw_f.view(R * (S + N_imp), -1))
            g_raw_c = ops.raw2outputs_bwd(raw_c, z_c, rays_d, noise_c, white, d_rgb0, d_disp0, d_acc0, None, d_depth0)
            grads_c = _backward_coarse(ctx, net_c, R, S, rays, z_c, g_raw_c)
            if net_f is None:
                for k in grads_c:
                    grads_c[k] += grads_f[k]
                grads_f = None
        else:
            g_raw_c = ops.raw2outputs_bwd(raw_c, z_c, rays_d, noise_c, white, d_rgb, d_disp, d_acc, None, d_depth)
            if d_raw is not None:
                g_raw_c = g_raw_c + d_raw
            grads_c = _backward_coarse(ctx, net_c, R, S, rays, z_c, g_raw_c)
        out = [grads_c[n] for n in named_c]
        if net_f is not None and N_imp > 0:
            out += [grads_f[n] for n, _ in net_f.named_parameters()]
        elif net_f is not None:
            out += [None for _ in net_f.parameters()]
        return (None, None, None, None, None, None, *out)

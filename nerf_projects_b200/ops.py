"""Tensor-level wrappers over the C ABI (no autograd here; see autograd.py).

Every function takes/returns CUDA fp32 torch tensors and enqueues work on the
current torch CUDA stream.  torch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L

# launch counter: every C-ABI call that enqueues kernels bumps it by the number
# of kernels it launches (bench.py reports it as `gpu_launches`).
_launches = 0


def launches() -> int:
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _empty(shape, like=None, device=None, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=like.device if like is not None else device)


# ------------------------------------------------------------------ rays

def get_rays(H: int, W: int, fx, fy, cx, cy, c2w: torch.Tensor):
    c2w = L.f32c(c2w[:3, :4])
    L.require_cuda(c2w)
    ro = _empty((H, W, 3), c2w)
    rd = _empty((H, W, 3), c2w)
    L.check(L.load().nerfb_get_rays(H, W, float(fx), float(fy), float(cx), float(cy), L.ptr(c2w), L.ptr(ro),
                                    L.ptr(rd), L.stream()), "nerfb_get_rays")
    _count()
    return ro, rd


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    sh = rays_o.shape
    o = L.f32c(rays_o).reshape(-1, 3)
    d = L.f32c(rays_d.expand(sh) if rays_d.shape != sh else rays_d).reshape(-1, 3)
    L.require_cuda(o, d)
    oo, od = torch.empty_like(o), torch.empty_like(d)
    L.check(L.load().nerfb_ndc_rays(o.shape[0], int(H), int(W), float(focal), float(near), L.ptr(o), L.ptr(d),
                                    L.ptr(oo), L.ptr(od), L.stream()), "nerfb_ndc_rays")
    _count()
    return oo.reshape(sh), od.reshape(sh)


def pack_rays(rays_o, rays_d, view_d, use_viewdirs, ndc, H, W, focal, near, far):
    o = L.f32c(rays_o).reshape(-1, 3)
    d = L.f32c(rays_d).reshape(-1, 3)
    v = L.f32c(view_d).reshape(-1, 3) if use_viewdirs else None
    L.require_cuda(o, d, v)
    n = o.shape[0]
    out = _empty((n, 11 if use_viewdirs else 8), o)
    L.check(L.load().nerfb_pack_rays(n, L.ptr(o), L.ptr(d), L.ptr(v), int(bool(use_viewdirs)), int(bool(ndc)),
                                     int(H), int(W), float(focal), float(near), float(far), L.ptr(out),
                                     L.stream()), "nerfb_pack_rays")
    _count()
    return out


def gather_ray_batch(H, W, fx, fy, cx, cy, poses, images, sel, img0=-1):
    """rays_o, rays_d [n,3] (+ target [n,3] when images is given) of the selected pixels; see nerf_b200.h."""
    L.require_cuda(poses, images, sel)
    assert poses.dtype == torch.float32 and poses.is_contiguous() and sel.dtype == torch.int64 and sel.is_contiguous()
    n = sel.numel()
    stride = poses.shape[-2] * poses.shape[-1]
    ro, rd = _empty((n, 3), poses), _empty((n, 3), poses)
    tgt = _empty((n, 3), poses) if images is not None else None
    if images is not None:
        assert images.dtype == torch.float32 and images.is_contiguous() and images.shape[-1] == 3
    L.check(L.load().nerfb_gather_ray_batch(n, int(H), int(W), float(fx), float(fy), float(cx), float(cy), L.ptr(poses),
                                            int(stride), int(poses.shape[0]), L.ptr(images), L.ptr(sel), int(img0),
                                            L.ptr(ro), L.ptr(rd), L.ptr(tgt), L.stream()), "nerfb_gather_ray_batch")
    _count()
    return ro, rd, tgt


def coarse_z(rays, S, lindisp=False, t_rand=None):
    rays = L.f32c(rays)
    L.require_cuda(rays, t_rand)
    R = rays.shape[0]
    z = _empty((R, S), rays)
    if t_rand is not None:
        t_rand = L.f32c(t_rand)
    L.check(L.load().nerfb_coarse_z(R, int(S), L.ptr(rays), rays.shape[1], int(bool(lindisp)), L.ptr(t_rand),
                                    L.ptr(z), L.stream()), "nerfb_coarse_z")
    _count()
    return z


# ------------------------------------------------------------- encoding

def enc_dim(L_: int) -> int:
    return 3 if L_ < 0 else 3 + 6 * L_


def embed(x, Lf: int, out=None, col0=0):
    x2 = L.f32c(x).reshape(-1, x.shape[-1])
    L.require_cuda(x2)
    n = x2.shape[0]
    if out is None:
        out = _empty((n, enc_dim(Lf)), x2)
    L.check(L.load().nerfb_embed(n, L.ptr(x2), x2.shape[1], int(Lf), L.ptr(out), out.shape[1], int(col0),
                                 L.stream()), "nerfb_embed")
    _count()
    return out


def embed_samples(R, S, rays, z, pts, viewdirs, Lp, Lv, enc, col_pts, enc_dir, col_dir):
    L.check(L.load().nerfb_embed_samples(int(R), int(S), L.ptr(rays), rays.shape[1] if rays is not None else 0,
                                         L.ptr(z), L.ptr(pts), L.ptr(viewdirs), int(Lp),
                                         int(Lv) if Lv is not None else L.NERFB_NO_VIEWS, L.ptr(enc), enc.shape[1],
                                         int(col_pts), L.ptr(enc_dir), enc_dir.shape[1] if enc_dir is not None else 0,
                                         int(col_dir), L.stream()), "nerfb_embed_samples")
    _count()


# -------------------------------------------------- fp32 linear layers

def _ld(t):
    assert t.stride(-1) == 1 or t.shape[-1] == 1, "innermost dimension must be contiguous"
    if t.dim() != 2 or t.shape[0] == 1:      # a single row: its stride is arbitrary (numpy transposes give 1), any ld >= width works
        return t.shape[-1]
    return t.stride(0)


def linear_fwd(X, W, b, Y, relu: bool, K: Optional[int] = None):
    """Y[:, :N] = act(X[:, :K] W^T + b); X/Y may be column windows of wider buffers."""
    M = X.shape[0]
    N, Kw = W.shape
    K = Kw if K is None else K
    L.check(L.load().nerfb_linear_fwd(M, N, K, L.ptr(X), _ld(X), L.ptr(W), _ld(W), L.ptr(b), L.ptr(Y), _ld(Y),
                                      int(relu), L.stream()), "nerfb_linear_fwd")
    _count()
    return Y


def linear_bwd_data(dY, W, dX, relu_src=None, accumulate=False):
    """dX = (dX +) dY W  (* relu_src>0).  W may be a column window W[:, a:b]."""
    M, N = dY.shape[0], W.shape[0]
    K = W.shape[1]
    L.check(L.load().nerfb_linear_bwd_data(M, N, K, L.ptr(dY), _ld(dY), L.ptr(W), _ld(W), L.ptr(dX), _ld(dX),
                                           L.ptr(relu_src), _ld(relu_src) if relu_src is not None else 0,
                                           int(accumulate), L.stream()), "nerfb_linear_bwd_data")
    _count()
    return dX


def linear_bwd_weight(dY, X, dW, db):
    """dW += dY^T X ; db += colsum(dY).  dY may be a column window (N = dW.shape[0])."""
    M = X.shape[0]
    N, K = dW.shape
    L.check(L.load().nerfb_linear_bwd_weight(M, N, K, L.ptr(dY), _ld(dY), L.ptr(X), _ld(X), L.ptr(dW), _ld(dW),
                                             L.ptr(db), L.stream()), "nerfb_linear_bwd_weight")
    _count(2 if db is not None else 1)


# ---------------------------------------------------------- compositing

def raw2outputs_fwd(raw, z, rays_d, noise=None, white_bkgd=False, want_weights=True):
    R, S, ch = raw.shape
    L.require_cuda(raw, z, rays_d, noise)
    rgb = _empty((R, 3), raw)
    disp = _empty((R,), raw)
    acc = _empty((R,), raw)
    depth = _empty((R,), raw)
    weights = _empty((R, S), raw) if want_weights else None
    L.check(L.load().nerfb_raw2outputs_fwd(R, S, L.ptr(raw), ch, L.ptr(z), L.ptr(rays_d), rays_d.stride(0),
                                           L.ptr(noise), int(bool(white_bkgd)), L.ptr(rgb), L.ptr(disp), L.ptr(acc),
                                           L.ptr(weights), L.ptr(depth), L.stream()), "nerfb_raw2outputs_fwd")
    _count()
    return rgb, disp, acc, weights, depth


def raw2outputs_bwd(raw, z, rays_d, noise, white_bkgd, d_rgb, d_disp, d_acc, d_weights, d_depth):
    R, S, ch = raw.shape
    d_raw = torch.empty_like(raw)
    gs = [None if g is None else L.f32c(g) for g in (d_rgb, d_disp, d_acc, d_weights, d_depth)]
    L.check(L.load().nerfb_raw2outputs_bwd(R, S, L.ptr(raw), ch, L.ptr(z), L.ptr(rays_d), rays_d.stride(0),
                                           L.ptr(noise), int(bool(white_bkgd)), *[L.ptr(g) for g in gs],
                                           L.ptr(d_raw), L.stream()), "nerfb_raw2outputs_bwd")
    _count()
    return d_raw


# ------------------------------------------------ hierarchical sampling

def sample_pdf(bins, weights, N, det=True, u=None, want_inds=False):
    L.require_cuda(bins, weights, u)
    assert bins.dim() == 2 and weights.dim() == 2 and bins.stride(1) == 1 and weights.stride(1) == 1
    R, M = bins.shape
    samples = _empty((R, N), bins)
    inds = _empty((R, N), bins, dtype=torch.int64) if want_inds else None
    if u is not None:
        u = L.f32c(u)
    L.check(L.load().nerfb_sample_pdf(R, M, int(N), L.ptr(bins), bins.stride(0), L.ptr(weights), weights.stride(0),
                                      int(bool(det)), L.ptr(u), L.ptr(samples), L.ptr(inds), L.stream()),
            "nerfb_sample_pdf")
    _count()
    return (samples, inds) if want_inds else samples


def hierarchical_z(z, weights, N, det=True, u=None, want_inds=False):
    L.require_cuda(z, weights, u)
    R, S = z.shape
    z_out = _empty((R, S + N), z)
    z_samples = _empty((R, N), z)
    z_std = _empty((R,), z)
    inds = _empty((R, N), z, dtype=torch.int64) if want_inds else None
    if u is not None:
        u = L.f32c(u)
    L.check(L.load().nerfb_hierarchical_z(R, S, int(N), L.ptr(z), L.ptr(weights), int(bool(det)), L.ptr(u),
                                          L.ptr(z_out), L.ptr(z_samples), L.ptr(inds), L.ptr(z_std), L.stream()),
            "nerfb_hierarchical_z")
    _count()
    return z_out, z_samples, z_std, inds


def composite_resample(raw, z, rays_d, noise, white_bkgd, N, det=True, u=None, want_weights=False):
    """Coarse raw2outputs + hierarchical_z in one launch (coarse weights stay on-chip) ->
    (rgb0, disp0, acc0, depth0, weights|None, z_out, z_samples, z_std)."""
    R, S, ch = raw.shape
    L.require_cuda(raw, z, rays_d, noise, u)
    rgb, disp, acc, depth = _empty((R, 3), raw), _empty((R,), raw), _empty((R,), raw), _empty((R,), raw)
    weights = _empty((R, S), raw) if want_weights else None
    z_out, z_samples, z_std = _empty((R, S + N), raw), _empty((R, N), raw), _empty((R,), raw)
    if u is not None:
        u = L.f32c(u)
    L.check(L.load().nerfb_composite_resample(R, S, int(N), L.ptr(raw), ch, L.ptr(z), L.ptr(rays_d), rays_d.stride(0),
                                              L.ptr(noise), int(bool(white_bkgd)), int(bool(det)), L.ptr(u), L.ptr(rgb),
                                              L.ptr(disp), L.ptr(acc), L.ptr(depth), L.ptr(weights), L.ptr(z_out),
                                              L.ptr(z_samples), L.ptr(z_std), L.stream()), "nerfb_composite_resample")
    _count()
    return rgb, disp, acc, depth, weights, z_out, z_samples, z_std


# ------------------------------------------------ tensor-core MLP (bf16)

def packed_weights_bytes(use_viewdirs: bool) -> int:
    return int(L.load().nerfb_packed_weights_bytes(int(bool(use_viewdirs))))


def pack_weights(param_ptrs: Sequence[Optional[int]], use_viewdirs: bool, out_ch: int, packed: torch.Tensor):
    arr = (C.c_void_p * L.NERFB_NUM_PARAMS)(*[p if p else None for p in param_ptrs])
    L.check(L.load().nerfb_pack_weights(arr, int(bool(use_viewdirs)), int(out_ch), L.ptr(packed), L.stream()),
            "nerfb_pack_weights")
    _count()


# optional per-launch CUDA-event timing of the fused MLP kernel (bench.py's roofline leg)
_mlp_events = None


def profile_mlp(enable: bool):
    """Start (True) / stop (False) recording CUDA events around every fused-MLP launch.
    Stopping returns [(milliseconds, rows, use_viewdirs), ...] (synchronises)."""
    global _mlp_events
    if enable:
        _mlp_events = []
        return None
    evs, _mlp_events = _mlp_events or [], None
    torch.cuda.synchronize()
    return [(a.elapsed_time(b), rows, vd) for a, b, rows, vd in evs]


def mlp_fwd_bf16(R, S, rays, z, pts, viewdirs, use_viewdirs, out_ch, packed, raw):
    if _mlp_events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _mlp_fwd_bf16(R, S, rays, z, pts, viewdirs, use_viewdirs, out_ch, packed, raw)
        e1.record()
        _mlp_events.append((e0, e1, int(R) * int(S), bool(use_viewdirs)))
        return raw
    return _mlp_fwd_bf16(R, S, rays, z, pts, viewdirs, use_viewdirs, out_ch, packed, raw)


def _mlp_fwd_bf16(R, S, rays, z, pts, viewdirs, use_viewdirs, out_ch, packed, raw):
    L.check(L.load().nerfb_mlp_fwd_bf16(int(R), int(S), L.ptr(rays), rays.shape[1] if rays is not None else 0,
                                        L.ptr(z), L.ptr(pts), L.ptr(viewdirs), int(bool(use_viewdirs)), int(out_ch),
                                        L.ptr(packed), L.ptr(raw), L.stream()), "nerfb_mlp_fwd_bf16")
    _count()
    return raw


# ---------------------------- inference: network pass + compositing in one launch (raw never materialised)

def mlp_fwd_composite_bf16(R, S, rays, z, lindisp, use_viewdirs, out_ch, packed, want_wpart: bool):
    """nerf_mlp_kernel<FWD_COMP> -> (rec [R,S/32,8], wpart [R,S] | None): the network pass with the compositing of every
    32-sample segment done in-kernel; finish with composite_finish (final pass) or resample_records (coarse pass)."""
    L.require_cuda(rays, z, packed)
    dev = rays.device
    rec = torch.empty((R, S // 32, 8), dtype=torch.float32, device=dev)
    wpart = torch.empty((R, S), dtype=torch.float32, device=dev) if want_wpart else None
    call = lambda: L.check(L.load().nerfb_mlp_fwd_composite_bf16(
        int(R), int(S), L.ptr(rays), rays.shape[1], L.ptr(z), int(bool(lindisp)), int(bool(use_viewdirs)), int(out_ch),
        L.ptr(packed), L.ptr(rec), L.ptr(wpart), L.stream()), "nerfb_mlp_fwd_composite_bf16")
    if _mlp_events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record()
        _mlp_events.append((e0, e1, int(R) * int(S), bool(use_viewdirs)))
    else:
        call()
    _count()
    return rec, wpart


def composite_finish(rec, white_bkgd):
    """(rgb [R,3], disp, acc, depth) from the segment records of mlp_fwd_composite_bf16."""
    R, nseg = rec.shape[0], rec.shape[1]
    rgb, disp, acc, depth = _empty((R, 3), rec), _empty((R,), rec), _empty((R,), rec), _empty((R,), rec)
    L.check(L.load().nerfb_composite_finish(int(R), int(nseg), L.ptr(rec), int(bool(white_bkgd)), L.ptr(rgb), L.ptr(disp),
                                            L.ptr(acc), L.ptr(depth), L.stream()), "nerfb_composite_finish")
    _count()
    return rgb, disp, acc, depth


def resample_records(R, S, N, rays, z, lindisp, wpart, rec, white_bkgd, det=True, u=None, want_samples=False):
    """Finish the fused coarse pass: (rgb0, disp0, acc0, depth0, z_out [R,S+N], z_samples|None, z_std)."""
    L.require_cuda(rays, z, wpart, rec, u)
    rgb, disp, acc, depth = _empty((R, 3), rays), _empty((R,), rays), _empty((R,), rays), _empty((R,), rays)
    z_out, z_std = _empty((R, S + N), rays), _empty((R,), rays)
    z_samples = _empty((R, N), rays) if want_samples else None
    if u is not None:
        u = L.f32c(u)
    L.check(L.load().nerfb_resample_records(int(R), int(S), int(N), L.ptr(rays), rays.shape[1], L.ptr(z),
                                            int(bool(lindisp)), L.ptr(wpart), L.ptr(rec), int(bool(white_bkgd)),
                                            int(bool(det)), L.ptr(u), L.ptr(rgb), L.ptr(disp), L.ptr(acc), L.ptr(depth),
                                            L.ptr(z_out), L.ptr(z_samples), L.ptr(z_std), L.stream()),
            "nerfb_resample_records")
    _count()
    return rgb, disp, acc, depth, z_out, z_samples, z_std


# ------------------------------------------ tensor-core MLP backward (bf16)

def train_store_bytes(nrows: int, which: int) -> int:
    return int(L.load().nerfb_train_store_bytes(int(nrows), int(which)))


def packed_weights_bwd_bytes() -> int:
    return int(L.load().nerfb_packed_weights_bwd_bytes())


def pack_weights_bwd(param_ptrs: Sequence[Optional[int]], packed_bwd: torch.Tensor):
    arr = (C.c_void_p * L.NERFB_NUM_PARAMS)(*[p if p else None for p in param_ptrs])
    L.check(L.load().nerfb_pack_weights_bwd(arr, L.ptr(packed_bwd), L.stream()), "nerfb_pack_weights_bwd")
    _count()


def mlp_fwd_save_bf16(R, S, rays, z, pts, viewdirs, packed, raw, acts, masks):
    L.check(L.load().nerfb_mlp_fwd_save_bf16(int(R), int(S), L.ptr(rays), rays.shape[1] if rays is not None else 0,
                                             L.ptr(z), L.ptr(pts), L.ptr(viewdirs), L.ptr(packed), L.ptr(raw),
                                             L.ptr(acts), L.ptr(masks), L.stream()), "nerfb_mlp_fwd_save_bf16")
    _count()
    return raw


def mlp_bwd_chain_bf16(nrows, d_raw, packed_bwd, masks, gstore):
    L.check(L.load().nerfb_mlp_bwd_chain_bf16(int(nrows), L.ptr(d_raw), L.ptr(packed_bwd), L.ptr(masks),
                                              L.ptr(gstore), L.stream()), "nerfb_mlp_bwd_chain_bf16")
    _count()


def mlp_wgrad_bf16(nrows, acts, gstore, d_raw_padded, grad_ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * L.NERFB_NUM_PARAMS)(*[p if p else None for p in grad_ptrs])
    L.check(L.load().nerfb_mlp_wgrad_bf16(int(nrows), L.ptr(acts), L.ptr(gstore), L.ptr(d_raw_padded), arr,
                                          L.stream()), "nerfb_mlp_wgrad_bf16")
    _count()


_bwd_flags = {}


def mlp_bwd_fused_bf16(nrows, d_raw_padded, packed_bwd, masks, acts, gstore, grad_ptrs: Sequence[Optional[int]], chain_sms=76):
    """Data-gradient chain + weight gradients of one network pass in ONE cooperative launch (nerf_bwd_fused_kernel)."""
    dev = acts.device
    need = 4 * ((int(nrows) + 511) // 512) * 20
    ent = _bwd_flags.get(dev)
    if ent is None or ent[0].numel() < need:
        ent = [torch.zeros(max(need, 1 << 16), dtype=torch.int32, device=dev), 0]
        _bwd_flags[dev] = ent
    ent[1] += 1
    arr = (C.c_void_p * L.NERFB_NUM_PARAMS)(*[p if p else None for p in grad_ptrs])
    L.check(L.load().nerfb_mlp_bwd_fused_bf16(int(nrows), L.ptr(d_raw_padded), L.ptr(packed_bwd), L.ptr(masks), L.ptr(acts),
                                              L.ptr(gstore), arr, L.ptr(ent[0]), int(ent[1]), int(chain_sms), L.stream()),
            "nerfb_mlp_bwd_fused_bf16")
    _count()


# no view directions (output_linear head): d_raw is [nrows,4] = channels 0..3, d_raw4 [nrows] = channel 4 or None

def pack_weights_bwd_nv(param_ptrs: Sequence[Optional[int]], out_ch: int, packed_bwd: torch.Tensor):
    arr = (C.c_void_p * L.NERFB_NUM_PARAMS)(*[p if p else None for p in param_ptrs])
    L.check(L.load().nerfb_pack_weights_bwd_nv(arr, int(out_ch), L.ptr(packed_bwd), L.stream()),
            "nerfb_pack_weights_bwd_nv")
    _count()


def mlp_fwd_save_nv_bf16(R, S, rays, z, pts, out_ch, packed, raw, acts, masks):
    L.check(L.load().nerfb_mlp_fwd_save_nv_bf16(int(R), int(S), L.ptr(rays), rays.shape[1] if rays is not None else 0,
                                                L.ptr(z), L.ptr(pts), int(out_ch), L.ptr(packed), L.ptr(raw),
                                                L.ptr(acts), L.ptr(masks), L.stream()), "nerfb_mlp_fwd_save_nv_bf16")
    _count()
    return raw


def mlp_bwd_chain_nv_bf16(nrows, d_raw, d_raw4, packed_bwd, masks, gstore):
    L.check(L.load().nerfb_mlp_bwd_chain_nv_bf16(int(nrows), L.ptr(d_raw), L.ptr(d_raw4), L.ptr(packed_bwd),
                                                 L.ptr(masks), L.ptr(gstore), L.stream()), "nerfb_mlp_bwd_chain_nv_bf16")
    _count()


def mlp_wgrad_nv_bf16(nrows, out_ch, acts, gstore, d_raw_padded, d_raw4_padded, grad_ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * L.NERFB_NUM_PARAMS)(*[p if p else None for p in grad_ptrs])
    L.check(L.load().nerfb_mlp_wgrad_nv_bf16(int(nrows), int(out_ch), L.ptr(acts), L.ptr(gstore), L.ptr(d_raw_padded),
                                             L.ptr(d_raw4_padded), arr, L.stream()), "nerfb_mlp_wgrad_nv_bf16")
    _count()


def umma_selftest(A, B, cta_group=1):
    N, K = B.shape
    assert A.shape == (128 * cta_group, K)
    Cm = _empty((128 * cta_group, N), A)
    L.check(L.load().nerfb_umma_selftest(int(cta_group), N, K, L.ptr(A), L.ptr(B), L.ptr(Cm), L.stream()),
            "nerfb_umma_selftest")
    _count()
    return Cm


def umma_selftest_bias(A, Bw, Bb, A8=None):
    N = Bw.shape[0]
    assert A.shape == (256, 64) and Bw.shape == (N, 64) and Bb.shape == (N, 16) and (A8 is None or A8.shape == (8, 16))
    Cm = _empty((256, N), A)
    L.check(L.load().nerfb_umma_selftest_bias(N, L.ptr(A), L.ptr(Bw), L.ptr(Bb), L.ptr(A8), L.ptr(Cm), L.stream()),
            "nerfb_umma_selftest_bias")
    _count()
    return Cm


def umma_selftest_mn(At, Bt):
    K, N = Bt.shape
    assert At.shape == (K, 128)
    Cm = _empty((128, N), At)
    L.check(L.load().nerfb_umma_selftest_mn(N, K, L.ptr(At), L.ptr(Bt), L.ptr(Cm), L.stream()),
            "nerfb_umma_selftest_mn")
    _count()
    return Cm


# ------------------------------------------------------------ optimiser

def adam_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0):
    n = len(params)
    for i0 in range(0, n, L.NERFB_ADAM_MAX):
        sl = slice(i0, min(n, i0 + L.NERFB_ADAM_MAX))
        k = sl.stop - sl.start
        mk = lambda ts: (C.c_void_p * k)(*[t.data_ptr() for t in ts[sl]])
        numel = (C.c_longlong * k)(*[t.numel() for t in params[sl]])
        L.check(L.load().nerfb_adam_step(k, mk(params), mk(grads), mk(exp_avg), mk(exp_avg_sq), numel, float(lr),
                                         float(beta1), float(beta2), float(eps), int(step), float(grad_scale),
                                         L.stream()), "nerfb_adam_step")
        _count()

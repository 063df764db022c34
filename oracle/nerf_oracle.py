"""CPU oracle for the original-NeRF render/train hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline.  The product path
(``nerf_projects_b200``) never imports this module and fails loudly when its
CUDA library is missing.

What it is: a from-scratch restatement, on CPU in fp32, of the algorithm the
reference implements with eager PyTorch ops.  The arithmetic library is the
same one the reference uses (ATen CPU kernels, torch 2.11), because the
reference's results are *defined* by those kernels (summation order of
``torch.sum``, fp64 running accumulator of ``torch.cumsum``/``cumprod``,
``searchsorted(right=True)``) -- see SURVEY.md section 7.3-A.  A second,
numpy-only restatement of ``sample_pdf`` that spells that arithmetic out bit
for bit lives in ``oracle/sample_pdf_np.py``.

Parity pinned: yes.  ``tools/make_golden.py`` executes the *unmodified*
reference (``/root/reference/nerf/{nerf,embedder,nerf_helpers}.py`` plus the
notebook cells of ``nerf.ipynb``) in the build container and commits its
outputs under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks
every function here against those vectors (bit-exact for sample_pdf
indices/samples, <=1e-6 for float maps).

Each function cites the reference lines it follows (``ipynb:N`` = raw JSON line
number of /root/reference/nerf/nerf.ipynb).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = torch.float32

# --------------------------------------------------------------------------
# deterministic, torch-version-independent parameter initialisation
# --------------------------------------------------------------------------

def layer_shapes(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4,
                 skips=(4,), use_viewdirs=True):
    """(name, out_features, in_features) in ``state_dict`` order of the
    reference ``NeRF`` module (nerf.py:32-55)."""
    out = []
    for i in range(D):
        if i == 0:
            fan_in = input_ch
        elif (i - 1) in skips:
            fan_in = W + input_ch
        else:
            fan_in = W
        out.append((f"pts_linears.{i}", W, fan_in))
    out.append(("views_linears.0", W // 2, input_ch_views + W))
    if use_viewdirs:
        out.append(("feature_linear", W, W))
        out.append(("alpha_linear", 1, W))
        out.append(("rgb_linear", 3, W // 2))
    else:
        out.append(("output_linear", output_ch, W))
    return out


def make_params(seed: int, **arch) -> Dict[str, np.ndarray]:
    """nn.Linear-style init (U(+-1/sqrt(fan_in)) for weight and bias, which is
    what kaiming_uniform(a=sqrt(5)) reduces to) drawn from numpy PCG64 so the
    values do not depend on the torch build.  Keys/shapes == reference
    ``NeRF.state_dict()`` (nerf.py:32-55)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    params = {}
    for name, fo, fi in layer_shapes(**arch):
        bound = 1.0 / math.sqrt(fi)
        params[name + ".weight"] = rng.uniform(-bound, bound, size=(fo, fi)).astype(np.float32)
        params[name + ".bias"] = rng.uniform(-bound, bound, size=(fo,)).astype(np.float32)
    return params


def smooth_spectrum(params: Dict[str, np.ndarray], falloff=0.5, input_ch=63, skips=(4,)):
    """Damps the weights that read frequency band k of gamma(x) by falloff**k (first layer and
    the gamma(x) columns of the skip layers).  A random-init network is white noise at the
    2^9 band -- sigma changes by O(1) between adjacent samples, unlike any trained field -- which
    turns the (legitimately discontinuous) resampling step into a chaos amplifier.  The damped
    network is a band-limited synthetic scene.  Deterministic; only those columns change."""
    p = dict(params)
    L = (input_ch - 3) // 6
    scale = np.ones(input_ch, np.float32)
    for k in range(L):
        scale[3 + 6 * k: 3 + 6 * k + 6] = np.float32(falloff ** k)
    names = ["pts_linears.0.weight"] + [f"pts_linears.{i + 1}.weight" for i in skips]
    for n in names:
        w = p[n].copy()
        w[:, :input_ch] *= scale
        p[n] = w
    return p


def density_head(params: Dict[str, np.ndarray]):
    """(key, row) of the parameters that produce sigma."""
    return ("alpha_linear", 0) if "alpha_linear.weight" in params else ("output_linear", 3)


def calibrate_density(params: Dict[str, np.ndarray], pts: np.ndarray, viewdirs: Optional[np.ndarray],
                      target_std=3.0, empty_fraction=0.6, **net):
    """Turns a random-init network into a non-trivial synthetic "scene"
    (SURVEY.md 7.3-B iii): rescales / shifts the density head so that over the
    given sample positions sigma has standard deviation ``target_std`` and is
    negative (empty space) on ``empty_fraction`` of them.  Plain random init has
    sigma of one sign almost everywhere, which makes compositing and sample_pdf
    degenerate.  Returns a new params dict (only the density row changes)."""
    p = dict(params)
    key, row = density_head(p)
    with torch.no_grad():
        raw = run_network(torch.from_numpy(pts[:, None, :]),
                          torch.from_numpy(viewdirs) if viewdirs is not None else None, to_torch(p), **net)
    s = raw[:, 0, 3].numpy().astype(np.float64)
    gain = target_std / max(s.std(), 1e-12)
    shift = np.quantile(s, empty_fraction)
    w = p[key + ".weight"].copy()
    b = p[key + ".bias"].copy()
    w[row] = (w[row] * gain).astype(np.float32)
    b[row] = np.float32((b[row] - shift) * gain)
    p[key + ".weight"], p[key + ".bias"] = w, b
    return p


def to_torch(params: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}

# --------------------------------------------------------------------------
# cameras (synthetic inputs of SURVEY.md 8d)
# --------------------------------------------------------------------------

def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """load_blender.py:10-34: translate along z, tilt by phi, spin by theta,
    then swap axes into the Blender world frame.  Returns 4x4 float32."""
    th, ph = theta_deg / 180.0 * np.pi, phi_deg / 180.0 * np.pi
    T = np.eye(4, dtype=np.float32); T[2, 3] = radius
    Rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0],
                   [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], dtype=np.float32)
    Rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0],
                   [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return (flip @ (Rt @ (Rp @ T))).astype(np.float32)


def blender_camera(H: int, W: int, theta_deg: float = 30.0):
    """Blender-shaped camera: camera_angle_x of the lego scene
    (load_blender.py:72-73; value in ipynb:1458), K as ipynb:1088-1094."""
    cax = 0.6911112070083618
    focal = 0.5 * W / np.tan(0.5 * cax)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float64)
    return K, pose_spherical(theta_deg, -30.0, 4.0)[:3, :4]


def llff_camera(H: int = 756, W: int = 1008, focal: float = 815.0):
    """Forward-facing camera looking down -z with a small offset/tilt."""
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float64)
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float32)
    t = np.array([[0.1], [-0.05], [0.2]], dtype=np.float32)
    return K, np.concatenate([R, t], axis=1).astype(np.float32)

# --------------------------------------------------------------------------
# rays
# --------------------------------------------------------------------------

def get_rays(H: int, W: int, K, c2w: torch.Tensor):
    """nerf_helpers.py:272-296.  Integer pixel coordinates (no +0.5),
    dirs = ((i-cx)/fx, -(j-cy)/fy, -1), rotated by c2w[:3,:3]; origin = c2w[:3,3]."""
    c2w = torch.as_tensor(c2w)
    xs = torch.linspace(0, W - 1, W, dtype=c2w.dtype)
    ys = torch.linspace(0, H - 1, H, dtype=c2w.dtype)
    px = xs[None, :].expand(H, W)
    py = ys[:, None].expand(H, W)
    cam = torch.stack([(px - K[0][2]) / K[0][0], -(py - K[1][2]) / K[1][1], -torch.ones_like(px)], -1)
    rays_d = torch.sum(cam[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H: int, W: int, focal: float, near: float, rays_o, rays_d):
    """nerf_helpers.py:342-369, same operation order."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx = -1.0 / (W / (2.0 * focal))
    sy = -1.0 / (H / (2.0 * focal))
    o0 = sx * rays_o[..., 0] / rays_o[..., 2]
    o1 = sy * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = sx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = sy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)

# --------------------------------------------------------------------------
# positional encoding and the MLP
# --------------------------------------------------------------------------

def embed(x: torch.Tensor, L: int) -> torch.Tensor:
    """embedder.py:36-80 with include_input, log_sampling: columns are
    [x | sin(1x) | cos(1x) | sin(2x) | cos(2x) | ... | cos(2^(L-1) x)], three
    channels per block.  L < 0 means identity (get_embedder i=-1)."""
    if L < 0:
        return x
    freqs = (2.0 ** torch.linspace(0.0, L - 1, steps=L)).tolist() if L > 0 else []
    cols = [x]
    for f in freqs:
        cols.append(torch.sin(x * f))
        cols.append(torch.cos(x * f))
    return torch.cat(cols, -1)


def mlp_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, *, D=8, input_ch=63,
                input_ch_views=27, skips=(4,), use_viewdirs=True) -> torch.Tensor:
    """nerf.py:57-111.  Skip concat is [gamma(x), h] after layer index in
    ``skips``; view concat is [feature, gamma(d)]; output is [rgb, sigma]."""
    lin = torch.nn.functional.linear
    pts, views = torch.split(x, [input_ch, input_ch_views], dim=-1)
    h = pts
    for i in range(D):
        h = torch.relu(lin(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([pts, h], -1)
    if use_viewdirs:
        sigma = lin(h, p["alpha_linear.weight"], p["alpha_linear.bias"])
        feat = lin(h, p["feature_linear.weight"], p["feature_linear.bias"])
        hv = torch.relu(lin(torch.cat([feat, views], -1), p["views_linears.0.weight"], p["views_linears.0.bias"]))
        rgb = lin(hv, p["rgb_linear.weight"], p["rgb_linear.bias"])
        return torch.cat([rgb, sigma], -1)
    return lin(h, p["output_linear.weight"], p["output_linear.bias"])


def enc_dim(L: int) -> int:
    return 3 if L < 0 else 3 + 6 * L


def run_network(pts: torch.Tensor, viewdirs: Optional[torch.Tensor], p, *, multires=10,
                multires_views=4, netchunk=65536, D=8, skips=(4,)) -> torch.Tensor:
    """ipynb:790-855 (+batchify ipynb:224-244): flatten, encode, broadcast the
    per-ray view direction to every sample, concat, chunked MLP, reshape."""
    flat = pts.reshape(-1, pts.shape[-1])
    enc = embed(flat, multires)
    use_vd = viewdirs is not None
    if use_vd:
        dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
        enc = torch.cat([enc, embed(dirs, multires_views)], -1)
    kw = dict(D=D, skips=skips, use_viewdirs=use_vd, input_ch=enc_dim(multires),
              input_ch_views=enc_dim(multires_views) if use_vd else 0)
    outs = [mlp_forward(p, enc[i:i + netchunk], **kw) for i in range(0, enc.shape[0], netchunk)]
    out = torch.cat(outs, 0)
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])

# --------------------------------------------------------------------------
# compositing and hierarchical sampling
# --------------------------------------------------------------------------

def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0.0, white_bkgd=False, noise=None):
    """ipynb:254-349.  ``noise`` (optional tensor like raw[...,3]) replaces the
    reference's torch.randn draw so tests can inject a fixed perturbation."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    far_gap = torch.full_like(dists[..., :1], 1e10)
    dists = torch.cat([dists, far_gap], -1) * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3]
    try:
        std = float(raw_noise_std)
    except (TypeError, ValueError):
        std = 0.0
    if std > 0.0:
        sigma = sigma + (noise if noise is not None else torch.randn_like(sigma) * std)
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * dists)
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    denom = torch.max(1e-10 * torch.ones_like(depth_map), acc_map)
    disp_map = 1.0 / torch.clamp(depth_map / denom, min=1e-10)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


def sample_pdf(bins, weights, N_samples, det=True, u=None, return_inds=False):
    """nerf_helpers.py:396-439.  ``u`` overrides the random draw (det=False)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, steps=N_samples, dtype=cdf.dtype)
            u = u.expand(list(cdf.shape[:-1]) + [N_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [N_samples], dtype=cdf.dtype)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    denom = c_hi - c_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - c_lo) / denom
    samples = b_lo + t * (b_hi - b_lo)
    return (samples, inds) if return_inds else samples


def coarse_z(near, far, N_samples, lindisp=False):
    """ipynb:418-425: z = near*(1-t) + far*t (or the disparity-linear form)."""
    t = torch.linspace(0.0, 1.0, steps=N_samples)
    if not lindisp:
        return near * (1.0 - t) + far * t
    return 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)


def render_rays(ray_batch, p_coarse, p_fine=None, *, N_samples=64, N_importance=0,
                lindisp=False, white_bkgd=False, raw_noise_std=0.0, retraw=False,
                t_rand=None, u=None, **net):
    """ipynb:359-492 with perturb handled through explicit ``t_rand`` (the
    stratified jitter, [R,N_samples] in [0,1)) and ``u`` (sample_pdf draws);
    both None reproduces perturb=0.  Also returns depth/z_vals for checks."""
    R = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z = coarse_z(near, far, N_samples, lindisp).expand(R, N_samples)
    if t_rand is not None:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mid, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mid], -1)
        z = lower + (upper - lower) * t_rand
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]
    raw = run_network(pts, viewdirs, p_coarse, **net)
    rgb, disp, acc, weights, depth = raw2outputs(raw, z, rays_d, raw_noise_std, white_bkgd)
    out = {}
    if N_importance > 0:
        out.update(rgb0=rgb, disp0=disp, acc0=acc, depth0=depth, weights0=weights, raw0=raw, z_coarse=z)
        z_mid = 0.5 * (z[..., 1:] + z[..., :-1])
        z_new, inds = sample_pdf(z_mid, weights[..., 1:-1], N_importance, det=(u is None and t_rand is None),
                                 u=u, return_inds=True)
        z_new = z_new.detach()
        z, _ = torch.sort(torch.cat([z, z_new], -1), -1)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]
        raw = run_network(pts, viewdirs, p_fine if p_fine is not None else p_coarse, **net)
        rgb, disp, acc, weights, depth = raw2outputs(raw, z, rays_d, raw_noise_std, white_bkgd)
        out.update(z_samples=z_new, inds=inds, z_std=torch.std(z_new, dim=-1, unbiased=False))
    out.update(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, weights=weights, z_vals=z)
    if retraw:
        out["raw"] = raw
    return out


def pack_rays(rays_o, rays_d, near, far, *, H=None, W=None, K=None, ndc=False, use_viewdirs=False):
    """ipynb:605-629: viewdirs from the *pre-NDC* directions, optional NDC
    warp with near=1, then [o, d, near, far, (viewdir)] as float32 rows."""
    sh = rays_d.shape
    vd = None
    if use_viewdirs:
        vd = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        vd = vd.reshape(-1, 3).float()
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1.0, rays_o, rays_d)
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    cols = [rays_o, rays_d, near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])]
    if vd is not None:
        cols.append(vd)
    return torch.cat(cols, -1), sh


def render(H, W, K, p_coarse, p_fine=None, *, chunk=32768, rays=None, c2w=None, ndc=True,
           near=0.0, far=1.0, use_viewdirs=False, **kw):
    """ipynb:558-640 + batchify_rays ipynb:514-548."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, torch.as_tensor(c2w))
    else:
        rays_o, rays_d = rays
    packed, sh = pack_rays(rays_o, rays_d, near, far, H=H, W=W, K=K, ndc=ndc, use_viewdirs=use_viewdirs)
    parts: Dict[str, list] = {}
    for i in range(0, packed.shape[0], chunk):
        r = render_rays(packed[i:i + chunk], p_coarse, p_fine, **kw)
        for k, v in r.items():
            parts.setdefault(k, []).append(v)
    out = {k: torch.cat(v, 0) for k, v in parts.items()}
    return {k: v.reshape(list(sh[:-1]) + list(v.shape[1:])) for k, v in out.items()}


def img2mse(x, y):
    """nerf_helpers.py:8."""
    return torch.mean((x - y) ** 2)


def train_loss_and_grads(ray_batch, target, p_coarse, p_fine, **kw):
    """One optimisation objective of train() (ipynb:1258-1274): MSE on the
    fine rgb plus MSE on rgb0; returns loss and d loss / d params."""
    pc = {k: v.clone().requires_grad_(True) for k, v in p_coarse.items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in p_fine.items()} if p_fine is not None else None
    out = render_rays(ray_batch, pc, pf, **kw)
    loss = img2mse(out["rgb_map"], target)
    if "rgb0" in out:
        loss = loss + img2mse(out["rgb0"], target)
    loss.backward()
    gc = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pc.items()}
    gf = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pf.items()} if pf else None
    return loss.detach(), gc, gf, out

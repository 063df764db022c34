#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It imports
``nerf.py``, ``embedder.py`` and ``nerf_helpers.py`` from
/root/reference/nerf and ``exec``s the notebook cells of ``nerf.ipynb`` that
define ``batchify``, ``raw2outputs``, ``render_rays``, ``batchify_rays``,
``render``, ``run_network`` and ``create_nerf`` (the notebook itself cannot be
run: no IPython/imageio here -- SURVEY.md 8c).  No reference source is copied
into this repository; only inputs/outputs are stored.

Inputs are generated with numpy PCG64 (``oracle.nerf_oracle.make_params`` and
the seeds below), so the fixtures can be regenerated bit-identically:

    python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/nerf"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import nerf_oracle as O  # noqa: E402  (only for make_params / cameras = input generation)

torch.set_default_dtype(torch.float32)
torch.set_num_threads(8)


def load_reference():
    import nerf as ref_nerf            # /root/reference/nerf/nerf.py
    import embedder as ref_embedder    # /root/reference/nerf/embedder.py
    import nerf_helpers as ref_helpers  # /root/reference/nerf/nerf_helpers.py
    import torch.nn.functional as F
    import time
    ns = dict(torch=torch, np=np, F=F, DEBUG=False, time=time, os=os,
              get_rays=ref_helpers.get_rays, ndc_rays=ref_helpers.ndc_rays,
              sample_pdf=ref_helpers.sample_pdf, get_embedder=ref_embedder.get_embedder,
              device=torch.device("cpu"), tqdm=lambda x: x)
    nb = json.load(open(os.path.join(REF, "nerf.ipynb")))
    wanted = {"batchify", "raw2outputs", "render_rays", "batchify_rays", "render", "run_network", "create_nerf"}
    for cell in nb["cells"]:
        if cell["cell_type"] != "code":
            continue
        src = "".join(cell["source"])
        if any(("def %s(" % w) in src for w in wanted):
            exec(compile(src, "<nerf.ipynb>", "exec"), ns)
    ns.update(ref_nerf=ref_nerf, ref_embedder=ref_embedder, ref_helpers=ref_helpers)
    return types.SimpleNamespace(**ns)


def capture_searchsorted():
    """Wrap torch.searchsorted so the reference sample_pdf's indices (which it
    does not return) can be recorded."""
    rec = {}
    orig = torch.searchsorted

    def wrapped(cdf, u, **kw):
        out = orig(cdf, u, **kw)
        rec["cdf"], rec["u"], rec["inds"] = cdf.clone(), u.clone(), out.clone()
        return out
    return rec, orig, wrapped


def ref_models(R, params_c, params_f, use_viewdirs, N_importance, multires=10, multires_views=4):
    """Reference create_nerf() on a throw-away basedir, then load our params."""
    tmp = tempfile.mkdtemp(prefix="golden_")
    args = types.SimpleNamespace(multires=multires, i_embed=0, use_viewdirs=use_viewdirs,
                                 multires_views=multires_views, N_importance=N_importance,
                                 netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256,
                                 netchunk=65536, lrate=5e-4, basedir=tmp, expname="g", ft_path=None,
                                 no_reload=True, perturb=0.0, N_samples=64, white_bkgd=False,
                                 raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        kw_train, kw_test, start, grad_vars, optim = R.create_nerf(args)
    kw_test["network_fn"].load_state_dict(O.to_torch(params_c))
    if kw_test["network_fine"] is not None:
        kw_test["network_fine"].load_state_dict(O.to_torch(params_f))
    return kw_train, kw_test, grad_vars


def subset_rays(H, W, K, c2w, n, seed):
    rays_o, rays_d = O.get_rays(H, W, K, torch.from_numpy(c2w))
    rng = np.random.Generator(np.random.PCG64(seed))
    idx = np.sort(rng.choice(H * W, size=n, replace=False))
    return rays_o.reshape(-1, 3)[idx].clone(), rays_d.reshape(-1, 3)[idx].clone(), idx


def save(name, **arrs):
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(out, **conv)
    print(f"{name}: {os.path.getsize(out)/1024:.1f} KiB  keys={sorted(conv)}")


def main():
    R = load_reference()
    rng = np.random.Generator(np.random.PCG64(2024))

    # ---- a1/a2: get_rays, ndc_rays -------------------------------------
    Kb, c2wb = O.blender_camera(20, 24, theta_deg=37.0)
    ro, rd = R.get_rays(20, 24, Kb, torch.from_numpy(c2wb))
    Kl, c2wl = O.llff_camera(18, 24, focal=19.5)
    rol, rdl = R.get_rays(18, 24, Kl, torch.from_numpy(c2wl))
    no, nd = R.ndc_rays(18, 24, Kl[0][0], 1.0, rol, rdl)
    save("rays", Kb=Kb, c2wb=c2wb, rays_o_b=ro, rays_d_b=rd, Kl=Kl, c2wl=c2wl,
         rays_o_l=rol, rays_d_l=rdl, ndc_o=no, ndc_d=nd)

    # ---- a7: embedder ---------------------------------------------------
    x = torch.from_numpy(rng.uniform(-6, 6, size=(96, 3)).astype(np.float32))
    e10, d10 = R.get_embedder(10, 0)
    e4, d4 = R.get_embedder(4, 0)
    save("embed", x=x, e10=e10(x), e4=e4(x), d10=d10, d4=d4)

    # ---- a8: NeRF forward ----------------------------------------------
    pv = O.make_params(11, use_viewdirs=True)
    pn = O.make_params(12, use_viewdirs=False, input_ch_views=0, output_ch=5)
    xin = torch.from_numpy(rng.uniform(-1, 1, size=(80, 90)).astype(np.float32))
    mv = R.ref_nerf.NeRF(8, 256, 63, 27, 5, [4], True); mv.load_state_dict(O.to_torch(pv))
    mn = R.ref_nerf.NeRF(8, 256, 63, 0, 5, [4], False); mn.load_state_dict(O.to_torch(pn))
    with torch.no_grad():
        save("mlp", x=xin, seed_v=11, seed_n=12, out_v=mv(xin), out_n=mn(xin[:, :63]))

    # ---- a9: raw2outputs -----------------------------------------------
    raw = torch.from_numpy(rng.normal(0, 1.0, size=(48, 64, 4)).astype(np.float32))
    raw[..., 3] = raw[..., 3] * 6.0 - 1.0
    raw[:8, :, 3] = -1.0            # empty rays
    raw[8:12, :, 3] = 50.0          # opaque at first sample
    z = torch.from_numpy(np.sort(rng.uniform(2, 6, size=(48, 64)).astype(np.float32), -1))
    d = torch.from_numpy(rng.normal(0, 1, size=(48, 3)).astype(np.float32))
    outs = {}
    for wb in (False, True):
        r = R.raw2outputs(raw, z, d, 0, wb)
        for n, v in zip(("rgb", "disp", "acc", "weights", "depth"), r):
            outs[f"{n}_wb{int(wb)}"] = v
    save("raw2outputs", raw=raw, z=z, rays_d=d, **outs)

    # ---- a10: sample_pdf ------------------------------------------------
    Rn = 384
    z_mid = torch.from_numpy(np.sort(rng.uniform(2, 6, size=(Rn, 63)).astype(np.float32), -1))
    w = rng.uniform(0, 1, size=(Rn, 62)).astype(np.float32)
    w[:64] = w[:64] ** 8                         # peaky
    w[64:96] = 0.0                               # all-zero -> uniform pdf from +1e-5
    w[96:128, :] = 0.0; w[96:128, 30] = 1.0      # single spike
    w[128:160] *= 1e-6                           # tiny
    w[160:192, 10:50] = 0.0                      # wide empty gap (denom < 1e-5 path)
    w = torch.from_numpy(w)
    rec, orig, wrapped = capture_searchsorted()
    torch.searchsorted = wrapped
    try:
        s128 = R.sample_pdf(z_mid, w, 128, det=True)
        i128 = rec["inds"].clone(); cdf128 = rec["cdf"].clone()
        s64 = R.sample_pdf(z_mid, w, 64, det=True)
        i64 = rec["inds"].clone()
        torch.manual_seed(7)
        srand = R.sample_pdf(z_mid, w, 128, det=False)
        irand = rec["inds"].clone(); urand = rec["u"].clone()
    finally:
        torch.searchsorted = orig
    save("sample_pdf", bins=z_mid, weights=w, s128=s128, i128=i128.to(torch.int16), cdf=cdf128,
         s64=s64, i64=i64.to(torch.int16), urand=urand, srand=srand, irand=irand.to(torch.int16))

    # ---- a3-a6,a11,a12: render() end to end -----------------------------
    def scene(ro, rd, pc, pf, use_viewdirs, ndc, H, W, K, near, far, lindisp, N_samples, target_std=3.0):
        """calibrate both density heads on the coarse sample positions of these rays"""
        packed, _ = O.pack_rays(ro, rd, near, far, H=H, W=W, K=K, ndc=ndc, use_viewdirs=use_viewdirs)
        z = O.coarse_z(packed[:, 6:7], packed[:, 7:8], N_samples, lindisp)
        pts = (packed[:, None, 0:3] + packed[:, None, 3:6] * z[..., None]).reshape(-1, 3).numpy()
        vd = packed[:, None, 8:11].expand(-1, N_samples, -1).reshape(-1, 3).numpy() if use_viewdirs else None
        pc2 = O.calibrate_density(pc, pts, vd, target_std=target_std)
        pf2 = O.calibrate_density(pf, pts, vd, target_std=target_std) if pf is not None else None
        return pc2, pf2

    def head_arrays(prefix, p):
        key, _ = O.density_head(p)
        return {f"{prefix}.{key}.weight": p[key + ".weight"], f"{prefix}.{key}.bias": p[key + ".bias"]}

    def run_render(tag, H, W, K, c2w, n, seed, use_viewdirs, N_imp, white, ndc, near, far, lindisp,
                   pc, pf, N_samples=64, calibrate=False, target_std=3.0):
        ro, rd, idx = subset_rays(H, W, K, c2w, n, seed)
        extra = {}
        if calibrate:
            same = pf is pc
            pc = O.smooth_spectrum(pc)
            pf = pc if same else (O.smooth_spectrum(pf) if pf is not None else None)
            pc, pf2 = scene(ro, rd, pc, None if same else pf, use_viewdirs, ndc, H, W, K, near, far, lindisp, N_samples,
                            target_std)
            pf = pc if same else pf2
            extra.update(head_arrays("override.c", pc))
            if pf is not None and not same:
                extra.update(head_arrays("override.f", pf))
        kw_train, kw_test, _ = ref_models(R, pc, pf, use_viewdirs, N_imp)
        kw = dict(kw_test); kw.update(white_bkgd=white, ndc=ndc, lindisp=lindisp, N_samples=N_samples)
        rec, orig, wrapped = capture_searchsorted()
        torch.searchsorted = wrapped
        try:
            with torch.no_grad():
                rgb, disp, acc, extras = R.render(H, W, K, chunk=32768, rays=(ro, rd), near=near, far=far,
                                                  retraw=True, **kw)
        finally:
            torch.searchsorted = orig
        out = dict(H=H, W=W, K=K, c2w=c2w, idx=idx, rays_o=ro, rays_d=rd, near=near, far=far,
                   rgb_map=rgb, disp_map=disp, acc_map=acc, raw=extras["raw"])
        for k in ("rgb0", "disp0", "acc0", "z_std"):
            if k in extras:
                out[k] = extras[k]
        if "inds" in rec:
            out["inds"] = rec["inds"].to(torch.int16)
        out.update(extra)
        save(tag, **out)
        return pc, pf

    Hc, Wc = 100, 100
    K0, c0 = O.blender_camera(Hc, Wc, 30.0)
    p0 = O.make_params(100, use_viewdirs=False, input_ch_views=0, output_ch=4)
    run_render("render_c0", Hc, Wc, K0, c0, 160, 1, False, 0, False, False, 2.0, 6.0, False, p0, None)

    K1, c1 = O.blender_camera(800, 800, 30.0)
    pc = O.make_params(0, use_viewdirs=True, output_ch=5)
    pf = O.make_params(1, use_viewdirs=True, output_ch=5)
    run_render("render_c1", 800, 800, K1, c1, 64, 2, True, 128, True, False, 2.0, 6.0, False, pc, pf)
    pcs, pfs = run_render("render_c1_sharp", 800, 800, K1, c1, 64, 3, True, 128, True, False, 2.0, 6.0, False,
                          pc, pf, calibrate=True)
    run_render("render_lindisp", 800, 800, K1, c1, 32, 4, True, 128, False, False, 2.0, 6.0, True, pc, pf,
               calibrate=True)
    K3, c3 = O.llff_camera(756, 1008, 815.0)
    run_render("render_llff", 756, 1008, K3, c3, 64, 5, True, 128, False, True, 0.0, 1.0, False, pc, pf,
               calibrate=True)
    # no-viewdirs hierarchical (output_ch=5, fine net None -> coarse net reused) and odd sample counts
    pn5 = O.make_params(7, use_viewdirs=False, input_ch_views=0, output_ch=5)
    run_render("render_nov_40_24", 800, 800, K1, c1, 32, 6, False, 24, True, False, 2.0, 6.0, False, pn5, pn5,
               N_samples=40, calibrate=True, target_std=1.5)   # 64 evals/ray only: gentler density
    p0s = O.make_params(101, use_viewdirs=False, input_ch_views=0, output_ch=4)
    run_render("render_c0_scene", Hc, Wc, K0, c0, 160, 8, False, 0, True, False, 2.0, 6.0, False, p0s, None,
               calibrate=True)

    # ---- a13: loss + gradients of one training objective ----------------
    ro, rd, idx = subset_rays(800, 800, K1, c1, 24, 9)
    target = torch.from_numpy(rng.uniform(0, 1, size=(24, 3)).astype(np.float32))
    kw_train, kw_test, grad_vars = ref_models(R, pcs, pfs, True, 128)
    kw = dict(kw_test); kw.update(white_bkgd=True, ndc=False, lindisp=False)
    rgb, disp, acc, extras = R.render(800, 800, K1, chunk=32768, rays=(ro, rd), near=2.0, far=6.0,
                                      retraw=True, **kw)
    loss = R.ref_helpers.img2mse(rgb, target) + R.ref_helpers.img2mse(extras["rgb0"], target)
    loss.backward()
    g = dict(rays_o=ro, rays_d=rd, target=target, loss=loss.detach())
    g.update(head_arrays("override.c", pcs)); g.update(head_arrays("override.f", pfs))
    srng = np.random.Generator(np.random.PCG64(5))
    for tag, net in (("c", kw["network_fn"]), ("f", kw["network_fine"])):
        for name, prm in net.named_parameters():
            gr = prm.grad if prm.grad is not None else torch.zeros_like(prm)
            flat = gr.reshape(-1)
            g[f"{tag}.{name}.norm"] = flat.double().norm().float()
            g[f"{tag}.{name}.sum"] = flat.double().sum().float()
            if flat.numel() <= 4096:
                g[f"{tag}.{name}.full"] = gr
            else:
                sel = np.sort(srng.choice(flat.numel(), size=2048, replace=False))
                g[f"{tag}.{name}.idx"] = sel.astype(np.int32)
                g[f"{tag}.{name}.val"] = flat[torch.from_numpy(sel)]
    save("train_grads", **g)


if __name__ == "__main__":
    main()

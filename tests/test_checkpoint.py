"""Checkpoint / Keras-weights compatibility pinned by the reference.

``tests/golden/ref_ckpt_w64.tar`` was WRITTEN BY THE REFERENCE (its create_nerf + torch.optim.Adam, two iterations of
its train() loop body, its torch.save dictionary -- tools/make_golden_ckpt.py, ipynb:876-961, 1258-1299); the companion
``ref_ckpt_w64.npz`` holds what the reference itself produced after RESUMING from that file and taking one more step.
``keras_weights.npz`` holds digests of the reference's ``NeRF.load_weights_from_keras`` (nerf.py:113-146).
"""
import os
import shutil
import types

import numpy as np
import pytest
import torch

from _cases import GOLD, T, gold

LRATE, LRATE_DECAY = 5e-4, 250


def _args(basedir, **over):
    a = dict(multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=128, netdepth=8, netwidth=64,
             netdepth_fine=8, netwidth_fine=64, netchunk=65536, lrate=LRATE, basedir=basedir, expname="ck", ft_path=None,
             no_reload=False, perturb=0.0, N_samples=64, white_bkgd=True, raw_noise_std=0.0, dataset_type="blender",
             no_ndc=False, lindisp=False)
    a.update(over)
    return types.SimpleNamespace(**a)


def _place_reference_checkpoint(tmp_path):
    d = tmp_path / "ck" / "checkpoints"
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(GOLD, "ref_ckpt_w64.tar"), d / "000002.tar")
    return str(tmp_path)


def _check_resumed_state(ckpt, kw, start, grad_vars, optim):
    assert start == 2 == ckpt["global_step"]
    for key, net in (("network_fn_state_dict", kw["network_fn"]), ("network_fine_state_dict", kw["network_fine"])):
        sd = net.state_dict()
        assert list(sd) == list(ckpt[key])
        for k in sd:
            assert torch.equal(sd[k].cpu(), ckpt[key][k]), k
    ref_opt = ckpt["optimizer_state_dict"]
    assert len(ref_opt["state"]) == len(grad_vars) == 48
    for i, p in enumerate(grad_vars):
        st, rs = optim.state[p], ref_opt["state"][i]
        assert float(st["step"]) == float(rs["step"]) == 2.0
        assert torch.equal(st["exp_avg"].cpu(), rs["exp_avg"]) and torch.equal(st["exp_avg_sq"].cpu(), rs["exp_avg_sq"])
    g, rg = optim.param_groups[0], ref_opt["param_groups"][0]
    assert g["lr"] == rg["lr"] == LRATE * (0.1 ** (2 / (LRATE_DECAY * 1000)))
    assert tuple(g["betas"]) == tuple(rg["betas"]) and g["eps"] == rg["eps"]


def test_create_nerf_resumes_from_reference_written_checkpoint_cpu(tmp_path):
    """create_nerf(no_reload=False) finds and loads the reference's own .tar: start, every weight, Adam's
    exp_avg / exp_avg_sq / step and the decayed learning rate are the reference's (host-side logic: no kernel runs)."""
    import nerf_projects_b200 as nb
    base = _place_reference_checkpoint(tmp_path)
    ckpt = torch.load(os.path.join(GOLD, "ref_ckpt_w64.tar"), map_location="cpu")
    assert sorted(ckpt) == ["global_step", "network_fine_state_dict", "network_fn_state_dict", "optimizer_state_dict"]
    kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(_args(base), device=torch.device("cpu"))
    _check_resumed_state(ckpt, kw_train, start, grad_vars, optim)
    # and what we save is loadable by torch.optim.Adam (the optimizer the reference builds) with the same state layout
    ours = optim.state_dict()
    ref_adam = torch.optim.Adam(params=[torch.nn.Parameter(p.detach().clone()) for p in grad_vars], lr=LRATE, betas=(0.9, 0.999))
    ref_adam.load_state_dict(ours)
    assert sorted(ours["state"][0]) == sorted(ckpt["optimizer_state_dict"]["state"][0])


@pytest.mark.gpu
def test_resume_from_reference_checkpoint_next_step_matches_reference(tmp_path):
    """Resume from the reference-written checkpoint on the GPU and take the next optimisation step (fp32 validation
    mode: the 64-wide networks are not the canonical tensor-core architecture): loss and every parameter match what the
    REFERENCE produced when it resumed from the same file (ref_ckpt_w64.npz)."""
    import nerf_projects_b200 as nb
    g = gold("ref_ckpt_w64")
    base = _place_reference_checkpoint(tmp_path)
    ckpt = torch.load(os.path.join(GOLD, "ref_ckpt_w64.tar"), map_location="cpu")
    nb.set_precision("fp32")
    try:
        kw_train, kw_test, start, grad_vars, optim = nb.create_nerf(_args(base))
        _check_resumed_state(ckpt, kw_train, start, grad_vars, optim)
        kw = dict(kw_train); kw.update(near=2.0, far=6.0)
        from oracle import nerf_oracle as O
        K, _ = O.blender_camera(800, 800, 30.0)
        ro, rd, target = T(g["rays_o"]).cuda(), T(g["rays_d"]).cuda(), T(g["target"]).cuda()
        rgb, disp, acc, extras = nb.render(800, 800, K, chunk=32768, rays=(ro, rd), retraw=True, **kw)
        loss = nb.img2mse(rgb, target) + nb.img2mse(extras["rgb0"], target)
        optim.zero_grad(); loss.backward(); optim.step()
    finally:
        nb.set_precision("bf16")
    assert abs(float(loss) - float(g["losses"][2])) <= 2e-6, (float(loss), g["losses"])
    worst = 0.0
    names = [f"c.{n}" for n, _ in kw_train["network_fn"].named_parameters()] + \
            [f"f.{n}" for n, _ in kw_train["network_fine"].named_parameters()]
    for name, p in zip(names, grad_vars):
        want = T(g["after3." + name])
        d = (p.detach().cpu() - want).abs().max().item()
        worst = max(worst, d)
        # one Adam step moves an entry by at most ~lr = 5e-4; agreement to 2 % of that on every entry
        assert d <= 1e-5, f"{name}: {d}"
    print(f"resume + step 3: loss {float(loss):.8f} vs reference {float(g['losses'][2]):.8f}; max |param - reference| {worst:.2e}")


def keras_weight_list(seed=4242):
    """the seeded Keras-layout list of tools/make_golden_ckpt.py: kernel [in, out], bias [out]"""
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = [(63, 256)] + [(256, 256)] * 4 + [(319, 256)] + [(256, 256)] * 2 + [(256, 256), (283, 128), (128, 3), (256, 1)]
    weights = []
    for fi, fo in dims:
        weights.append((rng.uniform(-1, 1, size=(fi, fo)) / np.sqrt(fi)).astype(np.float32))
        weights.append(rng.uniform(-0.05, 0.05, size=(fo,)).astype(np.float32))
    x = rng.uniform(-1, 1, size=(32, 90)).astype(np.float32)
    return weights, x


def _check_keras_state(sd, g):
    for name, t in sd.items():
        t = t.detach().cpu().double()
        assert list(t.shape) == list(g[f"sd.{name}.shape"]), name
        assert abs(float(t.sum()) - float(g[f"sd.{name}.sum"])) <= 1e-9 * max(1.0, float(g[f"sd.{name}.abssum"])), name
        assert abs(float(t.abs().sum()) - float(g[f"sd.{name}.abssum"])) <= 1e-9 * float(g[f"sd.{name}.abssum"]), name
        assert np.array_equal(t.reshape(-1)[:8].float().numpy(), g[f"sd.{name}.first"]), name
        assert np.array_equal(t.reshape(-1)[-8:].float().numpy(), g[f"sd.{name}.last"]), name


def test_load_weights_from_keras_matches_reference_cpu():
    """NeRF.load_weights_from_keras (nerf.py:113-146): same index map and transposes as the reference (state_dict digests
    of the reference's own result), and the guard for use_viewdirs=False."""
    import nerf_projects_b200 as nb
    g = gold("keras_weights")
    weights, x = keras_weight_list(int(g["seed"]))
    assert np.array_equal(x, g["x"])
    m = nb.NeRF(8, 256, 63, 27, 5, [4], True)
    m.load_weights_from_keras(weights)
    _check_keras_state(m.state_dict(), g)
    with pytest.raises(AssertionError):
        nb.NeRF(8, 256, 63, 0, 4, [4], False).load_weights_from_keras(weights)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
def test_keras_weights_forward_on_gpu(precision, tol):
    """The network loaded from Keras weights evaluates like the reference's (module forward in fp32; the fused
    tensor-core kernel through run_network in bf16, which must see the NEW weights: packed-image cache invalidated)."""
    import nerf_projects_b200 as nb
    g = gold("keras_weights")
    weights, x = keras_weight_list(int(g["seed"]))
    m = nb.NeRF(8, 256, 63, 27, 5, [4], True).cuda()
    with torch.no_grad():
        y0 = m(T(x).cuda())                       # random init: forces a packed-weights cache entry in bf16 mode below
    m.load_weights_from_keras(weights)
    _check_keras_state(m.state_dict(), g)
    with torch.no_grad():
        y = m(T(x).cuda())
    assert (y.cpu() - T(g["y"])).abs().max().item() <= 2e-5
    assert not torch.allclose(y, y0)
    if precision == "bf16":
        # fused kernel on explicit (pre-embedding) inputs: compare with the oracle on the same points
        from oracle import nerf_oracle as O
        e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0)
        rng = np.random.Generator(np.random.PCG64(5))
        pts = T(rng.uniform(-1.5, 1.5, size=(8, 16, 3)).astype(np.float32)).cuda()
        vd = torch.nn.functional.normalize(T(rng.normal(size=(8, 3)).astype(np.float32)), dim=-1).cuda()
        nb.set_precision("bf16")
        with torch.no_grad():
            m2 = nb.NeRF(8, 256, 63, 27, 5, [4], True).cuda()
            before = nb.run_network(pts, vd, m2, e, ed)
            m2.load_weights_from_keras(weights)
            after = nb.run_network(pts, vd, m2, e, ed)
        params = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        want = O.run_network(pts.cpu(), vd.cpu(), O.to_torch(params), multires=10, multires_views=4)
        assert not torch.allclose(before, after), "fused kernel still uses the weights packed before load_weights_from_keras"
        assert (after.cpu() - want).abs().max().item() <= tol

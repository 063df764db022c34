"""Stage-by-stage comparison of the CUDA pipeline with the CPU oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import nerf_projects_b200 as nb
from nerf_projects_b200 import ops, mlp
from oracle import nerf_oracle as O
from _cases import gold, case_params, T, RENDER_CASES

tag = sys.argv[1] if len(sys.argv) > 1 else "render_c1_sharp"
g = gold(tag)
pc, pf = case_params(tag)
c = RENDER_CASES[tag]
vd = c["use_viewdirs"]; S = c["N_samples"]; NI = c["N_importance"]; white = c["white_bkgd"]
och = pc["output_linear.weight"].shape[0] if not vd else 5
net_c = nb.NeRF(8, 256, 63, 27 if vd else 0, och, [4], vd).cuda(); net_c.load_state_dict(O.to_torch(pc))
net_f = nb.NeRF(8, 256, 63, 27 if vd else 0, och, [4], vd).cuda(); net_f.load_state_dict(O.to_torch(pf))
rays, _ = O.pack_rays(T(g["rays_o"]), T(g["rays_d"]), float(g["near"]), float(g["far"]), H=int(g["H"]), W=int(g["W"]), K=g["K"], ndc=c["ndc"], use_viewdirs=vd)
ref = O.render_rays(rays, O.to_torch(pc), O.to_torch(pf), N_samples=S, N_importance=NI, white_bkgd=white, lindisp=c["lindisp"], retraw=True)
R = rays.shape[0]
rc = rays.cuda()
z_c = ops.coarse_z(rc, S, c["lindisp"])
print("z_c equal:", torch.equal(z_c.cpu(), ref["z_coarse"]))
for prec in ("fp32", "bf16"):
    nb.set_precision(prec)
    raw_c = mlp.forward(net_c, R, S, rc, z_c).view(R, S, -1)
    e = (raw_c.cpu() - ref["raw0"]).abs()
    print(f"[{prec}] raw_c err: rgb max {e[...,:3].max():.2e}  sigma max {e[...,3].max():.2e} (sigma range {ref['raw0'][...,3].min():.1f}..{ref['raw0'][...,3].max():.1f})")
    rgb0, disp0, acc0, w0, depth0 = ops.raw2outputs_fwd(raw_c, z_c, rc[:, 3:6], None, white)
    print(f"[{prec}] weights0 err max {(w0.cpu()-ref['weights0']).abs().max():.2e}  rgb0 err {(rgb0.cpu()-ref['rgb0']).abs().max():.2e} acc0 err {(acc0.cpu()-ref['acc0']).abs().max():.2e}")
    # compositing alone on the oracle's raw
    r2 = ops.raw2outputs_fwd(ref["raw0"].cuda().contiguous(), z_c, rc[:, 3:6], None, white)
    print(f"[{prec}] composite(oracle raw) weights err {(r2[3].cpu()-ref['weights0']).abs().max():.2e} rgb err {(r2[0].cpu()-ref['rgb0']).abs().max():.2e}")
    z_f, z_s, z_std, inds = ops.hierarchical_z(z_c, w0, NI, True, None, want_inds=True)
    dz = (z_s.cpu() - ref["z_samples"]).abs()
    print(f"[{prec}] z_samples: {(dz>1e-5).sum().item()} of {dz.numel()} differ by >1e-5, max {dz.max():.3e}; inds differ {(inds.cpu()!=ref['inds']).sum().item()}")
    # hierarchical on the oracle's weights (should be bit exact)
    z_f2, z_s2, _, inds2 = ops.hierarchical_z(z_c, ref["weights0"].cuda().contiguous(), NI, True, None, want_inds=True)
    print(f"[{prec}] hier(oracle weights): samples equal {torch.equal(z_s2.cpu(), ref['z_samples'])} z_f equal {torch.equal(z_f2.cpu(), ref['z_vals'])}")
    raw_f = mlp.forward(net_f, R, S+NI, rc, z_f2).view(R, S+NI, -1)
    e = (raw_f.cpu() - ref["raw"]).abs()
    print(f"[{prec}] raw_f (oracle z) err: rgb max {e[...,:3].max():.2e} sigma max {e[...,3].max():.2e}")
    rgb, disp, acc, w, depth = ops.raw2outputs_fwd(raw_f, z_f2, rc[:, 3:6], None, white)
    print(f"[{prec}] final (oracle z): rgb err {(rgb.cpu()-ref['rgb_map']).abs().max():.2e} acc err {(acc.cpu()-ref['acc_map']).abs().max():.2e}")
    raw_f = mlp.forward(net_f, R, S+NI, rc, z_f).view(R, S+NI, -1)
    rgb, disp, acc, w, depth = ops.raw2outputs_fwd(raw_f, z_f, rc[:, 3:6], None, white)
    er = (rgb.cpu()-ref['rgb_map']).abs().max(-1)[0]
    print(f"[{prec}] final (own z): rgb err max {er.max():.2e}  rays>1e-4: {(er>1e-4).sum().item()}/{R}  acc err {(acc.cpu()-ref['acc_map']).abs().max():.2e}")
    # where do the errors come from: last-sample sign?
    print('depth err', (depth.cpu()-ref['depth_map']).abs().max().item(), 'z range', float(ref['z_vals'].min()), float(ref['z_vals'].max()))
    sl = ref["raw"][:, -1, 3]
    print(f"[{prec}] min |sigma_last| fine {sl.abs().min():.3e}; coarse {ref['raw0'][:,-1,3].abs().min():.3e}")

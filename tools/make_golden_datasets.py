#!/usr/bin/env python
"""Generate tests/golden/datasets.npz by running the UNMODIFIED reference loaders
(/root/reference/nerf/load_blender.py, load_llff.py) on the seeded directory trees of oracle/datasets_synth.py.
Build container only.  ``imageio`` (which the reference imports) is not installed here, so it is satisfied by a
shim whose ``imread`` decodes with OpenCV -- lossless PNG, identical pixels; no reference code is modified or copied.

    python tools/make_golden_datasets.py
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/nerf")

from oracle import datasets_synth as DS  # noqa: E402


def _imageio_shim():
    import cv2

    def imread(path, *a, **k):
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img.ndim == 3 and img.shape[2] == 4:
            return cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB) if img.ndim == 3 else img
    m = types.ModuleType("imageio")
    m.imread = imread
    v2 = types.ModuleType("imageio.v2")
    v2.imread = imread
    m.v2 = v2
    sys.modules["imageio"], sys.modules["imageio.v2"] = m, v2


def main():
    _imageio_shim()
    import load_blender as LB
    import load_llff as LL
    out = {}
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        b = DS.write_blender(os.path.join(tmp, "blender"))
        for tag, kw in (("b_full", dict(half_res=False, testskip=1)), ("b_half", dict(half_res=True, testskip=2)),
                        ("b_skip0", dict(half_res=False, testskip=0))):
            imgs, poses, rp, hwf, i_split = LB.load_blender_data(b, **kw)
            out.update({f"{tag}.imgs": imgs, f"{tag}.poses": poses, f"{tag}.render_poses": rp.numpy(),
                        f"{tag}.hwf": np.array(hwf, np.float64), f"{tag}.i_train": i_split[0],
                        f"{tag}.i_val": i_split[1], f"{tag}.i_test": i_split[2]})
        ll = DS.write_llff(os.path.join(tmp, "llff"))
        for tag, kw in (("l_default", dict(factor=4)), ("l_spherify", dict(factor=4, spherify=True)),
                        ("l_norecenter", dict(factor=4, recenter=False, bd_factor=None))):
            images, poses, bds, rp, i_test = LL.load_llff_data(ll, **kw)
            out.update({f"{tag}.images": images, f"{tag}.poses": poses, f"{tag}.bds": bds, f"{tag}.render_poses": rp,
                        f"{tag}.i_test": np.int64(i_test)})
    import load_LINEMOD as LM
    import load_deepvoxels as DV
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        # LINEMOD frames carry full paths: a fixed location so that the fixture does not depend on the temp name
        lm = DS.write_linemod(os.path.join(tmp, "linemod"))
        for tag, kw in (("m_full", dict(half_res=False, testskip=1)), ("m_half", dict(half_res=True, testskip=2))):
            imgs, poses, rp, hwf, K, i_split, near, far = LM.load_LINEMOD_data(lm, **kw)
            out.update({f"{tag}.imgs": imgs, f"{tag}.poses": poses, f"{tag}.hwf": np.array(hwf, np.float64),
                        f"{tag}.K": np.array(K, np.float64), f"{tag}.near": np.float64(near), f"{tag}.far": np.float64(far),
                        f"{tag}.i_test": i_split[2]})
        dv = DS.write_deepvoxels(os.path.join(tmp, "dv"))
        for tag, skip in (("d_skip8", 8), ("d_skip3", 3)):
            imgs, poses, rp, hwf, i_split = DV.load_dv_data(scene="cube", basedir=dv, testskip=skip)
            out.update({f"{tag}.imgs": imgs, f"{tag}.poses": poses, f"{tag}.render_poses": rp,
                        f"{tag}.hwf": np.array(hwf, np.float64), f"{tag}.i_val": i_split[1], f"{tag}.i_test": i_split[2]})
    out["pose_spherical"] = np.stack([LB.pose_spherical(a, p, r).numpy() for a, p, r in
                                      ((0., -30., 4.), (37.5, -30., 4.), (-180., 12., 2.5), (91., 0., 1.))])
    path = os.path.join(ROOT, "tests", "golden", "datasets.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: v.shape for k, v in out.items() if "imgs" in k or "images" in k})


if __name__ == "__main__":
    main()

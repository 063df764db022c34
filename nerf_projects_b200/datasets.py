"""Dataset formats on the input side of the path: the Blender-synthetic and LLFF
loaders the reference's ``train()`` calls (ipynb: ``load_blender_data`` / ``load_llff_data``), same names,
arguments and return tuples as /root/reference/nerf/load_blender.py:37-91 and load_llff.py:255-318.

Host-side numpy: files are decoded once, the arrays then go to the GPU (``RayBatcher`` keeps images and poses
resident and generates ray batches there).  PNG/JPG decoding uses OpenCV (``imageio`` is not in this image;
decoded 8-bit pixels are identical).  ``_minify`` of the reference shells out to ImageMagick ``mogrify``; here a
missing ``images_<factor>`` directory is produced with ``cv2.resize(INTER_AREA)`` instead (same geometry, not
bit-identical pixels) -- pre-minified directories are used as they are.
``load_LINEMOD_data`` (load_LINEMOD.py:36-95) and ``load_dv_data`` (load_deepvoxels.py:6-110) complete the dataset
types ``train()`` dispatches on.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

_IMG_EXT = ("JPG", "jpg", "png")


def _imread(path: str) -> np.ndarray:
    """8-bit image as RGB(A), like imageio.imread."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3 and img.shape[2] == 4:
        return cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
    if img.ndim == 3:
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    return img


# ------------------------------------------------------------------ Blender

def pose_spherical(theta, phi, radius):
    """load_blender.py:29-34: camera on a sphere looking at the origin -> [4,4] c2w (torch fp32).
    translate(z=radius), rotate by phi about x, by theta about y, then swap to the Blender world axes."""
    ph, th = phi / 180. * np.pi, theta / 180. * np.pi
    T = torch.eye(4)
    T[2, 3] = radius
    Rx = torch.eye(4)
    Rx[1, 1], Rx[1, 2], Rx[2, 1], Rx[2, 2] = np.cos(ph), -np.sin(ph), np.sin(ph), np.cos(ph)
    Ry = torch.eye(4)
    Ry[0, 0], Ry[0, 2], Ry[2, 0], Ry[2, 2] = np.cos(th), -np.sin(th), np.sin(th), np.cos(th)
    axes = torch.tensor([[-1., 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return axes @ (Ry @ (Rx @ T))


def load_blender_data(basedir, half_res=False, testskip=1):
    """load_blender.py:37-91 -> (imgs [N,H,W,4] in [0,1], poses [N,4,4], render_poses [40,4,4] torch,
    [H, W, focal], [i_train, i_val, i_test]).  All training frames; every ``testskip``-th val/test frame."""
    imgs, poses, i_split, start = [], [], [], 0
    meta = None
    for split in ("train", "val", "test"):
        with open(os.path.join(basedir, f"transforms_{split}.json")) as fp:
            meta = json.load(fp)
        step = 1 if (split == "train" or testskip == 0) else testskip
        frames = meta["frames"][::step]
        imgs.append((np.array([_imread(os.path.join(basedir, f["file_path"] + ".png")) for f in frames]) / 255.)
                    .astype(np.float32))
        poses.append(np.array([f["transform_matrix"] for f in frames]).astype(np.float32))
        i_split.append(np.arange(start, start + len(frames)))
        start += len(frames)
    imgs, poses = np.concatenate(imgs, 0), np.concatenate(poses, 0)
    H, W = imgs[0].shape[:2]
    focal = .5 * W / np.tan(.5 * float(meta["camera_angle_x"]))     # the last split's meta, as the reference
    render_poses = torch.stack([pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, 40 + 1)[:-1]], 0)
    if half_res:
        import cv2
        H, W, focal = H // 2, W // 2, focal / 2.
        half = np.zeros((imgs.shape[0], H, W, 4))                    # float64, as the reference returns
        for i, im in enumerate(imgs):
            half[i] = cv2.resize(im, (W, H), interpolation=cv2.INTER_AREA)
        imgs = half
    return imgs, poses, render_poses, [H, W, focal], i_split


# --------------------------------------------------------------------- LLFF

def normalize(x):
    return x / np.linalg.norm(x)


def viewmatrix(z, up, pos):
    """[3,4] camera frame with forward axis z, approximate up vector, at pos (load_llff.py:123-129)."""
    fwd = normalize(z)
    right = normalize(np.cross(up, fwd))
    true_up = normalize(np.cross(fwd, right))
    return np.stack([right, true_up, fwd, pos], 1)


def poses_avg(poses):
    """Average camera [3,5] of a pose set [N,3,5] (load_llff.py:135-144)."""
    frame = viewmatrix(normalize(poses[:, :3, 2].sum(0)), poses[:, :3, 1].sum(0), poses[:, :3, 3].mean(0))
    return np.concatenate([frame, poses[0, :3, -1:]], 1)


def _to44(p34):
    """[...,3,4] -> [...,4,4] with a (0,0,0,1) bottom row."""
    bottom = np.broadcast_to(np.array([0., 0., 0., 1.]), p34.shape[:-2] + (1, 4))
    return np.concatenate([p34, bottom], -2)


def recenter_poses(poses):
    """Express every pose in the frame of the average camera (load_llff.py:161-173)."""
    out = poses + 0
    rel = np.linalg.inv(_to44(poses_avg(poses)[:3, :4])) @ _to44(poses[:, :3, :4])
    out[:, :3, :4] = rel[:, :3, :4]
    return out


def render_path_spiral(c2w, up, rads, focal, zdelta, zrate, rots, N):
    """Spiral of N cameras around c2w looking at a point `focal` ahead (load_llff.py:148-157)."""
    rads4 = np.array(list(rads) + [1.])
    look_at = np.dot(c2w[:3, :4], np.array([0, 0, -focal, 1.]))
    out = []
    for t in np.linspace(0., 2. * np.pi * rots, int(N) + 1)[:-1]:
        eye = np.dot(c2w[:3, :4], np.array([np.cos(t), -np.sin(t), -np.sin(t * zrate), 1.]) * rads4)
        out.append(np.concatenate([viewmatrix(normalize(eye - look_at), up, eye), c2w[:, 4:5]], 1))
    return out


def spherify_poses(poses, bds):
    """Inward-facing captures (load_llff.py:179-238): centre on the point closest to all optical axes, scale the
    cameras to the unit sphere, and build a 120-view circular path at the cameras' mean height."""
    d, o = poses[:, :3, 2:3], poses[:, :3, 3:4]
    A = np.eye(3) - d * np.transpose(d, [0, 2, 1])
    center = np.squeeze(-np.linalg.inv((np.transpose(A, [0, 2, 1]) @ A).mean(0)) @ (-A @ o).mean(0))
    ax_z = normalize((poses[:, :3, 3] - center).mean(0))
    ax_x = normalize(np.cross([.1, .2, .3], ax_z))
    ax_y = normalize(np.cross(ax_z, ax_x))
    world = np.stack([ax_x, ax_y, ax_z, center], 1)
    reset = np.linalg.inv(_to44(world[None])) @ _to44(poses[:, :3, :4])
    rad = np.sqrt(np.mean(np.sum(np.square(reset[:, :3, 3]), -1)))
    sc = 1. / rad
    reset[:, :3, 3] *= sc
    bds *= sc                      # in place, like the reference
    rad *= sc
    zh = np.mean(reset[:, :3, 3], 0)[2]
    rc = np.sqrt(rad ** 2 - zh ** 2)
    ring = []
    for th in np.linspace(0., 2. * np.pi, 120):
        eye = np.array([rc * np.cos(th), rc * np.sin(th), zh])
        fwd = normalize(eye)
        right = normalize(np.cross(fwd, np.array([0, 0, -1.])))
        ring.append(np.stack([right, normalize(np.cross(fwd, right)), fwd, eye], 1))
    ring = np.stack(ring, 0)
    hwf = poses[0, :3, -1:]
    ring = np.concatenate([ring, np.broadcast_to(hwf, ring[:, :3, -1:].shape)], -1)
    reset = np.concatenate([reset[:, :3, :4], np.broadcast_to(hwf, reset[:, :3, -1:].shape)], -1)
    return reset, ring, bds


def _list_images(d):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(_IMG_EXT)]


def _minify(basedir, factors=(), resolutions=()):
    """Create images_<f> / images_<W>x<H> when missing (load_llff.py:9-60; OpenCV instead of mogrify)."""
    import cv2
    src = [f for f in (os.path.join(basedir, "images", n) for n in sorted(os.listdir(os.path.join(basedir, "images"))))
           if f.endswith(("JPG", "jpg", "png", "jpeg", "PNG"))]
    for r in list(factors) + list(resolutions):
        name = f"images_{r}" if isinstance(r, int) else f"images_{r[1]}x{r[0]}"
        out = os.path.join(basedir, name)
        if os.path.exists(out):
            continue
        print("Minifying", r, basedir)
        os.makedirs(out)
        for f in src:
            im = cv2.imread(f, cv2.IMREAD_UNCHANGED)
            size = (int(round(im.shape[1] / r)), int(round(im.shape[0] / r))) if isinstance(r, int) else (r[1], r[0])
            cv2.imwrite(os.path.join(out, os.path.splitext(os.path.basename(f))[0] + ".png"),
                        cv2.resize(im, size, interpolation=cv2.INTER_AREA))


def _load_data(basedir, factor=None, width=None, height=None, load_imgs=True):
    """load_llff.py:65-119: poses_bounds.npy -> poses [3,5,N] (H, W, focal in column 4), bds [2,N], imgs [H,W,3,N]."""
    arr = np.load(os.path.join(basedir, "poses_bounds.npy"))
    poses = arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0])
    bds = arr[:, -2:].transpose([1, 0])
    sh = _imread(_list_images(os.path.join(basedir, "images"))[0]).shape
    sfx = ""
    if factor is not None:
        sfx = f"_{factor}"
        _minify(basedir, factors=[factor])
    elif height is not None:
        factor = sh[0] / float(height)
        width = int(sh[1] / factor)
        _minify(basedir, resolutions=[[height, width]])
        sfx = f"_{width}x{height}"
    elif width is not None:
        factor = sh[1] / float(width)
        height = int(sh[0] / factor)
        _minify(basedir, resolutions=[[height, width]])
        sfx = f"_{width}x{height}"
    else:
        factor = 1
    imgdir = os.path.join(basedir, "images" + sfx)
    if not os.path.exists(imgdir):
        print(imgdir, "does not exist, returning")
        return None
    files = _list_images(imgdir)
    if poses.shape[-1] != len(files):
        print("Mismatch between imgs {} and poses {} !!!!".format(len(files), poses.shape[-1]))
        return None
    sh = _imread(files[0]).shape
    poses[:2, 4, :] = np.array(sh[:2]).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1. / factor
    if not load_imgs:
        return poses, bds
    imgs = np.stack([_imread(f)[..., :3] / 255. for f in files], -1)
    print("Loaded image data", imgs.shape, poses[:, -1, 0])
    return poses, bds, imgs


def load_llff_data(basedir, factor=8, recenter=True, bd_factor=.75, spherify=False, path_zflat=False):
    """load_llff.py:255-318 -> (images [N,H,W,3] fp32, poses [N,3,5] fp32, bds [N,2], render_poses [M,3,5], i_test)."""
    poses, bds, imgs = _load_data(basedir, factor=factor)
    print("Loaded", basedir, bds.min(), bds.max())
    # LLFF stores [down, right, back]; NeRF wants [right, up, back]; image index to axis 0
    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    images = np.moveaxis(imgs, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1. if bd_factor is None else 1. / (bds.min() * bd_factor)   # nearest depth -> 1/bd_factor
    poses[:, :3, 3] *= sc
    bds *= sc
    if recenter:
        poses = recenter_poses(poses)
    if spherify:
        poses, render_poses, bds = spherify_poses(poses, bds)
    else:
        c2w = poses_avg(poses)
        print("recentered", c2w.shape)
        print(c2w[:3, :4])
        up = normalize(poses[:, :3, 1].sum(0))
        close_depth, inf_depth = bds.min() * .9, bds.max() * 5.
        dt = .75
        focal = 1. / ((1. - dt) / close_depth + dt / inf_depth)      # focus depth of the spiral
        zdelta = close_depth * .2
        rads = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
        n_views, n_rots = 120, 2
        if path_zflat:
            c2w[:3, 3] = c2w[:3, 3] + (-close_depth * .1) * c2w[:3, 2]
            rads[2] = 0.
            n_rots, n_views = 1, n_views / 2
        render_poses = render_path_spiral(c2w, up, rads, focal, zdelta, zrate=.5, rots=n_rots, N=n_views)
    render_poses = np.array(render_poses).astype(np.float32)
    c2w = poses_avg(poses)
    print("Data:")
    print(poses.shape, images.shape, bds.shape)
    i_test = np.argmin(np.sum(np.square(c2w[:3, 3] - poses[:, :3, 3]), -1))
    print("HOLDOUT view is", i_test)
    return images.astype(np.float32), poses.astype(np.float32), bds, render_poses, i_test


# ------------------------------------------------------------------ LINEMOD

def load_LINEMOD_data(basedir, half_res=False, testskip=1):
    """load_LINEMOD.py:36-95 -> (imgs, poses, render_poses, [H, W, focal], K, i_split, near, far).  Like the Blender
    format, but ``file_path`` is used as written, intrinsics come from the first frame's ``intrinsic_matrix`` and
    near / far from the train and test metadata (floor of the smaller near, ceil of the larger far)."""
    metas = {}
    for split in ("train", "val", "test"):
        with open(os.path.join(basedir, f"transforms_{split}.json")) as fp:
            metas[split] = json.load(fp)
    imgs, poses, i_split, start = [], [], [], 0
    for split in ("train", "val", "test"):
        step = 1 if (split == "train" or testskip == 0) else testskip
        frames = metas[split]["frames"][::step]
        if split == "test":
            for k, f in enumerate(frames):
                print(f"{k}th test frame: {f['file_path']}")
        imgs.append((np.array([_imread(f["file_path"]) for f in frames]) / 255.).astype(np.float32))
        poses.append(np.array([f["transform_matrix"] for f in frames]).astype(np.float32))
        i_split.append(np.arange(start, start + len(frames)))
        start += len(frames)
    imgs, poses = np.concatenate(imgs, 0), np.concatenate(poses, 0)
    H, W = imgs[0].shape[:2]
    K = metas["test"]["frames"][0]["intrinsic_matrix"]          # the last split read, as the reference
    focal = float(K[0][0])
    print(f"Focal: {focal}")
    render_poses = torch.stack([pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, 40 + 1)[:-1]], 0)
    if half_res:
        import cv2
        H, W, focal = H // 2, W // 2, focal / 2.
        half = np.zeros((imgs.shape[0], H, W, imgs.shape[-1]))
        for i, im in enumerate(imgs):
            half[i] = cv2.resize(im, (W, H), interpolation=cv2.INTER_AREA)
        imgs = half
    near = np.floor(min(metas["train"]["near"], metas["test"]["near"]))
    far = np.ceil(max(metas["train"]["far"], metas["test"]["far"]))
    return imgs, poses, render_poses, [H, W, focal], K, i_split, near, far


# --------------------------------------------------------------- DeepVoxels

def _dv_focal(path, side):
    """Focal length of intrinsics.txt rescaled to a `side`-pixel image (load_deepvoxels.py:9-48): line 1 = f cx cy,
    line 5 = height width."""
    with open(path) as f:
        lines = f.read().splitlines()
    focal = float(lines[0].split()[0])
    height = float(lines[4].split()[0])
    return side / height * focal


def _dv_poses(posedir):
    """All pose/*.txt (4x4, row-major) with the y and z camera axes flipped -> [N,3,4] fp32 (load_deepvoxels.py:64-75)."""
    files = [f for f in sorted(os.listdir(posedir)) if f.endswith("txt")]
    mats = np.stack([np.array([float(x) for x in open(os.path.join(posedir, f)).read().split()])
                     .reshape(4, 4).astype(np.float32) for f in files], 0)
    flip = np.diag([1., -1., -1., 1.])
    return (mats @ flip)[:, :3, :4].astype(np.float32)


def _dv_images(imgdir, step=1):
    files = [f for f in sorted(os.listdir(imgdir)) if f.endswith("png")][::step]
    return np.stack([_imread(os.path.join(imgdir, f)) / 255. for f in files], 0).astype(np.float32)


def load_dv_data(scene="cube", basedir="/data/deepvoxels", testskip=8):
    """load_deepvoxels.py:6-110 -> (imgs, poses [N,3,4], render_poses = test poses, [512, 512, focal], i_split);
    all training views, every ``testskip``-th validation / test view."""
    H = W = 512
    root = lambda split: "{}/{}/{}".format(basedir, split, scene)
    focal = _dv_focal(os.path.join(root("train"), "intrinsics.txt"), H)
    poses = _dv_poses(os.path.join(root("train"), "pose"))
    testposes = _dv_poses(os.path.join(root("test"), "pose"))[::testskip]
    valposes = _dv_poses(os.path.join(root("validation"), "pose"))[::testskip]
    parts = [_dv_images(os.path.join(root("train"), "rgb")), _dv_images(os.path.join(root("validation"), "rgb"), testskip),
             _dv_images(os.path.join(root("test"), "rgb"), testskip)]
    counts = np.cumsum([0] + [x.shape[0] for x in parts])
    i_split = [np.arange(counts[i], counts[i + 1]) for i in range(3)]
    imgs = np.concatenate(parts, 0)
    poses = np.concatenate([poses, valposes, testposes], 0)
    print(poses.shape, imgs.shape)
    return imgs, poses, testposes, [H, W, focal], i_split

"""TEST INFRASTRUCTURE ONLY: writes tiny seeded Blender-synthetic and LLFF directory trees (the on-disk formats
load_blender.py / load_llff.py read) so loader parity can be checked without real datasets."""
import json
import os

import numpy as np


def write_blender(root, H=12, W=16, n=(5, 4, 6), seed=0):
    import cv2
    rng = np.random.Generator(np.random.PCG64(seed))
    os.makedirs(root, exist_ok=True)
    for split, cnt in zip(("train", "val", "test"), n):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(cnt):
            rgba = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
            cv2.imwrite(os.path.join(root, split, f"r_{i}.png"), cv2.cvtColor(rgba, cv2.COLOR_RGBA2BGRA))
            m = np.eye(4)
            m[:3, :4] = rng.normal(size=(3, 4))
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.1, "transform_matrix": m.tolist()})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": 0.6911112070083618 + 0.01 * len(split), "frames": frames}, f)
    return root


def write_llff(root, H=24, W=32, n=7, factor=4, seed=1):
    import cv2
    rng = np.random.Generator(np.random.PCG64(seed))
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, f"images_{factor}"), exist_ok=True)
    rows = []
    for i in range(n):
        big = rng.integers(0, 256, size=(H * factor, W * factor, 3), dtype=np.uint8)
        small = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        cv2.imwrite(os.path.join(root, "images", f"img_{i:03d}.png"), big)
        cv2.imwrite(os.path.join(root, f"images_{factor}", f"img_{i:03d}.png"), small)
        # forward-facing cameras jittered around the origin looking down -z (LLFF [down, right, back] convention)
        ang = rng.normal(scale=0.08, size=3)
        Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
        Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
        R = Ry @ Rx                                   # columns: right, up, back
        llff_R = np.stack([-R[:, 1], R[:, 0], R[:, 2]], 1)
        t = rng.normal(scale=0.4, size=3)
        pose = np.concatenate([llff_R, t[:, None], np.array([[H * factor], [W * factor], [815.0]])], 1)
        rows.append(np.concatenate([pose.reshape(-1), [1.2 + rng.uniform(0, 0.5), 9.0 + rng.uniform(0, 6)]]))
    np.save(os.path.join(root, "poses_bounds.npy"), np.array(rows))
    return root


def write_linemod(root, H=10, W=14, n=(4, 3, 5), seed=2):
    """LINEMOD-format tree: like Blender, but file_path is a full path, frames carry intrinsic_matrix, metas carry near/far."""
    import cv2
    rng = np.random.Generator(np.random.PCG64(seed))
    os.makedirs(root, exist_ok=True)
    for si, (split, cnt) in enumerate(zip(("train", "val", "test"), n)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(cnt):
            rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
            path = os.path.join(root, split, f"{i:04d}.png")
            cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
            m = np.eye(4)
            m[:3, :4] = rng.normal(size=(3, 4))
            Km = [[572.4 + si, 0.0, 325.3], [0.0, 573.6, 242.0], [0.0, 0.0, 1.0]]
            frames.append({"file_path": path, "transform_matrix": m.tolist(), "intrinsic_matrix": Km})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"near": 0.31 + 0.2 * si, "far": 1.42 + 0.3 * si, "frames": frames}, f)
    return root


def write_deepvoxels(root, scene="cube", side=9, n=(5, 9, 10), seed=3):
    """DeepVoxels-format tree: {train,validation,test}/<scene>/{intrinsics.txt, pose/*.txt, rgb/*.png}."""
    import cv2
    rng = np.random.Generator(np.random.PCG64(seed))
    for split, cnt in zip(("train", "validation", "test"), n):
        d = os.path.join(root, split, scene)
        os.makedirs(os.path.join(d, "pose"), exist_ok=True)
        os.makedirs(os.path.join(d, "rgb"), exist_ok=True)
        with open(os.path.join(d, "intrinsics.txt"), "w") as f:
            f.write("525.0 256.0 256.0 0.\n0. 0. 0.\n1.\n0.5\n480 640\n")
        for i in range(cnt):
            np.savetxt(os.path.join(d, "pose", f"{i:06d}.txt"), rng.normal(size=(1, 16)), fmt="%.8f")
            cv2.imwrite(os.path.join(d, "rgb", f"{i:06d}.png"), rng.integers(0, 256, size=(side, side, 3), dtype=np.uint8))
    return root

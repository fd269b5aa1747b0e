"""Generates tests/golden/augment.npz by RUNNING THE REFERENCE'S OWN RenderedDataset.__getitem__
(anakin/artiboost/rendered_dataset.py:155-274) on seeded inputs, with every random draw replaced by a recorded value.
Run in the build container:  python tests/golden/make_golden_augment.py

Patched in the reference's module namespaces (behaviour otherwise untouched):
  * prepare_essential -> returns our image / annotations instead of the pickle + render-queue round trip (:103-153);
  * torch.distributions Uniform / Normal -> replay `center_jit`, `scale_jit`, `rot_rad`, `blur_u` (:178-189,256);
  * img_augment.get_color_params / random.shuffle -> replay the four factors and the execution order (:6-45);
  * np.uint8 inside img_augment -> numpy-1.x wrap-around of negative hue shifts (numpy 2 raises; the reference pins
    numpy 1.x), and transform_img records the inverse affine coefficients it hands to Pillow.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from anakin.artiboost import rendered_dataset as RD  # noqa: E402
from anakin.utils import img_augment as IA  # noqa: E402

from oracle import augment as A  # noqa: E402  (only for OPS naming)

N = 12
RAW, OUT = (96, 80), (64, 64)  # (W, H) of the rendered image / the network input: small, non-square raw size
K = np.array([[81.5, 0, 48.0], [0, 81.5, 40.0], [0, 0, 1]], np.float64)


class U8(np.uint8):
    def __new__(cls, v=0):
        return np.uint8(int(np.trunc(v)) & 0xFF)


def main():
    rng = np.random.RandomState(7)
    npx = types.ModuleType("npx")
    npx.__dict__.update(np.__dict__)
    npx.uint8 = U8
    IA.np = npx
    state = {}

    class FakeDist:
        def __init__(self, *a, **k):
            pass

        def sample(self, shape=()):
            v = state["queue"].pop(0)
            return torch.as_tensor(v)

    RD.Uniform = RD.Normal = FakeDist
    IA.get_color_params = lambda **k: state["color"]

    class FakeRandom:
        @staticmethod
        def shuffle(lst):
            order = state["order"]            # execution order as indices into the list-building order
            lst[:] = [lst[i] for i in order]

    IA.random = FakeRandom
    orig_transform_img = IA.transform_img

    def recording_transform_img(img, affine_trans, res):
        rev = np.linalg.inv(affine_trans)
        state["rev"] = np.array([rev[0, 0], rev[0, 1], rev[0, 2], rev[1, 0], rev[1, 1], rev[1, 2]], np.float64)
        state["pre_warp"] = np.array(img)
        return orig_transform_img(img, affine_trans, res)

    IA.transform_img = recording_transform_img

    ds = RD.RenderedDataset.__new__(RD.RenderedDataset)
    ds.cam_intr, ds.image_size, ds.raw_size, ds.center_idx = K, list(OUT), list(RAW), 0
    ds.bbox_expand_ratio, ds.require_full_image, ds.crop_model = 1.2, False, "root_obj"
    ds.aug = True
    ds.hue, ds.saturation, ds.contrast, ds.brightness, ds.blur_radius = 0.075, 0.1, 0.1, 0.1, 0.1
    ds.scale_jittering, ds.center_jittering, ds.max_rot = 0.1, 0.1, 0.2 * np.pi
    ds.sides, ds.njoints, ds.ncorners = "right", 21, 8
    ds.obj_map = {"obj": 3}

    rec = {k: [] for k in ("img", "joints", "pose", "corners_can", "center_jit", "scale_jit", "rot_rad", "rot_cs", "blur_radius",
                           "factors", "order", "rev", "pre_warp", "image", "cam_intr", "root_joint", "joints_3d", "joints_2d",
                           "joints_vis", "corners_3d", "corners_2d", "corners_vis", "obj_transf")}
    for i in range(N):
        # a blocky image with sharp edges (exercises the blur rounding) plus noise
        blocks = rng.randint(0, 2, size=(RAW[1] // 8 + 1, RAW[0] // 8 + 1, 3)) * 200
        img = np.kron(blocks, np.ones((8, 8, 1)))[:RAW[1], :RAW[0]] + rng.randint(0, 56, size=(RAW[1], RAW[0], 3))
        img = img.astype(np.uint8)
        root = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.45, 0.55)])
        joints = (root + rng.normal(0, 0.04, size=(21, 3))).astype(np.float32)
        if i == 3:
            joints[:, 0] += 1.0            # hand far outside the raw image: joints_vis must be all zero
        ang = rng.uniform(-1, 1, 3)
        Rm = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = Rm * np.sign(np.linalg.det(Rm))
        pose[:3, 3] = root + rng.normal(0, 0.02, 3)
        corners_can = (np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]) * rng.uniform(0.03, 0.08, 3)).astype(np.float32)
        center_jit = rng.uniform(-1, 1, 2).astype(np.float32)
        scale_jit = np.float32(rng.normal(0, 0.1 / 3.0) * (4.0 if i == 5 else 1.0))  # i == 5: clipped
        rot_rad = float(np.float32(rng.uniform(-0.2 * np.pi, 0.2 * np.pi)))
        blur_u = float(np.float32(rng.uniform(0, 1)))
        factors = [float(np.float32(rng.uniform(0.9, 1.1))), float(np.float32(rng.uniform(0.9, 1.1))),
                   float(np.float32(rng.uniform(0.9, 1.1))), float(np.float32(rng.uniform(-0.075, 0.075)))]  # B, C, S, H
        order = rng.permutation(4)

        def essentials(index, img=img, joints=joints, pose=pose, corners_can=corners_can):
            j2 = (K @ joints.T).T
            j2 = j2[:, 0:2] / (j2[:, 2:3] + 1e-8)
            c3 = (pose[:3, :3] @ corners_can.T).T + pose[:3, 3]
            c2 = (K @ c3.T).T
            c2 = c2[:, 0:2] / (c2[:, 2:3] + 1e-8)
            return ({"img": img, "hand_joints_3d": joints, "hand_joints_2d": j2, "obj_corners_can": corners_can, "obj_corners_3d": c3,
                     "obj_corners_2d": c2, "cam_intr": K, "hand_side": "right", "pose": pose, "objname": "obj"},
                    {"obj_id": 0, "persp_id": 1, "grasp_id": 2})

        ds.prepare_essential = essentials
        state["queue"] = [center_jit, scale_jit, rot_rad, blur_u]
        state["color"] = tuple(factors)
        state["order"] = [int(o) for o in order]
        s = ds[i]
        rec["img"].append(img); rec["joints"].append(joints); rec["pose"].append(pose); rec["corners_can"].append(corners_can)
        rec["center_jit"].append(center_jit); rec["scale_jit"].append(scale_jit); rec["rot_rad"].append(rot_rad)
        rec["rot_cs"].append(np.array([np.cos(rot_rad), np.sin(rot_rad)], np.float32))
        rec["blur_radius"].append(np.float32(blur_u * 0.1)); rec["factors"].append(np.array(factors, np.float32))
        rec["order"].append(order.astype(np.int32)); rec["rev"].append(state["rev"]); rec["pre_warp"].append(state["pre_warp"])
        rec["image"].append(s["image"].numpy())
        for k in ("cam_intr", "root_joint", "joints_3d", "joints_2d", "joints_vis", "corners_3d", "corners_2d", "corners_vis", "obj_transf"):
            rec[k].append(np.asarray(s[k], np.float32))
    out = {k: np.stack(v) for k, v in rec.items()}
    path = os.path.join(HERE, "augment.npz")
    np.savez_compressed(path, K=K, raw_size=np.array(RAW), out_size=np.array(OUT), **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

"""GPU: ab_crop_augment (csrc/augment.cu) through artiboost_b200.artiboost.RenderedDataset vs the oracle
(oracle/augment.py, pinned against Pillow and the reference's __getitem__).  Image bytes exact; annotations fp32."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import augment as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG_DATASET = {"AUG": True, "AUG_PARAM": {"SCALE_JIT": 0.1, "CENTER_JIT": 0.1, "MAX_ROT": 0.2}}
ANN = ("cam_intr", "root_joint", "joints_3d", "joints_2d", "corners_3d", "corners_2d", "obj_transf")


def preset(out_size, center_idx=0, crop_model="root_obj", full=False):
    return {"IMAGE_SIZE": list(out_size), "CENTER_IDX": center_idx, "BBOX_EXPAND_RATIO": 1.2, "FULL_IMAGE": full, "CROP_MODEL": crop_model}


def make_inputs(rng, B, raw, K, far=()):
    W, H = raw
    imgs, joints, poses, ccs = [], [], [], []
    for i in range(B):
        blocks = rng.randint(0, 2, size=(H // 8 + 1, W // 8 + 1, 3)) * 200
        img = np.kron(blocks, np.ones((8, 8, 1)))[:H, :W] + rng.randint(0, 56, size=(H, W, 3))
        imgs.append(img.astype(np.uint8))
        root = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.45, 0.55)])
        j = (root + rng.normal(0, 0.04, size=(21, 3))).astype(np.float32)
        if i in far:
            j[:, 0] += 1.0
        joints.append(j)
        q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        p = np.eye(4, dtype=np.float32)
        p[:3, :3] = q * np.sign(np.linalg.det(q))
        p[:3, 3] = root + rng.normal(0, 0.02, 3)
        poses.append(p)
        ccs.append((np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]) * rng.uniform(0.03, 0.08, 3)).astype(np.float32))
    return np.stack(imgs), np.stack(joints), np.stack(poses), np.stack(ccs)


def run_gpu(ds, imgs, joints, poses, draws=None, order=None):
    B, H, W, _ = imgs.shape
    rgba = torch.from_numpy(np.concatenate([imgs, np.full((B, H, W, 1), 255, np.uint8)], -1)).to(DEV)
    views = {"rgba": rgba, "joints": torch.from_numpy(joints).to(DEV), "obj_pose": torch.from_numpy(poses).to(DEV),
             "obj_id": torch.arange(B, device=DEV, dtype=torch.int32), "persp_id": torch.zeros(B, device=DEV, dtype=torch.int32),
             "grasp_id": torch.zeros(B, device=DEV, dtype=torch.int32)}
    out = ds(views, draws=draws, order=order, return_affine=True)
    assert int(out["_status"]) == 0
    return {k: v.cpu().numpy() for k, v in out.items()}


def oracle_draws(draws, order, i):
    d, o = draws[i], order[i]
    return {"center_jit": d[0:2], "scale_jit": d[2], "rot_cs": d[3:5], "blur_radius": d[5], "brightness": d[6], "contrast": d[7],
            "saturation": d[8], "hue": d[9], "order": o}


@pytest.mark.parametrize("raw,out,crop_model,center_idx", [((96, 80), (64, 64), "root_obj", 0), ((128, 128), (112, 96), "hand_obj", 9),
                                                            ((256, 256), (224, 224), "hand", 0)])
def test_crop_augment_matches_oracle(lib_built, raw, out, crop_model, center_idx):
    from artiboost_b200.artiboost import RenderedDataset
    rng = np.random.RandomState(raw[0] + out[0])
    B = 10
    f = 217.5 * raw[0] / 256
    K = np.array([[f, 0, raw[0] / 2.0], [0, f, raw[1] / 2.0], [0, 0, 1]], np.float32)
    imgs, joints, poses, ccs = make_inputs(rng, B, raw, K, far=(2,))
    gen = torch.Generator(device=DEV).manual_seed(5)
    ds = RenderedDataset(ccs, K, CFG_DATASET, preset(out, center_idx, crop_model), raw_size=raw, device=DEV, generator=gen)
    draws, order = ds.draw(B)
    draws[1, 3], draws[1, 4] = 1.0, 0.0      # no rotation: Pillow's pure-scaling path (running double sums)
    draws[3, 5] = 0.0                        # blur radius 0: Pillow skips the filter
    draws[4, 5] = 0.1                        # the reference's maximum radius
    draws[5, 2] = 0.5                        # scale jitter beyond the clip
    got = run_gpu(ds, imgs, joints, poses, draws, order)
    dn, on = draws.cpu().numpy(), order.cpu().numpy()
    cfg = {"crop_model": crop_model, "bbox_expand_ratio": 1.2, "aug": True, "center_jit": 0.1, "scale_jit": 0.1, "center_idx": center_idx,
           "image_size": out, "raw_size": raw}
    for i in range(B):
        ref = A.rendered_sample(imgs[i], joints[i], poses[i], ccs[i], K, oracle_draws(dn, on, i), cfg)
        assert np.array_equal(got["affine"][i], ref["affine"][:2].reshape(-1)), f"forward affine [{i}]"
        assert np.array_equal(got["inv_affine"][i], ref["inv_affine"]), f"inverse affine [{i}]"
        assert np.array_equal(got["image"][i], ref["image"]), f"image [{i}]: {(got['image'][i] != ref['image']).sum()} values differ"
        for k in ANN:
            np.testing.assert_allclose(got[k][i], ref[k], rtol=1e-5, atol=1e-5, err_msg=f"{k}[{i}]")
        assert np.array_equal(got["joints_vis"][i], ref["joints_vis"]) and np.array_equal(got["corners_vis"][i], ref["corners_vis"])
    assert got["joints_vis"][2].sum() == 0  # hand outside the raw image
    assert set(np.unique(got["image"]).tolist()) != {-0.5}


def test_no_augmentation_and_full_image(lib_built):
    from artiboost_b200.artiboost import RenderedDataset
    rng = np.random.RandomState(2)
    raw = out = (64, 64)
    K = np.array([[54.0, 0, 32.0], [0, 54.0, 32.0], [0, 0, 1]], np.float32)
    imgs, joints, poses, ccs = make_inputs(rng, 4, raw, K)
    for full in (False, True):
        ds = RenderedDataset(ccs, K, {"AUG": False, "AUG_PARAM": None}, preset(out, full=full), raw_size=raw, device=DEV)
        got = run_gpu(ds, imgs, joints, poses)
        cfg = {"crop_model": "root_obj", "bbox_expand_ratio": 1.2, "aug": False, "center_idx": 0, "image_size": out, "raw_size": raw,
               "full_image": full}
        for i in range(4):
            ref = A.rendered_sample(imgs[i], joints[i], poses[i], ccs[i], K, {}, cfg)
            assert np.array_equal(got["image"][i], ref["image"])
            for k in ANN:
                np.testing.assert_allclose(got[k][i], ref[k], rtol=1e-5, atol=1e-5, err_msg=k)
        if full:  # the whole image, unchanged: identity warp
            ident = (imgs.astype(np.float32) / np.float32(255) - np.float32(0.5)).transpose(0, 3, 1, 2)
            assert np.array_equal(got["image"], ident)


def test_against_the_reference_fixture(lib_built):
    """CUDA vs the outputs recorded from the reference's own RenderedDataset.__getitem__ (tests/golden/augment.npz)."""
    from artiboost_b200.artiboost import RenderedDataset
    g = golden("augment.npz")
    raw, out = tuple(int(v) for v in g["raw_size"]), tuple(int(v) for v in g["out_size"])
    B = len(g["img"])
    ds = RenderedDataset(g["corners_can"], g["K"].astype(np.float32), CFG_DATASET, preset(out), raw_size=raw, device=DEV)
    f = g["factors"]
    draws = np.concatenate([g["center_jit"], g["scale_jit"][:, None], g["rot_cs"], g["blur_radius"][:, None], f[:, 0:1], f[:, 1:2], f[:, 2:3],
                            f[:, 3:4]], 1).astype(np.float32)
    got = run_gpu(ds, g["img"], g["joints"], g["pose"], torch.from_numpy(draws).to(DEV), torch.from_numpy(g["order"]).to(DEV))
    for k in ANN:
        np.testing.assert_allclose(got[k], g[k], rtol=2e-5, atol=2e-4 if k.endswith("2d") or k == "cam_intr" else 2e-6, err_msg=k)
    assert np.array_equal(got["joints_vis"], g["joints_vis"]) and np.array_equal(got["corners_vis"], g["corners_vis"])
    moved = (np.abs(got["image"] - g["image"]).max(1) > 0).mean()
    assert moved < 5e-3, moved  # LAPACK's fp32 inverse vs the closed form: a few source pixels on the boundary move


def test_argument_errors(lib_built):
    from artiboost_b200 import lib
    from artiboost_b200.artiboost import RenderedDataset
    K = np.eye(3, dtype=np.float32)
    with pytest.raises(lib.AbError):
        RenderedDataset(np.zeros((1, 8, 3)), K, CFG_DATASET, preset((64, 64)), device="cpu")
    with pytest.raises(NotImplementedError):
        RenderedDataset(np.zeros((1, 8, 3)), K, CFG_DATASET, preset((64, 64), crop_model="nope"), device=DEV)
    ds = RenderedDataset(np.zeros((1, 8, 3)), K, CFG_DATASET, preset((64, 64)), raw_size=(64, 64), device=DEV)
    with pytest.raises(lib.AbError):
        ds({"rgba": torch.zeros((1, 32, 32, 4), dtype=torch.uint8, device=DEV)})

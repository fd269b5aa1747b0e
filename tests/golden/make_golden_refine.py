"""Generates tests/golden/{refiner,scrambler23,preprocessor_staged}.npz by RUNNING THE REFERENCE'S OWN
anakin/artiboost/{refiner,scrambler,preprocessor}.py (from /root/reference) on the seeded inputs of refine_fixture.py.
Run in the build container:  python tests/golden/make_golden_refine.py

Third-party pieces that are absent here are replaced by the oracle's restatements (so the pins cover the reference's
own composition: iteration structure, feature layout, BatchNorm / residual wiring, indexing, rigid maps):
  chamfer_distance.ChamferDistance -> brute-force squared distances + first-minimum argmin (oracle/refine.py chamfer_nn)
  manotorch.axislayer.AxisLayer    -> oracle/refine.py axis_layer
  manotorch.manolayer / pytorch3d  -> ref_shim.py (oracle/mano_lbs.py, oracle/rotations.py)
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import refine_fixture as fx  # noqa: E402

ref_shim.install()
from artiboost_b200 import assets  # noqa: E402
from oracle import ccv, refine as orf  # noqa: E402


class ChamferDistance(torch.nn.Module):
    def forward(self, x, y):
        d1, i1, d2, i2 = [], [], [], []
        for a, b in zip(x.detach().numpy(), y.detach().numpy()):
            d, i = orf.chamfer_nn(a, b)
            d1.append(d * d), i1.append(i)
            d, i = (np.zeros(len(b), np.float32), np.zeros(len(b), np.int64))  # y -> x is discarded by the reference (:59-66)
            d2.append(d), i2.append(i)
        t = lambda v, dt: torch.from_numpy(np.stack(v).astype(dt))  # noqa: E731
        return t(d1, np.float32), t(d2, np.float32), t(i1, np.int32), t(i2, np.int32)


class AxisLayer(torch.nn.Module):
    def forward(self, joints, transf):
        return tuple(torch.from_numpy(a) for a in orf.axis_layer(joints.numpy(), transf.numpy()))


sys.modules["chamfer_distance"].ChamferDistance = ChamferDistance
sys.modules["manotorch.axislayer"].AxisLayer = AxisLayer


class _FixedDist:
    """Replays recorded Normal draws in call order."""

    def __init__(self, seq):
        self.seq = list(seq)

    def sample(self, shape):
        t = self.seq.pop(0)
        assert tuple(shape) == tuple(t.shape), (shape, t.shape)
        return torch.from_numpy(t).clone()


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def build_ref_refiner(model, iters=3):
    from anakin.artiboost.refiner import HORefiner
    ref_shim.MANO_MODEL["model"] = model
    state = {k: torch.from_numpy(v) for k, v in fx.refinenet_state().items()}
    path = os.path.join(tempfile.mkdtemp(), "refinenet.pt")
    torch.save(state, path)
    ref = HORefiner({"PRETRAINED": path, "ITERS": iters})
    np.random.seed(5)
    ref.setup(fx.object_meshes())
    return ref


def gen_refiner(model):
    ref = build_ref_refiner(model)
    pose, tsl, R, obj_id = fx.refiner_inputs()
    names = [fx.OBJ_NAMES[i] for i in obj_id]
    with torch.no_grad():
        out = ref({"hand_pose": torch.from_numpy(pose), "hand_tsl": torch.from_numpy(tsl), "obj_rot": torch.from_numpy(R)}, names)
        # first-iteration distances, as HORefiner.forward computes them (refiner.py:187-194)
        from anakin.artiboost.refiner import point2point_signed
        hv = ref.refine_net.mano_layer(torch.from_numpy(pose)).verts + torch.from_numpy(tsl).unsqueeze(1)
        vo = torch.transpose(torch.bmm(torch.from_numpy(R), torch.transpose(ref.resampled_objs_buffer[list(obj_id)], -2, -1)), -2, -1)
        h2o = point2point_signed(hv, vo).abs()
    pts = ref.resampled_objs_buffer.numpy()
    save("refiner.npz", pose=pose, tsl=tsl, obj_rot=R, obj_id=obj_id, h2o=h2o.numpy(),
         hand_verts=out["hand_verts"].numpy(), joints=out["joints"].numpy(), hand_pose=out["hand_pose"].numpy(),
         hand_tsl=out["hand_tsl"].numpy(), pts_head=pts[:, :16].copy(), pts_sum=pts.astype(np.float64).sum(axis=(1, 2)))


def gen_scramblers(model):
    from anakin.artiboost.scrambler import RandomScrambler2, RandomScrambler3
    from oracle.mano_lbs import ManoLayer
    rng = np.random.RandomState(17)
    nz = fx.scrambler_noise()
    B = nz["tsl"].shape[0]
    pose = rng.normal(0, 0.3, (B, 48)).astype(np.float32)
    shape = rng.normal(0, 1, (B, 10)).astype(np.float32)
    tsl = rng.normal(0, 0.05, (B, 3)).astype(np.float32)
    out = ManoLayer(model, dtype=np.float32)(pose, shape)
    feed = lambda: {"hand_pose": torch.from_numpy(pose.copy()), "hand_tsl": torch.from_numpy(tsl.copy()),  # noqa: E731
                    "hand_verts": torch.from_numpy(out.verts), "joints": torch.from_numpy(out.joints),
                    "hand_transf": torch.from_numpy(out.transforms_abs)}
    cfg = {"HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1}
    s2 = RandomScrambler2(cfg)
    s2.hand_tsl_dist = _FixedDist([nz["tsl"]])
    s2.hand_pose_dist = _FixedDist([nz["splay"], nz["bend5"], nz["thumb"]])
    r2 = s2(feed())
    s3 = RandomScrambler3(cfg)
    s3.hand_tsl_dist = _FixedDist([nz["tsl"]])
    s3.hand_pose_dist = _FixedDist([nz["splay"], nz["bend14"], nz["thumb"]])
    r3 = s3(feed())
    save("scrambler23.npz", pose=pose, shape=shape, tsl=tsl, joints=out.joints, transf=out.transforms_abs,
         pose2=r2["hand_pose"].numpy(), tsl2=r2["hand_tsl"].numpy(), pose3=r3["hand_pose"].numpy(), tsl3=r3["hand_tsl"].numpy())


def gen_preprocessor_staged(model):
    """PreProcessorPoseGenerator with RandomScrambler2 + HORefiner (2 iterations keeps the fixture run short)."""
    from anakin.artiboost.preprocessor import PreProcessorPoseGenerator
    from anakin.artiboost.scrambler import RandomScrambler2
    from manotorch.manolayer import ManoLayer
    ref = build_ref_refiner(model, iters=2)
    rng = np.random.RandomState(19)
    nz = fx.scrambler_noise(seed=23, B=4)
    B = 4
    pose = rng.normal(0, 0.3, (B, 48)).astype(np.float32)
    shape = rng.normal(0, 1, (B, 10)).astype(np.float32)
    tsl = (rng.normal(0, 0.02, (B, 3)) + [0.0, 0.0, 0.12]).astype(np.float32)
    obj_id = rng.randint(len(fx.OBJ_NAMES), size=B)
    persp, free, zoff = [], [], []
    for i in range(B):
        r, f, z = ccv.view_from_id(int(rng.randint(288)), 12, 24, (0.45, 0.55), *rng.rand(4))
        persp.append(r), free.append(f), zoff.append(z)
    persp, free, zoff = (np.stack(x) for x in (persp, free, zoff))
    scr = RandomScrambler2({"HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1})
    scr.hand_tsl_dist = _FixedDist([nz["tsl"]])
    scr.hand_pose_dist = _FixedDist([nz["splay"], nz["bend5"], nz["thumb"]])
    gen = PreProcessorPoseGenerator(ref, scr, ManoLayer(), ref.refine_net.mano_layer)
    t = torch.from_numpy
    feed = {"index": torch.arange(B), "obj_id": t(obj_id), "obj_name": [fx.OBJ_NAMES[i] for i in obj_id],
            "persp_id": torch.zeros(B), "grasp_id": torch.zeros(B), "hand_pose": t(pose.copy()), "hand_shape": t(shape.copy()),
            "hand_tsl": t(tsl.copy()), "persp_rotmat": t(persp.copy()), "camera_free_transf": t(free.copy()),
            "z_offset": t(zoff.copy())}
    with torch.no_grad():
        out = gen(feed)
    save("preprocessor_staged.npz", pose=pose, shape=shape, tsl=tsl, obj_id=obj_id, persp=persp, free=free, zoff=zoff,
         obj_pose=out["final_obj_pose"].numpy(), verts=out["final_hand_verts"].numpy(), joints=out["final_joints"].numpy())


if __name__ == "__main__":
    model = assets.make_synthetic_mano(seed=0)
    gen_refiner(model)
    gen_scramblers(model)
    gen_preprocessor_staged(model)

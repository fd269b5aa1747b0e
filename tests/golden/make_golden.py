"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (from /root/reference) on seeded inputs.
Run in the build container:  python tests/golden/make_golden.py
The fixtures pin oracle/ (tests/test_oracle_golden.py); the GPU parity tests then compare CUDA against oracle/.

  mano_iknet.npz     anakin/postprocess/iknet/manolayer.py  ManoLayer.__call__  (jax.numpy -> numpy), the in-tree MANO LBS
  view_engine.npz    anakin/artiboost/view_engine.py        ViewEngine.get_view (torch/np RNG replayed into explicit draws)
  ovg_set.npz        anakin/artiboost/ovg_set.py            OVGSet.row_col_calc / compute_occurence_count_map
  scrambler.npz      anakin/artiboost/scrambler.py          RandomScrambler.forward (Normal draws recorded)
  preprocessor.npz   anakin/artiboost/preprocessor.py       PreProcessorPoseGenerator.forward + refiner.NullRefine
                     (third-party MANO / pytorch3d underneath are oracle shims: pins the composition)
  ortho6d.npz        anakin/utils/transform.py              compute_rotation_matrix_from_ortho6d, batch_uvd2xyz
  update_method.npz  anakin/artiboost/artiboost_loader.py   update_method_1 arithmetic (:503-523)
  update_method234.npz  same file, update_method_2 / _3 / _4 (:526-598)
  losses.npz         anakin/criterions/{criterion,jointloss,ordinal}.py  Criterion.compute_losses with pinned random draws
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from artiboost_b200 import assets  # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def gen_mano(model):
    import importlib.util
    spec = importlib.util.spec_from_file_location(  # by path: the package __init__ pulls in jax.experimental
        "ref_iknet_manolayer", os.path.join(ref_shim.REF_ROOT, "anakin/postprocess/iknet/manolayer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    RefMano = mod.ManoLayer
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "models"))
    assets.dump_mano_pkl(model, os.path.join(tmp, "models", "MANO_RIGHT.pkl"))
    rng = np.random.RandomState(1)
    B = 6
    pose = rng.normal(0, 0.35, size=(B, 48))
    pose[0] = 0.0                      # identity pose: verts == shaped template
    pose[1, 3:] = 0.0                  # root rotation only
    betas = rng.normal(0, 1, size=(B, 10))
    betas[0] = 0.0
    out = {}
    for cidx in (0, 9):
        layer = RefMano(center_idx=cidx, flat_hand_mean=True, side="right", mano_root=tmp, use_pca=False)
        verts, jtr, full_pose = layer(pose, betas)
        out[f"verts_c{cidx}"] = np.asarray(verts)
        out[f"joints_c{cidx}"] = np.asarray(jtr)
    save("mano_iknet.npz", pose=pose, betas=betas, **out)


def gen_view():
    from anakin.artiboost.view_engine import ViewEngine
    cfg = {"PERSP_U_BINS": 12, "PERSP_THETA_BINS": 24, "CAMERA_Z_RANGE": [0.45, 0.55]}
    ve = ViewEngine(cfg)
    ids = [0, 1, 23, 24, 100, 143, 144, 200, 287, 57, 11 * 24 + 5, 6 * 24]
    rec = {k: [] for k in ("persp_id", "r_u", "r_theta", "r_roll", "r_z", "rotmat", "free", "z_offset")}
    for i, pid in enumerate(ids):
        seed = 100 + i
        torch.manual_seed(seed)
        np.random.seed(seed)
        rot, free, zoff = ve.get_view(torch.tensor(pid))
        torch.manual_seed(seed)
        np.random.seed(seed)
        r_u, r_theta = float(torch.rand(1)), float(torch.rand(1))
        r_roll = np.random.rand()
        r_z = float(torch.rand(()))
        assert abs((0.45 + r_z * (0.55 - 0.45)) - zoff[2]) < 1e-6, "RNG replay of Uniform.sample diverged"
        for k, v in zip(rec.keys(), (pid, r_u, r_theta, r_roll, r_z, rot, free, zoff)):
            rec[k].append(v)
    # the two poles of caculate_align_mat
    pole = np.stack([ViewEngine.caculate_align_mat(np.array([0.0, 0.0, 1.0])),
                     ViewEngine.caculate_align_mat(np.array([0.0, 0.0, -1.0]))])
    save("view_engine.npz", u_bins=12, theta_bins=24, z_range=np.array([0.45, 0.55]), pole=pole,
         **{k: np.asarray(v) for k, v in rec.items()})


def gen_ovg():
    from anakin.artiboost.ovg_set import OVGSet
    rng = np.random.RandomState(2)
    n_obj, n_persp, n_grasp = 4, 288, 50
    tidx = torch.from_numpy(rng.randint(0, n_obj * n_persp * n_grasp, size=4000))
    b, r, c = OVGSet.row_col_calc(tidx, n_persp, n_grasp)
    occ = OVGSet.compute_occurence_count_map(b, r, c, n_obj, n_persp, n_grasp)
    save("ovg_set.npz", shape=np.array([n_obj, n_persp, n_grasp]), tidx=tidx.numpy(), obj=b.numpy(), persp=r.numpy(),
         grasp=c.numpy(), occ_nonzero=np.argwhere(occ.numpy() > 0), occ_counts=occ.numpy()[occ.numpy() > 0])


class _FixedDist:
    def __init__(self, t):
        self.t = t

    def sample(self, shape):
        assert tuple(shape) == tuple(self.t.shape)
        return self.t.clone()


def gen_scrambler_and_preprocessor(model):
    from anakin.artiboost.scrambler import RandomScrambler
    from anakin.artiboost.refiner import NullRefine
    from anakin.artiboost.preprocessor import PreProcessorPoseGenerator
    from manotorch.manolayer import ManoLayer
    from oracle import ccv

    ref_shim.MANO_MODEL["model"] = model
    rng = np.random.RandomState(3)
    B = 8
    cfg = {"HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1}
    n_tsl = torch.from_numpy(rng.normal(0, 0.01, size=(B, 3)).astype(np.float32))
    n_ang = torch.from_numpy(rng.normal(0, 0.1, size=(B, 16)).astype(np.float32))
    scr = RandomScrambler(cfg)
    scr.hand_tsl_dist, scr.hand_pose_dist = _FixedDist(n_tsl), _FixedDist(n_ang)
    pose = torch.from_numpy(rng.normal(0, 0.3, size=(B, 48)).astype(np.float32))
    pose[0, 6:9] = 0.0  # a zero joint rotation: exercises the 1e-7 clip
    tsl = torch.from_numpy(rng.normal(0, 0.05, size=(B, 3)).astype(np.float32))
    res = scr({"hand_pose": pose, "hand_tsl": tsl})
    save("scrambler.npz", pose=pose.numpy(), tsl=tsl.numpy(), n_tsl=n_tsl.numpy(), n_ang=n_ang.numpy(),
         out_pose=res["hand_pose"].numpy(), out_tsl=res["hand_tsl"].numpy())

    # ---- pose generator: inputs drawn like OVGSet.__getitem__ would deliver them
    shape = torch.from_numpy(rng.normal(0, 1, size=(B, 10)).astype(np.float32))
    shape[1] = 0.0
    persp, free, zoff = [], [], []
    for i in range(B):
        r, f, z = ccv.view_from_id(int(rng.randint(288)), 12, 24, (0.45, 0.55), *rng.rand(4))
        persp.append(r), free.append(f), zoff.append(z)
    persp, free, zoff = (torch.from_numpy(np.stack(x)) for x in (persp, free, zoff))
    for tag, use_noise in (("clean", False), ("noise", True)):
        refiner = NullRefine(cfg=None)
        scr = RandomScrambler(cfg)
        if use_noise:
            scr.hand_tsl_dist, scr.hand_pose_dist = _FixedDist(n_tsl), _FixedDist(n_ang)
        else:
            scr.hand_tsl_dist, scr.hand_pose_dist = _FixedDist(torch.zeros(B, 3)), _FixedDist(torch.zeros(B, 16))
        gen = PreProcessorPoseGenerator(refiner, scr, ManoLayer(), refiner.refine_net.mano_layer)
        feed = {"index": torch.arange(B), "obj_id": torch.zeros(B), "obj_name": ["x"] * B, "persp_id": torch.zeros(B),
                "grasp_id": torch.zeros(B), "hand_pose": pose.clone(), "hand_shape": shape.clone(),
                "hand_tsl": tsl.clone(), "persp_rotmat": persp.clone(), "camera_free_transf": free.clone(),
                "z_offset": zoff.clone()}
        with torch.no_grad():
            out = gen(feed)
        if use_noise:
            save("preprocessor.npz", pose=pose.numpy(), shape=shape.numpy(), tsl=tsl.numpy(), persp=persp.numpy(),
                 free=free.numpy(), zoff=zoff.numpy(), n_tsl=n_tsl.numpy(), n_ang=n_ang.numpy(),
                 obj_pose=out["final_obj_pose"].numpy(), verts=out["final_hand_verts"].numpy(),
                 joints=out["final_joints"].numpy(), **{"clean_" + k: v for k, v in clean.items()})
        else:
            clean = {"obj_pose": out["final_obj_pose"].numpy(), "verts": out["final_hand_verts"].numpy(),
                     "joints": out["final_joints"].numpy()}


def gen_transform():
    from anakin.utils.transform import batch_uvd2xyz, compute_rotation_matrix_from_ortho6d
    rng = np.random.RandomState(4)
    p6 = torch.from_numpy(rng.normal(size=(16, 6)).astype(np.float32))
    R = compute_rotation_matrix_from_ortho6d(p6)
    B = 5
    uvd = torch.from_numpy(rng.rand(B, 22, 3).astype(np.float32))
    root = torch.from_numpy((np.array([0, 0, 0.5]) + rng.normal(0, 0.05, size=(B, 3))).astype(np.float32))
    K = torch.tensor([[217.5, 0, 128], [0, 217.5, 128], [0, 0, 1]], dtype=torch.float32)[None].repeat(B, 1, 1)
    xyz = batch_uvd2xyz(uvd=uvd, root_joint=root, intr=K, inp_res=(256, 256))
    save("ortho6d.npz", p6=p6.numpy(), R=R.numpy(), uvd=uvd.numpy(), root=root.numpy(), K=K.numpy(), xyz=xyz.numpy())


def gen_update_method():
    import logging
    sys.modules["anakin.utils.logger"] = type(sys)("anakin.utils.logger")
    sys.modules["anakin.utils.logger"].logger = logging.getLogger("ref")
    from anakin.artiboost import artiboost_loader as al
    rng = np.random.RandomState(5)
    w = torch.from_numpy(rng.uniform(0.1, 10, size=(4, 288, 50)).astype(np.float32))
    w[0, 0, :5] = 0.0  # blacklisted cells get lifted to 0.1 by the clamp (reference quirk)
    w0 = w.clone()
    cells = np.stack([rng.randint(4, size=300), rng.randint(288, size=300), rng.randint(50, size=300)], 1)
    cells = np.unique(cells, axis=0)
    vals = rng.uniform(5, 60, size=len(cells))
    res = {tuple(int(x) for x in c): float(v) for c, v in zip(cells, vals)}
    w1 = al.ArtiBoostLoader.update_method_1(w, res, 0.1, 10.0)["sample_weight_map"]
    save("update_method.npz", w0=w0.numpy(), cells=cells, vals=vals, w1=w1.numpy())
    # update_method_2..4 (:526-598); method_4 before and after 75 % of the epochs
    L = al.ArtiBoostLoader
    kw = dict(dist_lower_threshold=8.0, dist_upper_threshold=16.0, n_epochs=100)
    w2 = L.update_method_2(w0.clone(), res, 0.1, 10.0)["sample_weight_map"]
    r3 = L.update_method_3(w0.clone(), res, 0.1, 10.0, **kw)
    r4a = L.update_method_4(w0.clone(), res, 0.1, 10.0, epoch_idx=10, **kw)
    r4b = L.update_method_4(w0.clone(), res, 0.1, 10.0, epoch_idx=80, **kw)
    save("update_method234.npz", w0=w0.numpy(), cells=cells, vals=vals, w2=w2.numpy(), w3=r3["sample_weight_map"].numpy(),
         ratio3=np.float64(r3["dist_lower_ratio"]), w4a=r4a["sample_weight_map"].numpy(), ratio4a=np.float64(r4a["dist_lower_ratio"]),
         w4b=r4b["sample_weight_map"].numpy(), ratio4b=np.float64(r4b["dist_lower_ratio"]))


def gen_losses():
    """anakin/criterions/{jointloss,ordinal,criterion}.py with the random draws pinned: random.shuffle -> identity (the
    first third of the pairs is kept), sample_view_vectors -> a recorded set of view vectors."""
    import random as pyrandom
    from anakin.criterions import ordinal
    from anakin.criterions.criterion import Criterion
    from anakin.criterions.jointloss import JointsLoss
    rng = np.random.RandomState(6)
    B = 5
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32))  # noqa: E731
    preds = {"joints_3d_abs": t(rng.normal(0, 0.05, (B, 21, 3)) + [0, 0, 0.5]), "corners_3d_abs": t(rng.normal(0, 0.08, (B, 8, 3)) + [0, 0, 0.5])}
    jv, cv = np.ones((B, 21), np.float32), np.ones((B, 8), np.float32)
    jv[1, 3:7] = 0
    cv[2, :] = 0
    targs = {"joints_3d": t(rng.normal(0, 0.05, (B, 21, 3))), "corners_3d": t(rng.normal(0, 0.08, (B, 8, 3))),
             "root_joint": t(rng.normal(0, 0.05, (B, 3)) + [0, 0, 0.5]), "joints_vis": t(jv), "corners_vis": t(cv)}
    vv = {20: ordinal.sample_view_vectors(20), 40: ordinal.sample_view_vectors(40)}
    orig_shuffle, orig_svv = pyrandom.shuffle, ordinal.sample_view_vectors
    pyrandom.shuffle = lambda x: None
    ordinal.random.shuffle = lambda x: None
    ordinal.sample_view_vectors = lambda n=20: vv[n]
    try:
        crit = Criterion({"LAMBDAS": [0.5, 0.2, 0.1]}, [JointsLoss(LAMBDA_JOINTS_3D=1.0, LAMBDA_CORNERS_3D=0.2),
                                                          ordinal.HandOrdLoss(), ordinal.SceneOrdLoss()])
        total, parts = crit.compute_losses(preds, targs)
    finally:
        pyrandom.shuffle, ordinal.sample_view_vectors = orig_shuffle, orig_svv
        ordinal.random.shuffle = orig_shuffle
    save("losses.npz", vv20=vv[20].numpy(), vv40=vv[40].numpy(), total=total.numpy(),
         **{"pred_" + k: v.numpy() for k, v in preds.items()}, **{"targ_" + k: v.numpy() for k, v in targs.items()},
         **{"part_" + k: (v.numpy() if v is not None else np.zeros(())) for k, v in parts.items()})


if __name__ == "__main__":
    model = assets.make_synthetic_mano(seed=0)
    gen_mano(model)
    gen_view()
    gen_ovg()
    gen_scrambler_and_preprocessor(model)
    gen_transform()
    gen_update_method()
    gen_losses()

"""ORACLE (test infrastructure, never imported by the product path).

CPU/NumPy restatement of MANO linear-blend skinning as the reference consumes it.

The arithmetic lives in third-party `manotorch` (git+https://github.com/lixiny/manotorch.git, UNPINNED,
requirements.txt:178; absent here).  The same algorithm is vendored in-tree at
anakin/postprocess/iknet/manolayer.py:182-276 (manopth lineage) and this file follows that:
  v_shaped  = template + shapedirs . betas                       (:199)
  J         = J_regressor . v_shaped                             (:200)
  v_posed   = v_shaped + posedirs . vec(R_1..15 - I)             (:202)
  chain     = root [R0|J0]; child = parent . [R_k | J_k - J_parent]   (:207-245)
  A_k       = G_k - [0 | G_k . (J_k,0)]                          (:252-255)
  verts     = (sum_k w_vk A_k) . [v_posed;1]                     (:257-262)
  joints    = 16 chain origins + tips [745,317,444,556,673], reordered (:263-270), minus centre joint (:272-274)
Pinned by tests/golden/mano_iknet_*.npz, produced by running that in-tree file itself (jax.numpy -> numpy
shim) on the synthetic MANO pickle: see tests/golden/make_golden.py.

API surface mirrors what the hot path reads from manotorch's MANOOutput (grasp_engine.py:137-145):
verts, joints, center_idx, center_joint, full_poses, betas, transforms_abs.  `transforms_abs` ordering and
`get_rotation_center` are [recalled] from manotorch (not verifiable here): transforms are returned in the
MANO chain order 0..15 (wrist, index, middle, little, ring, thumb) -- the order its consumers in the reference
require: manotorch's AxisLayer pairs transforms_abs[:, 1:] with the chain-ordered keypoint list
[5,6,7, 9,10,11, 17,18,19, 13,14,15, 1,2,3] and anakin/artiboost/scrambler.py:124-181 indexes the resulting axes with
chain-order pose indices; rotation centre = root joint of the shaped template.
"""
from collections import namedtuple

import numpy as np

PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
TIP_VERTS = [745, 317, 444, 556, 673]
JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
TRANSF_REORDER = list(range(16))

MANOOutput = namedtuple("MANOOutput", ["verts", "joints", "center_idx", "center_joint", "full_poses", "betas",
                                       "transforms_abs"])


def rodrigues(aa):
    """Exact-math axis-angle -> R, evaluated in float64 then cast.  The in-tree layer goes through a quaternion and
    adds 1e-8 before the norm (iknet/manolayer.py:161-172); that perturbs R by <=2e-8, far inside the 1e-4 budget."""
    aa = np.asarray(aa)
    a = aa.astype(np.float64)
    th = np.linalg.norm(a, axis=-1, keepdims=True)
    small = th < 1e-12
    k = a / np.where(small, 1.0, th)
    kx, ky, kz = k[..., 0], k[..., 1], k[..., 2]
    z = np.zeros_like(kx)
    K = np.stack([z, -kz, ky, kz, z, -kx, -ky, kx, z], axis=-1).reshape(a.shape[:-1] + (3, 3))
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    R = np.eye(3) + s * K + (1 - c) * (K @ K)
    return R.astype(aa.dtype)


class ManoLayer:
    """manotorch.ManoLayer(rot_mode="axisang", use_pca=False, flat_hand_mean=True) semantics on NumPy."""

    def __init__(self, model, center_idx=None, dtype=np.float32):
        self.dtype = dtype
        self.center_idx = center_idx
        self.v_template = np.asarray(model["v_template"], dtype)          # [778,3]
        self.shapedirs = np.asarray(model["shapedirs"], dtype)            # [778,3,10]
        self.posedirs = np.asarray(model["posedirs"], dtype)              # [778,3,135]
        self.J_regressor = np.asarray(model["J_regressor"], dtype)        # [16,778]
        self.weights = np.asarray(model["weights"], dtype)                # [778,16]
        self.faces = np.asarray(model["f"], np.int64)
        self.th_faces = self.faces

    def get_rotation_center(self, betas=None):
        B = 1 if betas is None else betas.shape[0]
        betas = np.zeros((B, 10), self.dtype) if betas is None else np.asarray(betas, self.dtype)
        v_shaped = self.v_template[None] + np.einsum("vdk,bk->bvd", self.shapedirs, betas)
        return np.einsum("jv,bvd->bjd", self.J_regressor, v_shaped)[:, 0]

    def __call__(self, pose_coeffs, betas=None):
        pose = np.asarray(pose_coeffs, self.dtype)
        B = pose.shape[0]
        betas = np.zeros((B, 10), self.dtype) if betas is None else np.asarray(betas, self.dtype)
        R = rodrigues(pose.reshape(B, 16, 3))                                              # [B,16,3,3]
        pose_map = (R[:, 1:] - np.eye(3, dtype=self.dtype)).reshape(B, 135)
        v_shaped = self.v_template[None] + np.einsum("vdk,bk->bvd", self.shapedirs, betas)
        J = np.einsum("jv,bvd->bjd", self.J_regressor, v_shaped)                           # [B,16,3]
        v_posed = v_shaped + np.einsum("vdk,bk->bvd", self.posedirs, pose_map)
        G = np.zeros((B, 16, 4, 4), self.dtype)
        for k in range(16):
            local = np.zeros((B, 4, 4), self.dtype)
            local[:, :3, :3] = R[:, k]
            local[:, 3, 3] = 1
            if PARENTS[k] < 0:
                local[:, :3, 3] = J[:, k]
                G[:, k] = local
            else:
                local[:, :3, 3] = J[:, k] - J[:, PARENTS[k]]
                G[:, k] = G[:, PARENTS[k]] @ local
        A = G.copy()
        A[:, :, :3, 3] -= np.einsum("bkij,bkj->bki", G[:, :, :3, :3], J)
        T = np.einsum("vk,bkij->bvij", self.weights, A)                                    # [B,778,4,4]
        verts = np.einsum("bvij,bvj->bvi", T[:, :, :3, :3], v_posed) + T[:, :, :3, 3]
        jtr = np.concatenate([G[:, :, :3, 3], verts[:, TIP_VERTS]], axis=1)[:, JOINT_REORDER]
        if self.center_idx is not None:
            center_joint = jtr[:, self.center_idx][:, None]
        else:
            center_joint = np.zeros((B, 1, 3), self.dtype)
        return MANOOutput(verts=verts - center_joint, joints=jtr - center_joint, center_idx=self.center_idx,
                          center_joint=center_joint, full_poses=pose, betas=betas,
                          transforms_abs=G[:, TRANSF_REORDER])

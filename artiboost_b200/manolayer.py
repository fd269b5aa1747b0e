"""MANO layer with the manotorch API surface the hot path uses, backed by the fused sm_100a LBS kernel.

Mirrors `manotorch.manolayer.ManoLayer` as the reference calls it: constructor keywords at
anakin/artiboost/grasp_engine.py:90-95 and anakin/artiboost/artiboost_loader.py:161-170, `forward` at
anakin/artiboost/preprocessor.py:25,62 and anakin/artiboost/refiner.py:138, `get_rotation_center` at
preprocessor.py:55, `MANOOutput` fields at grasp_engine.py:137-145.  Inference only (the hot path runs under
`torch.no_grad()`, artiboost_loader.py:369); the kernel is ab_mano_forward (include/artiboost_b200.h).
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import lib

MANOOutput = namedtuple("MANOOutput", ["verts", "joints", "center_idx", "center_joint", "full_poses", "betas",
                                       "transforms_abs"])
MANOOutput.__new__.__defaults__ = (None,) * 7


class ManoLayer(nn.Module):

    def __init__(self, rot_mode: str = "axisang", side: str = "right", center_idx: Optional[int] = None,
                 mano_assets_root: str = "assets/mano_v1_2", use_pca: bool = False, flat_hand_mean: bool = True,
                 ncomps: int = 15, mano_model: Optional[Dict[str, np.ndarray]] = None, **kwargs):
        """`mano_model` (a dict as returned by assets.make_synthetic_mano / assets.load_mano_pkl) overrides loading
        `<mano_assets_root>/models/MANO_<SIDE>.pkl`."""
        super().__init__()
        if rot_mode != "axisang":
            raise NotImplementedError("only rot_mode='axisang' is on the ArtiBoost hot path")
        if use_pca or not flat_hand_mean:
            raise NotImplementedError("the ArtiBoost hot path uses use_pca=False, flat_hand_mean=True")
        if mano_model is None:
            import os

            from .assets import load_mano_pkl
            mano_model = load_mano_pkl(os.path.join(mano_assets_root, "models", f"MANO_{side.upper()}.pkl"))
        self.rot_mode, self.side, self.center_idx, self.ncomps = rot_mode, side, center_idx, ncomps
        f32 = np.float32
        v_template = np.asarray(mano_model["v_template"], np.float64)               # [778,3]
        shapedirs = np.asarray(mano_model["shapedirs"], np.float64)                 # [778,3,10]
        posedirs = np.asarray(mano_model["posedirs"], np.float64)                   # [778,3,135]
        j_reg = np.asarray(mano_model["J_regressor"], np.float64)                   # [16,778]
        weights = np.asarray(mano_model["weights"], np.float64)                     # [778,16]
        assert v_template.shape == (778, 3) and shapedirs.shape == (778, 3, 10) and posedirs.shape == (778, 3, 135)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(f32)))  # noqa: E731
        # manotorch-compatible views of the model
        self.register_buffer("th_v_template", t(v_template)[None])
        self.register_buffer("th_J_regressor", t(j_reg))
        self.register_buffer("th_faces", torch.from_numpy(np.asarray(mano_model["f"], np.int64)))
        # kernel layouts: k-major blend-shape matrices, regressed joints folded through the regressor
        self.register_buffer("k_v_template", t(v_template.reshape(-1)))
        self.register_buffer("k_shapedirs_t", t(shapedirs.reshape(778 * 3, 10).T))
        self.register_buffer("k_posedirs_t", t(posedirs.reshape(778 * 3, 135).T))
        self.register_buffer("k_j_template", t((j_reg @ v_template).reshape(-1)))
        self.register_buffer("k_j_shapedirs", t(np.einsum("jv,vdk->jdk", j_reg, shapedirs).reshape(48, 10)))
        self.register_buffer("k_weights", t(weights))

    # ------------------------------------------------------------------------------------------------ C-ABI
    def model_struct(self) -> lib.ManoModelStruct:
        lib.require_cuda(self.k_v_template, "ManoLayer buffers")
        return lib.ManoModelStruct(self.k_v_template.data_ptr(), self.k_shapedirs_t.data_ptr(),
                                   self.k_posedirs_t.data_ptr(), self.k_j_template.data_ptr(),
                                   self.k_j_shapedirs.data_ptr(), self.k_weights.data_ptr())

    def get_rotation_center(self, betas: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Root joint of the shaped template, [B,3] (preprocessor.py:55)."""
        jt = self.k_j_template[:3]
        if betas is None:
            return jt[None]
        return jt[None] + betas.to(jt.dtype) @ self.k_j_shapedirs[:3].T

    @torch.no_grad()
    def lbs_into(self, pose: torch.Tensor, betas: Optional[torch.Tensor], post_rt: Optional[torch.Tensor],
                 verts: torch.Tensor, joints: torch.Tensor, transf: Optional[torch.Tensor] = None) -> None:
        """One ab_mano_forward launch into caller-owned buffers; post_rt [B,12] = rigid map fused into the store
        (x' = R x + t).  Used by the refiner, whose MANO forwards are each followed by a translation / rigid map."""
        B, dev = pose.shape[0], pose.device
        m = self.model_struct()
        with torch.cuda.device(dev):
            rc = lib.load().ab_mano_forward(C.byref(m), B, lib.ptr(pose), lib.ptr(betas), lib.ptr(post_rt), -1,
                                            lib.ptr(verts), lib.ptr(joints), lib.ptr(transf), lib.stream_ptr(dev))
        lib.check(rc, "ab_mano_forward")

    @torch.no_grad()
    def forward(self, pose_coeffs: torch.Tensor, betas: Optional[torch.Tensor] = None, **kwargs) -> MANOOutput:
        lib.require_cuda(pose_coeffs, "pose_coeffs")
        if pose_coeffs.dim() != 2 or pose_coeffs.shape[1] != 48:
            raise ValueError(f"pose_coeffs must be [B,48] axis-angle, got {tuple(pose_coeffs.shape)}")
        B = pose_coeffs.shape[0]
        dev = pose_coeffs.device
        pose = pose_coeffs.contiguous().float()
        b = None if betas is None else betas.contiguous().float()
        verts = torch.empty((B, 778, 3), device=dev, dtype=torch.float32)
        joints = torch.empty((B, 21, 3), device=dev, dtype=torch.float32)
        transf = torch.empty((B, 16, 4, 4), device=dev, dtype=torch.float32)
        m = self.model_struct()
        with torch.cuda.device(dev):
            rc = lib.load().ab_mano_forward(C.byref(m), B, lib.ptr(pose), lib.ptr(b), None,
                                            -1, lib.ptr(verts), lib.ptr(joints), lib.ptr(transf),
                                            lib.stream_ptr(dev))
        lib.check(rc, "ab_mano_forward")
        if self.center_idx is not None:
            center_joint = joints[:, self.center_idx:self.center_idx + 1].clone()
            verts -= center_joint
            joints -= center_joint
        else:
            center_joint = torch.zeros((B, 1, 3), device=dev, dtype=torch.float32)
        return MANOOutput(verts=verts, joints=joints, center_idx=self.center_idx, center_joint=center_joint,
                          full_poses=pose, betas=b if b is not None else torch.zeros((B, 10), device=dev),
                          transforms_abs=transf)

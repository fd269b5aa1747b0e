"""Pose generator (anakin/artiboost/preprocessor.py:11-99) as one C-ABI call.

The reference runs three MANO forwards, ~40 small torch kernels and two rotation round trips per batch; here
`forward` is ab_pose_generate: a per-sample prelude + ONE fused LBS launch whose store applies the camera
transform.  The fused call covers the `random` / `naive` / `null` scramblers with the `null` refiner.  The anatomical
scramblers (`random_2` / `random_3`) and the `hand_obj` refiner (HORefiner, the shipped config's default) take the
staged path: ab_pose_prelude -> [MANO forward + ab_scramble_anatomical] -> [refinement loop] -> last LBS launch with the
camera transform fused into its store."""
import ctypes as C

import torch
import torch.nn as nn

from .. import lib
from .refiner import HORefiner, NullRefine
from .scrambler import Scrambler


class PreProcessorPoseGenerator(nn.Module):

    def __init__(self, refiner, scrambler, ge_mano_layer, rf_mano_layer):
        super().__init__()
        self.refiner = refiner
        self.scrambler: Scrambler = scrambler
        self.ge_mano_layer = ge_mano_layer
        self.rf_mano_layer = rf_mano_layer
        if not isinstance(refiner, (NullRefine, HORefiner)):
            raise NotImplementedError(f"unknown refiner {type(refiner).__name__}: REFINER.TYPE is 'null' or 'hand_obj'")
        self.generator = None  # optional torch.Generator for the scrambler draws

    @torch.no_grad()
    def forward(self, synth_extend):
        hand_pose = synth_extend["hand_pose"]
        lib.require_cuda(hand_pose, "synth_extend['hand_pose']")
        dev = hand_pose.device
        B = hand_pose.shape[0]
        f = lambda k: synth_extend[k].to(dev).float().contiguous()  # noqa: E731
        hand_pose, hand_shape, hand_tsl = f("hand_pose"), f("hand_shape"), f("hand_tsl")
        persp, free, zoff = f("persp_rotmat"), f("camera_free_transf"), f("z_offset")
        anatomical = getattr(self.scrambler, "needs_hand_transf", False)
        if anatomical or not isinstance(self.refiner, NullRefine):
            return self._forward_staged(synth_extend, hand_pose, hand_shape, hand_tsl, persp, free, zoff, anatomical)
        if synth_extend.get("noise") is not None:   # drawn with the samples by the fused launch (ab_synth_draw)
            n_tsl, n_ang = synth_extend["noise"]
        else:
            n_tsl, n_ang = self.scrambler.sample_noise(B, dev, self.generator) if self.scrambler is not None else (None, None)
        obj_pose = torch.empty((B, 4, 4), device=dev, dtype=torch.float32)
        verts = torch.empty((B, 778, 3), device=dev, dtype=torch.float32)
        joints = torch.empty((B, 21, 3), device=dev, dtype=torch.float32)
        L = lib.load()
        ws = torch.empty(max(int(L.ab_pose_generate_workspace_bytes(B)), 4), dtype=torch.uint8, device=dev)
        m = self.rf_mano_layer.model_struct()
        with torch.cuda.device(dev):
            rc = L.ab_pose_generate(C.byref(m), B, lib.ptr(hand_pose), lib.ptr(hand_shape), lib.ptr(hand_tsl),
                                    lib.ptr(persp), lib.ptr(free), lib.ptr(zoff), lib.ptr(n_tsl), lib.ptr(n_ang),
                                    lib.ptr(obj_pose), lib.ptr(verts), lib.ptr(joints), lib.ptr(ws),
                                    lib.stream_ptr(dev))
        lib.check(rc, "ab_pose_generate")
        return {
            "index": synth_extend.get("index"),
            "obj_id": synth_extend.get("obj_id"),
            "obj_name": synth_extend.get("obj_name"),
            "persp_id": synth_extend.get("persp_id"),
            "grasp_id": synth_extend.get("grasp_id"),
            "final_obj_pose": obj_pose,
            "final_hand_verts": verts,
            "final_joints": joints,
        }

    def _forward_staged(self, synth_extend, hand_pose, hand_shape, hand_tsl, persp, free, zoff, anatomical, noise=None):
        """preprocessor.py:20-99 with an anatomical scrambler and / or the hand_obj refiner.  `noise` overrides the
        scrambler's draws (tests)."""
        dev, B = hand_pose.device, hand_pose.shape[0]
        L = lib.load()
        e = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)  # noqa: E731
        obj_pose, pose, tsl, cso, post = e(B, 4, 4), e(B, 48), e(B, 3), e(B, 3), e(B, 12)
        n_tsl = n_ang = None
        if not anatomical and self.scrambler is not None:
            n_tsl, n_ang = noise if noise is not None else self.scrambler.sample_noise(B, dev, self.generator)
        m = self.rf_mano_layer.model_struct()
        with torch.cuda.device(dev):
            rc = L.ab_pose_prelude(C.byref(m), B, lib.ptr(hand_pose), lib.ptr(hand_shape), lib.ptr(hand_tsl),
                                   lib.ptr(persp), lib.ptr(free), lib.ptr(zoff), lib.ptr(n_tsl), lib.ptr(n_ang),
                                   lib.ptr(obj_pose), lib.ptr(pose), lib.ptr(tsl), lib.ptr(cso), lib.ptr(post),
                                   lib.stream_ptr(dev))
        lib.check(rc, "ab_pose_prelude")
        verts, joints = e(B, 778, 3), e(B, 21, 3)
        if anatomical:
            # MANO forward #2 (preprocessor.py:62-63) for the joints and hand_transf the AxisLayer reads
            transf = e(B, 16, 4, 4)
            self.rf_mano_layer.lbs_into(pose, hand_shape, None, verts, joints, transf)
            a_tsl, splay, bend, thumb = noise if noise is not None else self.scrambler.sample_anatomical_noise(
                B, dev, self.generator)
            res = self.scrambler({"hand_pose": pose, "hand_tsl": tsl, "joints": joints, "hand_verts": verts,
                                  "hand_transf": transf}, noise=(a_tsl, splay, bend, thumb))
            pose, tsl = res["hand_pose"], res["hand_tsl"]
        if isinstance(self.refiner, HORefiner):
            names = synth_extend.get("obj_name")
            if names is not None:
                obj_id = torch.tensor([self.refiner.obj_idx[n] for n in names], dtype=torch.int32, device=dev)
            else:  # batched pipeline: HORefiner.setup saw the object engine's meshes in obj_id order
                obj_id = synth_extend["obj_id"].to(dev).to(torch.int32)
            # obj_rot = obj_pose[:, :3, :3] (preprocessor.py:79), read in place from the 4x4 poses
            _, _, verts, joints = self.refiner.refine(pose, tsl, obj_pose, obj_id, rigid=free, offset=cso)
        else:
            Rf = free[:, :3, :3]
            post[:, :9] = Rf.reshape(B, 9)
            post[:, 9:] = torch.bmm(Rf, (tsl + cso).unsqueeze(-1)).squeeze(-1)
            self.rf_mano_layer.lbs_into(pose.contiguous(), None, post, verts, joints)  # NullRefine: betas=None
        return {
            "index": synth_extend.get("index"),
            "obj_id": synth_extend.get("obj_id"),
            "obj_name": synth_extend.get("obj_name"),
            "persp_id": synth_extend.get("persp_id"),
            "grasp_id": synth_extend.get("grasp_id"),
            "final_obj_pose": obj_pose,
            "final_hand_verts": verts,
            "final_joints": joints,
        }

"""Pose generator (anakin/artiboost/preprocessor.py:11-99) as one C-ABI call.

The reference runs three MANO forwards, ~40 small torch kernels and two rotation round trips per batch; here
`forward` is ab_pose_generate: a per-sample prelude + ONE fused LBS launch whose store applies the camera
transform.  The fused call covers the `random` / `naive` / `null` scramblers and the `null` refiner; any other
refiner raises (its weights are not available, see refiner.py)."""
import ctypes as C

import torch
import torch.nn as nn

from .. import lib
from .refiner import NullRefine
from .scrambler import Scrambler


class PreProcessorPoseGenerator(nn.Module):

    def __init__(self, refiner, scrambler, ge_mano_layer, rf_mano_layer):
        super().__init__()
        self.refiner = refiner
        self.scrambler: Scrambler = scrambler
        self.ge_mano_layer = ge_mano_layer
        self.rf_mano_layer = rf_mano_layer
        if not isinstance(refiner, NullRefine):
            raise NotImplementedError("the fused pose generator supports REFINER.TYPE 'null' only")
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

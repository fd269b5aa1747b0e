"""HybridBaseline ("clasbased", anakin/models/hybridbaseline.py:17-96): backbone -> 3-D heatmap head (21 joints + box
root) + MLP box rotation -> uvd->xyz -> 8 corners.  Emits the same seven outputs."""
import os
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn as nn

from .registry import MODEL, build_backbone, build_head, build_model, enable_lower_param
from .transform import batch_uvd2xyz, compute_rotation_matrix_from_ortho6d


@MODEL.register_module
class HybridBaseline(nn.Module):

    @enable_lower_param
    def __init__(self, **cfg):
        super().__init__()
        self.center_idx = cfg["DATA_PRESET"].get("CENTER_IDX", 9)
        self.inp_res = cfg["DATA_PRESET"]["IMAGE_SIZE"]
        self.backbone = build_backbone(cfg["BACKBONE"], default_args=cfg["DATA_PRESET"])
        self.hybrid_head = build_head(cfg["HYBRID_HEAD"], default_args=cfg["DATA_PRESET"])
        self.box_head = build_model(cfg["BOX_HEAD"], default_args=cfg["DATA_PRESET"])
        self.init_weights(pretrained=cfg["PRETRAINED"])
        # training-step fusion (models/fused_tail.py, installed by train.TrainStep): tail + criterion + their gradient in one launch
        self.fused_tail = None

    def forward(self, inputs: Dict):
        batch_size, n_channel, height, width = inputs["image"].shape
        feats = self.backbone.forward_acts(inputs["image"])
        pose_results = self.hybrid_head.forward_act(feats["res_layer4"])
        box_rot_6d = self.box_head(feats["res_layer4_mean_bf16"])
        if self.fused_tail is not None and torch.is_grad_enabled() and self.fused_tail.usable(inputs):
            preds, loss, parts = self.fused_tail(pose_results["kp3d"], box_rot_6d, inputs)
            preds["_fused_loss"], preds["_fused_parts"] = loss, parts
            return preds
        if (not torch.is_grad_enabled() and pose_results["kp3d"].is_cuda and os.environ.get("AB_FUSED_TAIL", "1") != "0"):
            from .fused_tail import tail_forward   # inference: the tail as ONE launch (the composition below defines it)
            return tail_forward(pose_results["kp3d"], box_rot_6d, inputs, self.center_idx, self.inp_res)
        pose_3d_abs = batch_uvd2xyz(uvd=pose_results["kp3d"], root_joint=inputs["root_joint"], intr=inputs["cam_intr"],
                                    inp_res=self.inp_res)
        joints_3d_abs = pose_3d_abs[:, 0:21, :]
        boxroot_3d_abs = pose_3d_abs[:, 21:22, :]
        corners_can_3d = inputs["corners_can"].to(boxroot_3d_abs.device)
        box_rot_rotmat = compute_rotation_matrix_from_ortho6d(box_rot_6d)
        corners_3d_abs = torch.matmul(box_rot_rotmat, corners_can_3d.permute(0, 2, 1)).permute(0, 2, 1) + boxroot_3d_abs
        root_joint = joints_3d_abs[:, self.center_idx, :]
        cam_intr = inputs["cam_intr"].to(corners_3d_abs.device)
        corners_2d = torch.matmul(cam_intr, corners_3d_abs.permute(0, 2, 1)).permute(0, 2, 1)
        corners_2d = corners_2d[:, :, 0:2] / corners_2d[:, :, 2:3]
        corners_2d[:, :, 0] /= width
        corners_2d[:, :, 1] /= height
        corners_2d_uvd = torch.cat((corners_2d, torch.zeros_like(corners_2d[:, :, 0:1])), dim=2)
        final_2d_uvd = torch.cat((pose_results["kp3d"][:, 0:21, :], corners_2d_uvd, pose_results["kp3d"][:, 21:22, :]), dim=1)
        return {
            "joints_3d_abs": joints_3d_abs,
            "corners_3d_abs": corners_3d_abs,
            "joints_3d": joints_3d_abs - root_joint.unsqueeze(1),
            "corners_3d": corners_3d_abs - root_joint.unsqueeze(1),
            "2d_uvd": final_2d_uvd,
            "boxroot_3d_abs": boxroot_3d_abs,
            "box_rot_rotmat": box_rot_rotmat,
        }

    def init_weights(self, pretrained=""):
        if pretrained == "":
            return
        if not os.path.isfile(pretrained):
            raise FileNotFoundError(f"=> No {type(self).__name__} checkpoints file found in {pretrained}")
        checkpoint = torch.load(pretrained, map_location="cpu")
        if isinstance(checkpoint, OrderedDict):
            state_dict = checkpoint
        elif isinstance(checkpoint, dict) and "state_dict" in checkpoint:
            state_dict = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in checkpoint["state_dict"].items())
        else:
            raise RuntimeError(f"=> No state_dict found in checkpoint file {pretrained}")
        self.load_state_dict(state_dict, strict=False)

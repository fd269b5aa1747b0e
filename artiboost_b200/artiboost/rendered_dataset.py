"""Batched, on-device counterpart of anakin.artiboost.rendered_dataset.RenderedDataset (rendered_dataset.py:20-340).

The reference turns ONE rendered image into a training sample per `__getitem__` call, on the CPU inside a DataLoader
worker (PIL blur / colour jitter / affine warp, numpy annotation transforms).  Here a whole batch of rendered views goes
through `ab_crop_augment` (csrc/augment.cu) in three launches and never leaves the GPU.  Same constructor configuration
(cfg_dataset AUG / AUG_PARAM, cfg_preset IMAGE_SIZE / CENTER_IDX / BBOX_EXPAND_RATIO / FULL_IMAGE / CROP_MODEL), same
sample keys (hoquery.py Queries / SynthQueries), same arithmetic down to the byte (oracle/augment.py)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .. import lib

CROP_MODELS = {"root_obj": 0, "hand_obj": 1, "hand": 2}


class RenderedDataset:

    def __init__(self, obj_corners, cam_intr, cfg_dataset: dict, cfg_preset: dict, crop_image: Optional[dict] = None,
                 raw_size=None, device="cuda", generator: Optional[torch.Generator] = None, obj_class_ids=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise lib.AbError("RenderedDataset runs on a CUDA device only: artiboost_b200 has no CPU path")
        self.obj_corners = torch.as_tensor(np.asarray(obj_corners, np.float32), device=self.device).contiguous()  # [n_obj, 8, 3]
        # Queries.OBJ_IDX is the 1-based YCB class id of the object (rendered_dataset.py:38,236: obj_map[objname])
        ids = np.arange(1, len(self.obj_corners) + 1) if obj_class_ids is None else np.asarray(obj_class_ids)
        self.obj_class_ids = torch.as_tensor(ids, dtype=torch.int32, device=self.device)
        self.cam_intr = np.asarray(cam_intr, np.float32).reshape(3, 3)
        self.image_size = tuple(cfg_preset["IMAGE_SIZE"])  # (W, H)
        self.raw_size = tuple(raw_size) if raw_size is not None else self.image_size
        self.center_idx = int(cfg_preset["CENTER_IDX"])
        self.bbox_expand_ratio = float(cfg_preset["BBOX_EXPAND_RATIO"] if crop_image is None else crop_image["BBOX_EXPAND_RATIO"])
        self.require_full_image = bool(cfg_preset["FULL_IMAGE"] if crop_image is None else crop_image["FULL_IMAGE"])
        self.crop_model = (cfg_preset.get("CROP_MODEL", "hand_obj") if crop_image is None else crop_image["CROP_MODEL"])
        if self.crop_model not in CROP_MODELS:
            raise NotImplementedError(self.crop_model)  # as rendered_dataset.py:303-304
        if self.require_full_image:
            self.bbox_expand_ratio = 1.0
        self.aug = bool(cfg_dataset["AUG"])
        aug_param = cfg_dataset.get("AUG_PARAM")
        if self.aug:  # rendered_dataset.py:61-71
            self.hue, self.saturation, self.contrast, self.brightness, self.blur_radius = 0.075, 0.1, 0.1, 0.1, 0.1
            self.scale_jittering = float(aug_param["SCALE_JIT"]) if aug_param is not None else 0.0
            self.center_jittering = float(aug_param["CENTER_JIT"]) if aug_param is not None else 0.0
            self.max_rot = float(aug_param["MAX_ROT"]) * math.pi if aug_param is not None else 0.0
        else:
            self.hue = self.saturation = self.contrast = self.brightness = self.blur_radius = 0.0
            self.scale_jittering = self.center_jittering = self.max_rot = 0.0
        self.generator = generator
        self._cfg = lib.AugmentCfgStruct(self.raw_size[0], self.raw_size[1], self.image_size[0], self.image_size[1], self.center_idx,
                                         CROP_MODELS[self.crop_model], int(self.require_full_image), int(self.aug),
                                         self.bbox_expand_ratio, self.center_jittering, self.scale_jittering,
                                         (lib.C.c_float * 9)(*self.cam_intr.reshape(-1).tolist()))

    # ------------------------------------------------------------------------------------------------ random draws
    def draw(self, B: int):
        """The per-sample draws of rendered_dataset.py:178-189,256 and img_augment.py:6-45, taken on the device.
        -> (draws f32 [B, 10], order i32 [B, 4])"""
        dev, g = self.device, self.generator
        u = torch.rand((B, 9), device=dev, generator=g)
        n = torch.randn(B, device=dev, generator=g)
        rot = (u[:, 2] * 2.0 - 1.0) * self.max_rot

        def around_one(x, col):  # random.uniform(max(0, 1 - x), 1 + x)
            lo = max(0.0, 1.0 - x)
            return lo + u[:, col] * (1.0 + x - lo)

        draws = torch.stack([u[:, 0] * 2.0 - 1.0, u[:, 1] * 2.0 - 1.0, n * (self.scale_jittering / 3.0), torch.cos(rot), torch.sin(rot),
                             u[:, 3] * self.blur_radius, around_one(self.brightness, 4), around_one(self.contrast, 5),
                             around_one(self.saturation, 6), (u[:, 7] * 2.0 - 1.0) * self.hue], dim=1).float().contiguous()
        order = torch.rand((B, 4), device=dev, generator=g).argsort(dim=1).to(torch.int32).contiguous()  # random.shuffle
        return draws, order

    # ------------------------------------------------------------------------------------------------ the batch
    @torch.no_grad()
    def __call__(self, views: Dict[str, torch.Tensor], draws: Optional[torch.Tensor] = None,
                 order: Optional[torch.Tensor] = None, return_affine: bool = False) -> Dict[str, torch.Tensor]:
        """views: rgba u8 [B,H,W,4], joints f32 [B,21,3], obj_pose f32 [B,4,4], obj_id (+ persp_id, grasp_id)."""
        rgba = views["rgba"]
        lib.require_cuda(rgba, "rgba")
        B, H, W, _ = rgba.shape
        if (W, H) != self.raw_size:
            raise lib.AbError(f"rendered views are {W}x{H} but raw_size is {self.raw_size}")
        dev = rgba.device
        if self.aug and draws is None:
            draws, order = self.draw(B)
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        Wo, Ho = self.image_size
        corners_can = self.obj_corners[views["obj_id"].long()].contiguous()
        out = {"image": f(B, 3, Ho, Wo), "cam_intr": f(B, 3, 3), "root_joint": f(B, 3), "joints_3d": f(B, 21, 3), "joints_2d": f(B, 21, 2),
               "joints_vis": f(B, 21), "corners_3d": f(B, 8, 3), "corners_2d": f(B, 8, 2), "corners_vis": f(B, 8), "obj_transf": f(B, 4, 4)}
        aff = (f(B, 6), f(B, 6)) if return_affine else (None, None)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        L = lib.load()
        ws = torch.empty(int(L.ab_augment_workspace_bytes(self._cfg, B)) + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        # converted copies are bound to names: a temporary would be freed (and its block handed to the next temporary)
        # as soon as .data_ptr() returns, before the launch
        rgba_c = rgba.contiguous()
        joints_c = views["joints"].float().contiguous()
        pose_c = views["obj_pose"].float().contiguous()
        with torch.cuda.device(dev):
            rc = L.ab_crop_augment(self._cfg, B, lib.ptr(rgba_c), lib.ptr(joints_c), lib.ptr(pose_c), lib.ptr(corners_can), P(draws), P(order),
                                   *[out[k].data_ptr() for k in ("image", "cam_intr", "root_joint", "joints_3d", "joints_2d", "joints_vis",
                                                                 "corners_3d", "corners_2d", "corners_vis", "obj_transf")],
                                   P(aff[0]), P(aff[1]), status.data_ptr(), ws_ptr, lib.stream_ptr(dev))
        lib.check(rc, "ab_crop_augment")
        out["corners_can"] = corners_can
        out["obj_idx"] = self.obj_class_ids[views["obj_id"].long()]
        out["is_synth"] = torch.ones(B, device=dev)
        # Queries.SAMPLE_IDX (rendered_dataset.py:272, hoquery.py:7): the index of the sample in the epoch's synthetic set;
        # views drawn on the fly carry it as "sample_idx", else they are numbered from `index_base`
        out["sample_idx"] = (views["sample_idx"].to(dev, torch.int64) if "sample_idx" in views
                             else torch.arange(B, device=dev, dtype=torch.int64) + int(views.get("index_base", 0)))
        for k in ("obj_id", "persp_id", "grasp_id"):
            if k in views:
                out[k] = views[k]
        out["_status"] = status  # device word: bit 0 blur radius out of range, bit 1 warp outside Pillow's fixed-point range
        if return_affine:
            out["affine"], out["inv_affine"] = aff
        return out

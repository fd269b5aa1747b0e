"""The online-synthesis pipeline as one object: CCV sample -> view -> grasp lookup -> pose generator -> rasterise.

This is what ArtiBoostLoader.generate_render_cache + RenderedDataset.prepare_essential do per epoch / per sample in
the reference (anakin/artiboost/artiboost_loader.py:352-387, rendered_dataset.py:103-123), kept entirely on device
and batched.  Assets are the synthetic stand-ins of artiboost_b200/assets.py unless the caller passes real ones.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import assets
from .artiboost import (GraspEngine, HORefiner, NullRefine, ObjEngine, OVGSet, PreProcessorPoseGenerator, Renderer,
                        Scrambler, ViewEngine, make_mesh)
from .artiboost.renderer import PYRENDER_EXTRINSIC, PointLight

# config/ho3dv2_clasbased_jlol_artiboost2.yaml:6-20,42-45 (HO3D CCV space), render camera rescaled to 256^2 (same FoV
# as the shipped 512^2 / f=435 camera, yaml:52-61)
DEFAULT_CFG = {
    "VIEW": {"PERSP_U_BINS": 12, "PERSP_THETA_BINS": 24, "CAMERA_Z_RANGE": [0.45, 0.55]},
    "GRASP_NUM": 50,
    "SCRAMBLER": {"TYPE": "random", "HAND_TSL_SIGMA": 0.01, "HAND_POSE_SIGMA": 0.1},
    # the shipped config's refiner is "hand_obj" (yaml:47-50); its GrabNet weights are a licensed asset, so the synthetic
    # default stays "null".  {"TYPE": "hand_obj", "PRETRAINED": path or None, "ITERS": 3} builds HORefiner (randomly
    # initialised when PRETRAINED is None -- same arithmetic, meaningless refinement).
    "REFINER": {"TYPE": "null"},
    "RENDER_SIZE": [256, 256],
    "CAM_PARAM": {"FX": 217.5, "FY": 217.5, "CX": 128.0, "CY": 128.0},
}


class SynthPipeline:

    def __init__(self, obj_names: Optional[List[str]] = None, device="cuda", seed: int = 0, cfg: Optional[dict] = None,
                 n_hand_tex: int = 51, n_bg: int = 8, chunk: int = 512, mano_model: Optional[Dict] = None,
                 objects: Optional[Dict[str, dict]] = None, grasps: Optional[Dict[str, list]] = None):
        self.cfg = cfg = dict(DEFAULT_CFG if cfg is None else cfg)
        self.device = dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("SynthPipeline runs on a CUDA device only")
        self.obj_names = list(obj_names) if obj_names is not None else list(assets.HO3D_TRAIN_OBJS)
        self.mano_model = mano_model if mano_model is not None else assets.make_synthetic_mano(seed)
        self.objects = objects if objects is not None else assets.make_synthetic_objects(self.obj_names, seed)
        self.grasps = grasps if grasps is not None else assets.make_synthetic_grasps(self.objects, cfg["GRASP_NUM"], seed)
        self.obj_engine = ObjEngine(self.objects, self.obj_names, device=dev)
        self.grasp_engine = GraspEngine(self.grasps, self.obj_names, n_grasp=cfg["GRASP_NUM"], device=dev)
        self.view_engine = ViewEngine(cfg["VIEW"])
        self.generator = torch.Generator(device=dev)
        self.generator.manual_seed(seed)
        shape = (len(self.obj_names), self.view_engine.n_persp_center, self.grasp_engine.n_grasp)
        self.sample_weight_map = torch.ones(shape, device=dev)
        self.occurence_map = torch.zeros(shape, dtype=torch.bool, device=dev)
        self.ovg_set = OVGSet(self.obj_engine, self.grasp_engine, self.view_engine, 0, 0, self.grasp_engine.n_grasp,
                              torch.zeros(shape, dtype=torch.bool), device=dev, generator=self.generator)
        rcfg = cfg.get("REFINER", {"TYPE": "null"})
        if rcfg["TYPE"] == "hand_obj":
            torch.manual_seed(seed)
            self.refiner = HORefiner(rcfg, mano_model=self.mano_model)
            np.random.seed(seed)  # resample_obj draws from np.random (refiner.py:180)
            self.refiner.setup(self.obj_engine.obj_trimeshes_mapping)
            self.refiner = self.refiner.to(dev)
        else:
            self.refiner = NullRefine(mano_model=self.mano_model).to(dev)
        self.scrambler = Scrambler.build(cfg["SCRAMBLER"]["TYPE"], cfg["SCRAMBLER"])
        self.pose_generator = PreProcessorPoseGenerator(self.refiner, self.scrambler, self.refiner.refine_net.mano_layer,
                                                        self.refiner.refine_net.mano_layer).to(dev)
        self.pose_generator.generator = self.generator
        W, H = cfg["RENDER_SIZE"]
        K = cfg["CAM_PARAM"]
        self.cam_intr = np.array([[K["FX"], 0, K["CX"]], [0, K["FY"], K["CY"]], [0, 0, 1]], np.float32)
        hand_faces = self.mano_model["f"]
        tex = assets.make_hand_textures(n_hand_tex, seed, template=self.mano_model["v_template"])
        self.hand_meshes = [make_mesh(self.mano_model["v_template"], hand_faces, tex[i]) for i in range(n_hand_tex)]
        self.backgrounds = list(assets.make_backgrounds(n_bg, int(1.5 * H), int(1.5 * W), seed)) if n_bg else None
        self.renderer = Renderer(W, H, gpu_id=dev.index or 0, chunk=chunk)
        self.renderer.rng = np.random.RandomState(seed + 1)
        self.renderer.setup(self.cam_intr, PYRENDER_EXTRINSIC, self.obj_engine.obj_trimeshes_mapping, self.hand_meshes,
                            self.backgrounds, [PointLight(np.array([0.9, 0.9, 0.9]), 5.0, np.eye(4))])

    @torch.no_grad()
    def sample_poses(self, n: int) -> dict:
        """CCV draw + view + grasp + pose generator for n views -> final_obj_pose / final_hand_verts / final_joints
        plus the (obj, persp, grasp) ids."""
        self.ovg_set.train()
        self.ovg_set.update_len(config_len_train=n)
        _, self.occurence_map = self.ovg_set.update(self.sample_weight_map, self.occurence_map)
        return self.pose_generator(self.ovg_set.get_batch(0, n))

    @torch.no_grad()
    def render(self, poses: dict, rand: Optional[dict] = None, out: Optional[dict] = None) -> dict:
        """rand: optional {"hand_tex", "light", "bg_sel"} device tensors (the per-view draws of renderer.py:102-104);
        drawn on device from the pipeline's generator when omitted."""
        B = poses["final_obj_pose"].shape[0]
        rand = rand or self.draw_render_randoms(B)
        return self.renderer.render_batch(poses["obj_id"], poses["final_obj_pose"], poses["final_hand_verts"],
                                          hand_tex=rand["hand_tex"], light=rand["light"], bg_sel=rand.get("bg_sel"),
                                          out=out)

    def draw_render_randoms(self, B: int) -> dict:
        dev, g, r = self.device, self.generator, self.renderer
        rand = {"hand_tex": torch.randint(r.n_hand_tex, (B,), device=dev, generator=g, dtype=torch.int32),
                "light": 1.0 + 4.0 * torch.rand(B, device=dev, generator=g)}
        if r.backgrounds is not None:
            nb, bh, bw = r.backgrounds.shape[:3]
            u = torch.rand((B, 4), device=dev, generator=g)
            bid = (u[:, 0] * nb).long().clamp_(max=nb - 1)
            # crop rule of renderer.py:125-136 for square-proportional backgrounds: height drawn in [H, bh]
            ch = r.height + (u[:, 1] * (bh - r.height + 1)).long().clamp_(max=bh - r.height)
            cw = (ch * r.width) // r.height
            y0 = (u[:, 2] * (bh - ch + 1).float()).long().clamp_(min=0)
            x0 = (u[:, 3] * (bw - cw + 1).float()).long().clamp_(min=0)
            y0 = torch.minimum(y0, bh - ch)
            x0 = torch.minimum(x0, bw - cw)
            rand["bg_sel"] = torch.stack([bid, x0, y0, cw, ch], 1).to(torch.int32)
        return rand

    @torch.no_grad()
    def synthesise(self, n: int, out: Optional[dict] = None) -> dict:
        poses = self.sample_poses(n)
        views = self.render(poses, out=out)
        views.update(obj_id=poses["obj_id"], persp_id=poses["persp_id"], grasp_id=poses["grasp_id"],
                     obj_pose=poses["final_obj_pose"], hand_verts=poses["final_hand_verts"],
                     joints=poses["final_joints"])
        return views

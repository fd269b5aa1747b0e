"""The online-synthesis pipeline as one object: CCV sample -> view -> grasp lookup -> pose generator -> rasterise.

This is what ArtiBoostLoader.generate_render_cache + RenderedDataset.prepare_essential do per epoch / per sample in
the reference (anakin/artiboost/artiboost_loader.py:352-387, rendered_dataset.py:103-123), kept entirely on device
and batched.  Assets are the synthetic stand-ins of artiboost_b200/assets.py unless the caller passes real ones.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import ctypes as C
import os

import numpy as np
import torch

from . import assets, lib
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
                 objects: Optional[Dict[str, dict]] = None, grasps: Optional[Dict[str, list]] = None,
                 filter_back: bool = True, fused_draw: bool = True, sample_seed: Optional[int] = None):
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
        # `seed` makes the (synthetic) assets; `sample_seed` (default: the same) seeds everything that is DRAWN -- CCV cells,
        # jitter, scrambler noise, the renderer's per-view choices.  Ranks of one job share the assets and differ in the draws.
        draw_seed = int(seed if sample_seed is None else sample_seed)
        self.generator.manual_seed(draw_seed)
        shape = (len(self.obj_names), self.view_engine.n_persp_center, self.grasp_engine.n_grasp)
        self._seed, self._offset = draw_seed, 0       # the Philox stream of the fused draw: (seed, per-call offset)
        self._space = None                            # ab_synth_space; completed once the renderer exists (texture / bg counts)
        self._cdf, self._cdf_key = None, None
        self._occ_map = torch.zeros(shape, dtype=torch.bool, device=dev)
        self._occ_count = torch.zeros(shape, dtype=torch.int32, device=dev)
        self._render_rand = None
        self.sample_weight_map = torch.ones(shape, device=dev)
        # artiboost_loader.py:125-130: back-of-hand cells are blacklisted and never drawn
        self.blacklist_map = (self.construct_blacklist_map() if filter_back
                              else torch.zeros(shape, dtype=torch.bool, device=dev))
        self.sample_weight_map[self.blacklist_map] = 0.0
        self.ovg_set = OVGSet(self.obj_engine, self.grasp_engine, self.view_engine, 0, 0, self.grasp_engine.n_grasp,
                              self.blacklist_map, device=dev, generator=self.generator)
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
        self.renderer.rng = np.random.RandomState(draw_seed + 1)
        self.renderer.setup(self.cam_intr, PYRENDER_EXTRINSIC, self.obj_engine.obj_trimeshes_mapping, self.hand_meshes,
                            self.backgrounds, [PointLight(np.array([0.9, 0.9, 0.9]), 5.0, np.eye(4))])

        sc = cfg["SCRAMBLER"]
        self.fused_draw = bool(fused_draw and rcfg["TYPE"] == "null" and sc["TYPE"] in ("random", "null"))
        self._sigmas = ((float(sc.get("HAND_TSL_SIGMA", 0.0)), float(sc.get("HAND_POSE_SIGMA", 0.0)))
                        if sc["TYPE"] == "random" else (0.0, 0.0))

    # ------------------------------------------------------------------------------------------ CCV space on device
    def _synth_space(self) -> "lib.SynthSpaceStruct":
        v, r = self.view_engine, getattr(self, "renderer", None)
        nb, bh, bw = (r.backgrounds.shape[:3] if r is not None and r.backgrounds is not None else (0, 0, 0))
        zmin, zmax = v.camera_z_range
        W, H = self.cfg["RENDER_SIZE"]
        sig = getattr(self, "_sigmas", (0.0, 0.0))
        return lib.SynthSpaceStruct(len(self.obj_names), v.n_persp_center, self.grasp_engine.n_grasp, v.persp_u_bins,
                                    v.persp_theta_bins, float(zmin), float(zmax), self.grasp_engine.table.data_ptr(),
                                    sig[0], sig[1], r.n_hand_tex if r is not None else 1, 1.0, 5.0, int(nb), int(bh), int(bw),
                                    int(W), int(H))

    @torch.no_grad()
    def construct_blacklist_map(self, rand2: Optional[torch.Tensor] = None, threshold: float = -0.8, return_th: bool = False):
        """ArtiBoostLoader._construct_blacklist_map (artiboost_loader.py:415-500) over the whole CCV space in one launch.
        rand2 f32 [n_obj, n_persp, n_grasp, 2]: the (u, theta) jitter get_view draws per cell; drawn here when omitted, like
        the reference's loop does."""
        dev = self.device
        shape = (len(self.obj_names), self.view_engine.n_persp_center, self.grasp_engine.n_grasp)
        if rand2 is None:
            rand2 = torch.rand(shape + (2,), device=dev, generator=self.generator)
        rand2 = rand2.to(dev).float().contiguous()
        out = torch.empty(shape, dtype=torch.uint8, device=dev)
        th = torch.empty(shape, dtype=torch.float32, device=dev) if return_th else None
        sp = self._synth_space()
        with torch.cuda.device(dev):
            rc = lib.load().ab_ccv_blacklist(C.byref(sp), lib.ptr(rand2), float(threshold), lib.ptr(out), lib.ptr(th),
                                             lib.stream_ptr(dev))
        lib.check(rc, "ab_ccv_blacklist")
        return (out.bool(), th) if return_th else out.bool()

    @property
    def occurence_map(self) -> torch.Tensor:
        """bool [n_obj, n_persp, n_grasp]: cells drawn so far (ovg_set.py:172-178).  The fused draw counts on the device;
        the counts are folded into the map when it is read."""
        if getattr(self, "_pose_stream", None) is not None:   # draws issued ahead by synthesise(prefetch=True) count too
            torch.cuda.current_stream(self.device).wait_stream(self._pose_stream)
        self._occ_map |= self._occ_count > 0
        self._occ_count.zero_()
        return self._occ_map

    @occurence_map.setter
    def occurence_map(self, value: torch.Tensor):
        self._occ_map = value.to(self.device).bool()
        self._occ_count.zero_()

    def _ccv_cdf(self) -> torch.Tensor:
        """The fp64 CDF of the weight map, recomputed only when the map changes (once per epoch: step_eval)."""
        w = self.sample_weight_map
        key = (w.data_ptr(), w._version)
        if self._cdf is None or key != self._cdf_key:
            w = w.detach().to(self.device).float().contiguous()
            if self._cdf is None:
                self._cdf = torch.empty(w.numel(), dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                lib.check(lib.load().ab_ccv_cdf(lib.ptr(w), w.numel(), lib.ptr(self._cdf), lib.stream_ptr(self.device)), "ab_ccv_cdf")
            self._cdf_key = (self.sample_weight_map.data_ptr(), self.sample_weight_map._version)
            self._cdf_src = w   # keeps a converted copy alive until the launch has run
        return self._cdf

    @torch.no_grad()
    def draw(self, n: int, uniforms: Optional[torch.Tensor] = None, return_uniforms: bool = False) -> dict:
        """One launch (ab_synth_draw): CCV cells, views, grasps, scrambler noise and the renderer's per-view draws for n
        samples.  `uniforms` f32 [n, 32] replaces the Philox stream (tests)."""
        dev = self.device
        if self._space is None:
            self._space = self._synth_space()
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)  # noqa: E731
        b = {"obj_id": i32(n), "persp_id": i32(n), "grasp_id": i32(n), "hand_pose": e(n, 48), "hand_shape": e(n, 10),
             "hand_tsl": e(n, 3), "persp_rotmat": e(n, 3, 3), "camera_free_transf": e(n, 4, 4), "z_offset": e(n, 3)}
        scr = self._sigmas != (0.0, 0.0)
        n_tsl, n_ang = (e(n, 3), e(n, 16)) if scr else (None, None)
        rr = {"hand_tex": i32(n), "light": e(n), "bg_sel": i32(n, 5)}
        u_out = e(n, lib.SYNTH_UNIFORMS) if return_uniforms else None
        if uniforms is not None:
            uniforms = uniforms.to(dev).float().contiguous()
        cdf = self._ccv_cdf()
        with torch.cuda.device(dev):
            rc = lib.load().ab_synth_draw(
                C.byref(self._space), lib.ptr(cdf), n, self._seed, self._offset, lib.ptr(uniforms), lib.ptr(b["obj_id"]),
                lib.ptr(b["persp_id"]), lib.ptr(b["grasp_id"]), lib.ptr(self._occ_count), lib.ptr(b["hand_pose"]),
                lib.ptr(b["hand_shape"]), lib.ptr(b["hand_tsl"]), lib.ptr(b["persp_rotmat"]), lib.ptr(b["camera_free_transf"]),
                lib.ptr(b["z_offset"]), lib.ptr(n_tsl), lib.ptr(n_ang), lib.ptr(rr["hand_tex"]), lib.ptr(rr["light"]),
                lib.ptr(rr["bg_sel"]), lib.ptr(u_out), lib.stream_ptr(dev))
        lib.check(rc, "ab_synth_draw")
        self._offset += lib.SYNTH_UNIFORMS // 4
        if self.renderer.backgrounds is None:
            rr.pop("bg_sel")
        b.update(index=None, obj_name=None, noise=(n_tsl, n_ang) if scr else None)
        self._render_rand = (n, rr)
        if return_uniforms:
            b["uniforms"] = u_out
        return b

    @torch.no_grad()
    def sample_poses(self, n: int) -> dict:
        """CCV draw + view + grasp + pose generator for n views -> final_obj_pose / final_hand_verts / final_joints
        plus the (obj, persp, grasp) ids.  With REFINER null and the `random` (or no) scrambler this is three launches: the
        fused draw (ab_synth_draw) and the pose generator's prelude + LBS; other configurations take the staged path."""
        if self.fused_draw:
            return self.pose_generator(self.draw(n))
        self._render_rand = None
        self.ovg_set.train()
        self.ovg_set.update_len(config_len_train=n)
        _, self.occurence_map = self.ovg_set.update(self.sample_weight_map, self.occurence_map)
        return self.pose_generator(self.ovg_set.get_batch(0, n))

    @torch.no_grad()
    def render(self, poses: dict, rand: Optional[dict] = None, out: Optional[dict] = None) -> dict:
        """rand: optional {"hand_tex", "light", "bg_sel"} device tensors (the per-view draws of renderer.py:102-104);
        drawn on device from the pipeline's generator when omitted."""
        B = poses["final_obj_pose"].shape[0]
        if rand is None and self._render_rand is not None and self._render_rand[0] == B:
            rand, self._render_rand = self._render_rand[1], None   # drawn together with these poses by the fused launch
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
    def synthesise(self, n: int, out: Optional[dict] = None, prefetch: bool = False) -> dict:
        """One batch of n views: sample_poses -> render.

        prefetch: software-pipeline consecutive calls.  The draw + pose generator of the NEXT call (same n) are issued on a
        high-priority side stream BEFORE this call's rasteriser is launched, so the three small latency-bound launches
        (fused draw, prelude, LBS: ~80 us per 512 samples) run beside the ~0.3 ms rasteriser instead of in front of it.
        The Philox offsets advance in call order either way, so a prefetching sequence produces the same views as a plain
        one (tests/test_gpu_synthesis.py).  Poses drawn ahead use the weight map of the moment they were drawn:
        `drop_prefetch()` (called by whoever changes `sample_weight_map`) discards them."""
        main = torch.cuda.current_stream(self.device)
        ahead, self._ahead = getattr(self, "_ahead", None), None
        if ahead is not None and ahead[0] == n:
            _, poses, rand, done = ahead
            main.wait_event(done)
        else:
            poses = self.sample_poses(n)
            rand, self._render_rand = (self._render_rand[1] if self._render_rand is not None else None), None
        if prefetch:
            if getattr(self, "_pose_stream", None) is None:
                self._pose_stream = torch.cuda.Stream(self.device, priority=int(os.environ.get("AB_SYNTH_PRIO", "-1")))
            fence = torch.cuda.Event()
            fence.record(main)   # everything enqueued so far (a weight-map update included); NOT this call's rasteriser
            self._pose_stream.wait_event(fence)
            if self.fused_draw:
                # three library launches and no torch op: only the launches move to the side stream (lib.use_stream); the
                # buffers stay with the current stream's allocator pool -- their writers wait for `fence`, their reader
                # (the next call's rasteriser, on the current stream) waits for `done`
                with lib.use_stream(self._pose_stream):
                    nxt = self.sample_poses(n)
                nrand, self._render_rand = (self._render_rand[1] if self._render_rand is not None else None), None
            else:
                with torch.cuda.stream(self._pose_stream):
                    nxt = self.sample_poses(n)
                    nrand, self._render_rand = (self._render_rand[1] if self._render_rand is not None else None), None
                    for d in (nxt, nrand or {}):
                        for t in d.values():
                            if torch.is_tensor(t):
                                t.record_stream(main)   # allocated on the side stream, consumed on the main one
            done = torch.cuda.Event()
            done.record(self._pose_stream)
            self._ahead = (n, nxt, nrand, done)
        views = self.render(poses, rand=rand, out=out)
        views.update(obj_id=poses["obj_id"], persp_id=poses["persp_id"], grasp_id=poses["grasp_id"],
                     obj_pose=poses["final_obj_pose"], hand_verts=poses["final_hand_verts"],
                     joints=poses["final_joints"])
        return views

    def drop_prefetch(self):
        """Discards poses drawn ahead by `synthesise(prefetch=True)` (their Philox offsets stay consumed)."""
        self._ahead = None

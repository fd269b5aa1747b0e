"""ORACLE (test infrastructure): the synthesis path end to end on the host, built only from oracle/ pieces and the
synthetic assets -- CCV draw -> view -> grasp lookup -> pose generator (oracle/ccv.py) -> rasterise (oracle/raster.c,
OpenMP over views).  bench.py times `render` as the CPU baseline (`cpu_baseline`, `--impl reference`); the reference's
own renderer (pyrender 0.1.43 + EGL) and MANO layer (manotorch) are third-party, absent here and cannot run.
"""
import numpy as np

from . import ccv, raster


class CpuSynth:

    def __init__(self, assets_mod, obj_names=None, seed=0, n_hand_tex=51, n_bg=8, size=(256, 256),
                 cam=(217.5, 217.5, 128.0, 128.0), n_grasp=50):
        a = assets_mod
        self.obj_names = list(obj_names or a.HO3D_TRAIN_OBJS)
        self.model = a.make_synthetic_mano(seed)
        self.objects = a.make_synthetic_objects(self.obj_names, seed)
        self.grasps = a.make_synthetic_grasps(self.objects, n_grasp, seed)
        self.n_grasp = n_grasp
        W, H = size
        self.cfg = dict(width=W, height=H, fx=cam[0], fy=cam[1], cx=cam[2], cy=cam[3], znear=0.05, cull_backface=1,
                        ambient=0.8, diffuse=0.25)
        tex = a.make_hand_textures(n_hand_tex, seed, template=self.model["v_template"])
        rgba = lambda c: np.concatenate([c, np.full(c.shape[:-1] + (1,), 255, np.uint8)], -1)  # noqa: E731
        voff = np.cumsum([0] + [self.objects[n]["vertices"].shape[0] for n in self.obj_names]).astype(np.int32)
        foff = np.cumsum([0] + [self.objects[n]["faces"].shape[0] for n in self.obj_names]).astype(np.int32)
        self.scene = dict(
            hand_faces=self.model["f"].astype(np.int32), hand_cols=rgba(tex),
            obj_verts=np.concatenate([self.objects[n]["vertices"] for n in self.obj_names]).astype(np.float32),
            obj_vert_off=voff,
            obj_faces=np.concatenate([self.objects[n]["faces"] for n in self.obj_names]).astype(np.int32),
            obj_face_off=foff, obj_cols=rgba(np.concatenate([self.objects[n]["colors"] for n in self.obj_names])),
            bgs=a.make_backgrounds(n_bg, int(1.5 * H), int(1.5 * W), seed) if n_bg else None)
        self.rng = np.random.RandomState(seed + 101)
        self.weight_map = np.ones((len(self.obj_names), 288, n_grasp), np.float32)

    def sample(self, n):
        """-> per-view inputs of the rasteriser, drawn like SynthPipeline.sample_poses + draw_render_randoms."""
        rng = self.rng
        o, p, g = ccv.sample_ovg(self.weight_map, rng.rand(n))
        pose, shape, tsl = [], [], []
        for oi, gi in zip(o, g):
            hp, hs, ht = self.grasps[self.obj_names[oi]][gi]
            pose.append(hp), shape.append(np.zeros(10) if hs is None else hs), tsl.append(ht)
        views = [ccv.view_from_id(int(pi), 12, 24, (0.45, 0.55), *rng.rand(4).astype(np.float32)) for pi in p]
        persp, free, zoff = (np.stack([v[i] for v in views]) for i in range(3))
        out = ccv.pose_generator(self.model, np.stack(pose), np.stack(shape), np.stack(tsl), persp, free, zoff,
                                 rng.normal(0, 0.01, (n, 3)).astype(np.float32),
                                 rng.normal(0, 0.1, (n, 16)).astype(np.float32))
        W, H = self.cfg["width"], self.cfg["height"]
        sel = None
        if self.scene["bgs"] is not None:
            nb, bh, bw = self.scene["bgs"].shape[:3]
            ch = rng.randint(H, bh + 1, size=n)
            cw = (ch * W) // H
            sel = np.stack([rng.randint(nb, size=n), (rng.rand(n) * (bw - cw + 1)).astype(np.int64),
                            (rng.rand(n) * (bh - ch + 1)).astype(np.int64), cw, ch], 1).astype(np.int32)
        return dict(hand_verts=out["final_hand_verts"], obj_pose=out["final_obj_pose"].reshape(n, 16),
                    obj_id=o.astype(np.int32), hand_tex=rng.randint(self.scene["hand_cols"].shape[0], size=n),
                    light=rng.uniform(1, 5, size=n).astype(np.float32), bg_sel=sel)

    def render(self, inp, n_threads=None, out=None):
        return raster.render_batch(self.cfg, self.scene, inp["hand_verts"], inp["hand_tex"], inp["obj_id"],
                                   inp["obj_pose"], inp["light"], inp["bg_sel"], n_threads=n_threads, out=out)

"""Batched on-device renderer behind the reference's `Renderer` interface.

Mirrors anakin/utils/renderer.py: `Renderer(width, height, gpu_id)` (:46-55), `setup(cam_intr, cam_extr, obj_meshes,
hand_meshes, backgrounds, lights)` (:57-99), `__call__(obj_name, obj_pose, hand_verts, motion_blur=0) -> uint8[H,W,3]
BGR` (:101-123), random background crop rule `get_rand_bg` (:125-136).  The reference draws one view per call
through pyrender/OpenGL; here `render_batch` rasterises B views per call with ab_render_batch
(include/artiboost_b200.h) and `__call__` is a batch of one.  Meshes are duck-typed like trimesh: `.vertices`,
`.faces` and per-vertex colours in `.visual.vertex_colors` (or `.colors`).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, NamedTuple, Optional, Sequence, Union

import numpy as np
import torch

from . import DUMMY
from .. import lib
from .patches import PatchTable

PointLight = NamedTuple("PointLight", [("color", np.ndarray), ("intensity", float), ("pose", np.ndarray)])
DirectionalLight = NamedTuple("DirectionalLight", [("color", np.ndarray), ("intensity", float), ("pose", np.ndarray)])

# OpenCV-convention camera expressed in pyrender's frame (anakin/utils/misc.py:87-95)
PYRENDER_EXTRINSIC = np.array([[1.0, 0, 0, 0], [0, -1.0, 0, 0], [0, 0, -1.0, 0], [0, 0, 0, 1.0]])


def make_mesh(vertices, faces, colors=None):
    """Minimal trimesh-like container."""
    return SimpleNamespace(vertices=np.asarray(vertices), faces=np.asarray(faces),
                           visual=SimpleNamespace(vertex_colors=None if colors is None else np.asarray(colors)))


def _vertex_colors(mesh, n: int) -> np.ndarray:
    col = getattr(getattr(mesh, "visual", None), "vertex_colors", None)
    if col is None:
        col = getattr(mesh, "colors", None)
    if col is None:  # pyrender's default material base colour 0.3 (anakin/utils/frender_utils.py:153-157)
        col = np.full((n, 3), 77, np.uint8)
    col = np.asarray(col)
    if col.dtype != np.uint8:
        col = np.clip(np.round(col * 255.0 if col.max() <= 1.0 else col), 0, 255).astype(np.uint8)
    out = np.full((n, 4), 255, np.uint8)
    out[:, :col.shape[1]] = col[:, :4]
    return out


class Renderer:

    def __init__(self, width: int, height: int, gpu_id: int = 0, chunk: int = 512) -> None:
        if not torch.cuda.is_available():
            raise lib.AbError("Renderer needs a CUDA device: artiboost_b200 has no CPU path")
        lib.load()
        self.width, self.height = int(width), int(height)
        self.device = torch.device("cuda", int(gpu_id))
        self.chunk = int(chunk)
        self._ws = None
        self.rng = np.random  # the reference draws texture / light / background from np.random (renderer.py:102-104)

    # -------------------------------------------------------------------------------------------------- setup
    def setup(self, cam_intr: np.ndarray, cam_extr: np.ndarray, obj_meshes: Dict[str, object],
              hand_meshes: Sequence[object], backgrounds: Optional[Sequence[object]] = None,
              lights: Optional[List[Union[PointLight, DirectionalLight]]] = None, cull_backface: bool = True,
              diffuse: float = 0.25, znear: float = 0.05):
        cam_intr, cam_extr = np.asarray(cam_intr), np.asarray(cam_extr)
        assert cam_intr.shape == (3, 3) and cam_extr.shape == (4, 4), "camera parameter format error"
        if not np.allclose(cam_extr, PYRENDER_EXTRINSIC) and not np.allclose(cam_extr, np.eye(4)):
            raise NotImplementedError("only the OpenCV-convention camera at the origin (CONST.PYRENDER_EXTRINSIC)")
        dev = self.device
        self.obj_names = [k for k in obj_meshes.keys() if k != DUMMY]
        self.obj_index = {k: i for i, k in enumerate(self.obj_names)}
        verts, faces, cols, voff, foff = [], [], [], [0], [0]
        for k in self.obj_names:
            m = obj_meshes[k]
            v = np.asarray(m.vertices, np.float32)
            f = np.asarray(m.faces, np.int32)
            verts.append(v)
            faces.append(np.concatenate([f, np.zeros((f.shape[0], 1), np.int32)], axis=1))
            cols.append(_vertex_colors(m, v.shape[0]))
            voff.append(voff[-1] + v.shape[0])
            foff.append(foff[-1] + f.shape[0])
        cat = lambda xs, shape, dt: (np.concatenate(xs) if xs else np.zeros(shape, dt))  # noqa: E731
        # the C-ABI takes raw pointers: every upload is forced to C order first
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev).contiguous()  # noqa: E731
        self.obj_verts = up(cat(verts, (0, 3), np.float32))
        self.obj_faces = up(cat(faces, (0, 4), np.int32))
        self.obj_colors = up(cat(cols, (0, 4), np.uint8))
        self._voff = (C.c_int32 * len(voff))(*voff)
        self._foff = (C.c_int32 * len(foff))(*foff)
        if len(hand_meshes) == 0:
            raise ValueError("at least one hand mesh is required")
        hf = np.asarray(hand_meshes[0].faces, np.int32)
        n_hv = int(np.asarray(hand_meshes[0].vertices).shape[0])
        self.n_hand_verts, self.n_hand_tex = n_hv, len(hand_meshes)
        self.hand_faces = up(np.concatenate([hf, np.zeros((hf.shape[0], 1), np.int32)], axis=1))
        self.hand_colors = up(np.stack([_vertex_colors(m, n_hv) for m in hand_meshes]))
        # backgrounds: resized to 1.5x the render size like the reference (renderer.py:98-99), nearest neighbour
        self.backgrounds = None
        if backgrounds:
            bh, bw = int(1.5 * self.height), int(1.5 * self.width)
            bgs = []
            for b in backgrounds:
                a = np.asarray(b)[..., :3].astype(np.uint8)
                ys = (np.arange(bh) * a.shape[0]) // bh
                xs = (np.arange(bw) * a.shape[1]) // bw
                bgs.append(a[ys][:, xs])
            self.backgrounds = up(np.stack(bgs))
            # the kernel reads an RGBX copy: one aligned 32-bit load per background pixel instead of three byte loads
            self._bgs4 = torch.cat([self.backgrounds, torch.zeros_like(self.backgrounds[..., :1])], -1).contiguous()
        # mesh patches of the tile rasteriser (native host builder, once per scene)
        self.obj_patches = PatchTable([(np.asarray(obj_meshes[k].vertices, np.float32), np.asarray(obj_meshes[k].faces))
                                       for k in self.obj_names], dev)
        self.hand_patches = PatchTable([(np.asarray(hand_meshes[0].vertices, np.float32), hf)], dev, with_pos=False)
        self.scene = lib.SceneStruct(
            len(self.obj_names), self.obj_verts.data_ptr(), self.obj_faces.data_ptr(), self.obj_colors.data_ptr(),
            C.cast(self._voff, lib.c_i32_p), C.cast(self._foff, lib.c_i32_p), n_hv, hf.shape[0], self.n_hand_tex,
            self.hand_faces.data_ptr(), self.hand_colors.data_ptr(),
            None if self.backgrounds is None else self._bgs4.data_ptr(),
            0 if self.backgrounds is None else self.backgrounds.shape[0],
            0 if self.backgrounds is None else self.backgrounds.shape[1],
            0 if self.backgrounds is None else self.backgrounds.shape[2], 4,
            C.pointer(self.obj_patches.struct) if self.obj_names else None, C.pointer(self.hand_patches.struct))
        self.camera = lib.CameraStruct(self.width, self.height, float(cam_intr[0, 0]), float(cam_intr[1, 1]),
                                       float(cam_intr[0, 2]), float(cam_intr[1, 2]), float(znear), int(cull_backface),
                                       0.8, float(diffuse), 128, 128, 128)  # ambient 0.8, bg 0.5 (renderer.py:77)
        self.lights = lights
        self.set_chunk(self.chunk)

    def set_chunk(self, chunk: int) -> None:
        """(Re)allocates the workspace for groups of `chunk` views (the per-tile patch lists of one group)."""
        self.chunk = int(chunk)
        n = lib.load().ab_render_workspace_bytes(C.byref(self.scene), C.byref(self.camera), self.chunk)
        if n == 0:
            raise lib.AbError("ab_render_workspace_bytes: inconsistent scene / camera / chunk")
        self._ws = torch.empty(int(n) + 256, dtype=torch.uint8, device=self.device)
        self._ws_off = (-self._ws.data_ptr()) % 256

    # ----------------------------------------------------------------------------------------------- per call
    def get_rand_bg_sel(self, n: int) -> np.ndarray:
        """[n,5] int32 {bg id, x0, y0, crop_w, crop_h}: the crop rule of renderer.py:125-136."""
        sel = np.full((n, 5), -1, np.int32)
        if self.backgrounds is None:
            return sel
        nb, bh, bw = self.backgrounds.shape[:3]
        for i in range(n):
            bid = self.rng.randint(nb)
            if bh - self.height > bw - self.width:
                cw = self.rng.randint(self.width, bw + 1)
                ch = int(self.height / self.width * cw)
            else:
                ch = self.rng.randint(self.height, bh + 1)
                cw = int(self.width / self.height * ch)
            y0 = self.rng.randint(bh - ch + 1)
            x0 = self.rng.randint(bw - cw + 1)
            sel[i] = (bid, x0, y0, cw, ch)
        return sel

    @torch.no_grad()
    def render_batch(self, obj_ids: torch.Tensor, obj_poses: torch.Tensor, hand_verts: torch.Tensor,
                     hand_tex: Optional[torch.Tensor] = None, light: Optional[torch.Tensor] = None,
                     bg_sel: Optional[torch.Tensor] = None, out: Optional[dict] = None,
                     want=("rgba", "depth", "seg")) -> dict:
        """obj_ids i32[B] (<0: hand only), obj_poses f32[B,4,4], hand_verts f32[B,778,3], all on this renderer's
        device.  hand_tex i32[B] / light f32[B] / bg_sel i32[B,5] are the per-view random draws (drawn here like
        renderer.py:102-104 when omitted).  -> {"rgba": u8[B,H,W,4], "depth": f32[B,H,W], "seg": u8[B,H,W]}."""
        if self._ws is None:
            raise lib.AbError("Renderer.setup() has not been called")
        dev = self.device
        B = int(obj_ids.shape[0])
        for name, t in (("obj_ids", obj_ids), ("obj_poses", obj_poses), ("hand_verts", hand_verts)):
            lib.require_cuda(t, name)
        if tuple(hand_verts.shape) != (B, self.n_hand_verts, 3) or tuple(obj_poses.shape) != (B, 4, 4):
            raise ValueError("render_batch: bad input shapes")
        if hand_tex is None:
            hand_tex = torch.from_numpy(self.rng.randint(self.n_hand_tex, size=B).astype(np.int32)).to(dev)
        if light is None:
            light = torch.from_numpy(self.rng.uniform(1.0, 5.0, size=B).astype(np.float32)).to(dev)
        if bg_sel is None and self.backgrounds is not None:
            bg_sel = torch.from_numpy(self.get_rand_bg_sel(B)).to(dev)
        obj_ids = obj_ids.to(torch.int32).contiguous()
        obj_poses = obj_poses.float().contiguous()
        hand_verts = hand_verts.float().contiguous()
        hand_tex = hand_tex.to(torch.int32).contiguous()
        light = light.float().contiguous()
        bg_sel = None if bg_sel is None else bg_sel.to(torch.int32).contiguous()
        H, W = self.height, self.width
        out = {} if out is None else out
        if "rgba" in want and "rgba" not in out:
            out["rgba"] = torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev)
        if "depth" in want and "depth" not in out:
            out["depth"] = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        if "seg" in want and "seg" not in out:
            out["seg"] = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.load().ab_render_batch(
                C.byref(self.scene), C.byref(self.camera), B, self.chunk, lib.ptr(hand_verts), lib.ptr(hand_tex),
                lib.ptr(obj_ids), None, lib.ptr(obj_poses), lib.ptr(light), lib.ptr(bg_sel), lib.ptr(out.get("rgba")),
                lib.ptr(out.get("depth")), lib.ptr(out.get("seg")), C.c_void_p(self._ws.data_ptr() + self._ws_off),
                lib.stream_ptr(dev))
        lib.check(rc, "ab_render_batch")
        return out

    @torch.no_grad()
    def render_batch_host(self, obj_ids: torch.Tensor, obj_poses: torch.Tensor, hand_verts: torch.Tensor,
                          hand_tex: torch.Tensor, light: torch.Tensor, bg_sel: Optional[torch.Tensor], out: dict,
                          sub_batch: int = 64) -> dict:
        """Host-buffer entry point (what a CPU-side consumer such as the reference's RenderedDataset needs): all inputs
        and `out` = {"rgba","depth","seg"} are PINNED host tensors.  Views are processed in sub-batches; the
        device->host copy of sub-batch i runs on a second stream while sub-batch i+1 is rasterised."""
        dev = self.device
        B = int(obj_ids.shape[0])
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(dev)
            self._host_bufs = {}
        main = torch.cuda.current_stream(dev)
        d_in = [x.to(dev, non_blocking=True) if x is not None else None
                for x in (obj_ids, obj_poses, hand_verts, hand_tex, light, bg_sel)]
        H, W = self.height, self.width
        key = (sub_batch, H, W)
        if key not in self._host_bufs:  # two device staging sets, alternated
            self._host_bufs[key] = [
                {"rgba": torch.empty((sub_batch, H, W, 4), dtype=torch.uint8, device=dev),
                 "depth": torch.empty((sub_batch, H, W), dtype=torch.float32, device=dev),
                 "seg": torch.empty((sub_batch, H, W), dtype=torch.uint8, device=dev),
                 "free": torch.cuda.Event()} for _ in range(2)]
        bufs = self._host_bufs[key]
        keys = [k for k in ("rgba", "depth", "seg") if k in out]   # only what the caller asked for is rendered and copied
        for j, v0 in enumerate(range(0, B, sub_batch)):
            n = min(sub_batch, B - v0)
            buf = bufs[j & 1]
            main.wait_event(buf["free"])  # the copy that last read this staging set has finished
            sl = slice(v0, v0 + n)
            stage = {k: buf[k][:n] for k in keys}
            self.render_batch(d_in[0][sl], d_in[1][sl], d_in[2][sl], d_in[3][sl], d_in[4][sl],
                              None if d_in[5] is None else d_in[5][sl], out=stage, want=keys)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(done)
                for k in keys:
                    out[k][sl].copy_(stage[k], non_blocking=True)
                buf["free"].record(self._copy_stream)
        main.wait_stream(self._copy_stream)
        return out

    def __call__(self, obj_name: str, obj_pose: np.ndarray, hand_verts: np.ndarray, motion_blur: int = 0):
        """One view, host arrays in, uint8[H,W,3] BGR out (renderer.py:101-123)."""
        if motion_blur:
            raise NotImplementedError("motion_blur is never enabled by the reference's callers (render_infra.py:57)")
        dev = self.device
        oid = -1 if obj_name == DUMMY else self.obj_index[obj_name]  # KeyError on unknown names, like show_node
        pose = np.eye(4, dtype=np.float32) if oid < 0 else np.asarray(obj_pose, np.float32).reshape(4, 4)
        out = self.render_batch(torch.tensor([oid], dtype=torch.int32, device=dev),
                                torch.from_numpy(pose)[None].to(dev),
                                torch.from_numpy(np.asarray(hand_verts, np.float32))[None].to(dev), want=("rgba",))
        return out["rgba"][0, :, :, :3].flip(-1).cpu().numpy()

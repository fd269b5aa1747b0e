"""Mesh patches for the tile rasteriser: host binding of ab_build_patches_host (include/artiboost_b200.h, "Mesh patches").

The reference hands whole meshes to pyrender, which uploads one VBO per object and leaves culling / clipping to the GL
pipeline (anakin/utils/renderer.py:79-93).  Here every mesh is cut once, at Renderer.setup time, into patches of at most
32 faces over at most 32 vertices with a bounding sphere and a normal cone; face order is not changed (each face keeps its
original index as the primitive id of the z-test key, renderer.py:90-93 insertion order).  The builder is native host
code inside the C-ABI library and needs no device.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from .. import lib


def build_patches(vertices: np.ndarray, faces: np.ndarray) -> dict:
    """One mesh -> host arrays {"pos" f32[n,32,4], "vid" i32[n,32], "face" u32[n,32], "prim" i32[n,32],
    "bound" f32[n,12]} (layout: include/artiboost_b200.h)."""
    L = lib.load()
    v = np.ascontiguousarray(vertices, np.float32)
    f = np.ascontiguousarray(np.asarray(faces)[:, :3], np.int32)
    if v.ndim != 2 or v.shape[1] != 3 or f.ndim != 2 or f.shape[0] == 0:
        raise ValueError("build_patches: vertices [V,3] and a non-empty faces [F,3] expected")
    cap = int(L.ab_patch_capacity(f.shape[0]))
    pos = np.empty((cap, 32, 4), np.float32)
    vid = np.empty((cap, 32), np.int32)
    face = np.empty((cap, 32), np.uint32)
    prim = np.empty((cap, 32), np.int32)
    bound = np.empty((cap, 12), np.float32)
    n = C.c_int32(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = L.ab_build_patches_host(p(v), v.shape[0], p(f), f.shape[0], 3, cap, p(pos), p(vid), p(face), p(prim), p(bound),
                                 C.byref(n))
    lib.check(rc, "ab_build_patches_host")
    n = int(n.value)
    return {"pos": pos[:n].copy(), "vid": vid[:n].copy(), "face": face[:n].copy(), "prim": prim[:n].copy(),
            "bound": bound[:n].copy()}


class PatchTable:
    """Concatenated patches of several meshes, uploaded; `.struct` is the ab_patch_table the C-ABI takes (it points into
    tensors and a ctypes array owned by this object: keep it alive as long as the scene)."""

    def __init__(self, meshes: Sequence[Tuple[np.ndarray, np.ndarray]], device, with_pos: bool = True):
        import torch
        per_mesh: List[dict] = [build_patches(v, f) for v, f in meshes]
        off = [0]
        for t in per_mesh:
            off.append(off[-1] + t["vid"].shape[0])
        self.host = per_mesh
        self.n_mesh = len(per_mesh)
        self._off = (C.c_int32 * len(off))(*off)
        self.patch_off = off
        cat = lambda k, shape, dt: (np.concatenate([t[k] for t in per_mesh]) if per_mesh else np.zeros(shape, dt))  # noqa: E731
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device).contiguous()  # noqa: E731
        self.pos = up(cat("pos", (0, 32, 4), np.float32)) if with_pos else None
        self.vid = up(cat("vid", (0, 32), np.int32))
        self.face = up(cat("face", (0, 32), np.uint32).view(np.int32))   # same bits; torch has no uint32 arithmetic to need
        self.prim = up(cat("prim", (0, 32), np.int32))
        self.bound = up(cat("bound", (0, 12), np.float32))
        ptr = lambda t: None if t is None or t.numel() == 0 else t.data_ptr()  # noqa: E731
        self.struct = lib.PatchTableStruct(self.n_mesh, C.cast(self._off, lib.c_i32_p), ptr(self.pos), ptr(self.vid),
                                           ptr(self.face), ptr(self.prim), ptr(self.bound))

"""Loaders for the REAL assets the reference trains with, producing the structures the B200 pipeline consumes
(SURVEY.md 8(f).4).  None of these assets ship with this repo (licensed / large): MANO pickles, YCB meshes
(`data/YCB_models_process/<obj>/ds_textured.obj`, `data/DexYCB/models/<obj>/textured_simple.obj`), HTML hand textures
(`data/HTML_supp/html_XXX/hand.obj`), grasp tables (`assets/grasp_engine/ycb_grasp/<obj>.pkl`), HO3D corner file
(`assets/ho3d_corners.pkl`), background images (`assets/synth_bg`).  The reference reads them through trimesh
(anakin/artiboost/object_engine.py:30-91, hand_texture.py:5-12); trimesh is not a dependency here, so a small Wavefront
OBJ reader stands in for `trimesh.load(path, process=False)` (same vertex order, no merging).  pyrender textures the
meshes with their UV maps; the batched rasteriser shades per-vertex colours, so textures are sampled once per vertex
at load time.
"""
from __future__ import annotations

import os
import pickle
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np

CAM_EXTR = np.array([[1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]])  # object_engine.py:35-40


def load_obj(path: str) -> SimpleNamespace:
    """Wavefront OBJ -> namespace(vertices [V,3] f64, faces [F,3] i64, uv [V,2] | None, vertex_colors u8 [V,3] | None,
    texture_path | None).  Polygons are fan-triangulated; per-vertex UV = the first `vt` a face corner pairs with the
    vertex (process=False semantics: vertices are never duplicated or merged, so ids match the reference's)."""
    verts, cols, uvs, faces, face_uv = [], [], [], [], []
    mtl, tex = None, None
    with open(path, "r", errors="replace") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append([float(p[1]), float(p[2]), float(p[3])])
                if len(p) >= 7:
                    cols.append([float(p[4]), float(p[5]), float(p[6])])
            elif line.startswith("vt "):
                p = line.split()
                uvs.append([float(p[1]), float(p[2])])
            elif line.startswith("f "):
                corners = line.split()[1:]
                vi, ti = [], []
                for c in corners:
                    q = c.split("/")
                    vi.append(int(q[0]))
                    ti.append(int(q[1]) if len(q) > 1 and q[1] else 0)
                for k in range(1, len(vi) - 1):
                    faces.append([vi[0], vi[k], vi[k + 1]])
                    face_uv.append([ti[0], ti[k], ti[k + 1]])
            elif line.startswith("mtllib "):
                mtl = line.split(None, 1)[1].strip()
    V = np.asarray(verts, np.float64).reshape(-1, 3)
    F = np.asarray(faces, np.int64).reshape(-1, 3)
    F = np.where(F < 0, F + len(V) + 1, F) - 1  # negative = relative indices
    uv = None
    if uvs and face_uv:
        T = np.asarray(face_uv, np.int64).reshape(-1, 3)
        T = np.where(T < 0, T + len(uvs) + 1, T) - 1
        uv = np.zeros((len(V), 2))
        seen = np.zeros(len(V), bool)
        fv, ft = F.reshape(-1), T.reshape(-1)
        ok = ft >= 0
        # first occurrence wins: walk the corners in reverse so earlier ones overwrite later ones
        uv[fv[ok][::-1]] = np.asarray(uvs)[ft[ok][::-1]]
        seen[fv[ok]] = True
        if not seen.any():
            uv = None
    if mtl is not None:
        mp = os.path.join(os.path.dirname(path), mtl)
        if os.path.exists(mp):
            for line in open(mp, "r", errors="replace"):
                if line.strip().startswith("map_Kd"):
                    tex = os.path.join(os.path.dirname(path), line.split(None, 1)[1].strip())
    vc = None
    if len(cols) == len(V) and len(V):
        c = np.asarray(cols)
        vc = np.clip(np.round(c * 255.0 if c.max() <= 1.0 else c), 0, 255).astype(np.uint8)
    return SimpleNamespace(vertices=V, faces=F, uv=uv, vertex_colors=vc, texture_path=tex)


def sample_texture(uv: np.ndarray, image: np.ndarray) -> np.ndarray:
    """Nearest-texel lookup with OpenGL's convention (v = 0 at the bottom row, wrap = repeat) -> u8 [V,3]."""
    h, w = image.shape[:2]
    u = np.mod(uv[:, 0], 1.0)
    v = np.mod(uv[:, 1], 1.0)
    x = np.clip(np.floor(u * w).astype(np.int64), 0, w - 1)
    y = np.clip(np.floor((1.0 - v) * h).astype(np.int64), 0, h - 1)
    return np.ascontiguousarray(image[y, x, :3]).astype(np.uint8)


def load_textured_mesh(path: str) -> SimpleNamespace:
    """-> trimesh-like namespace(vertices, faces, visual.vertex_colors u8 [V,3]) as Renderer.setup consumes it."""
    m = load_obj(path)
    colors = m.vertex_colors
    if colors is None and m.uv is not None and m.texture_path and os.path.exists(m.texture_path):
        from PIL import Image
        colors = sample_texture(m.uv, np.asarray(Image.open(m.texture_path).convert("RGB")))
    return SimpleNamespace(vertices=m.vertices, faces=m.faces, visual=SimpleNamespace(vertex_colors=colors))


def center_vert_bbox(vertices, bbox_center=None, bbox_scale=None, scale=False):
    """anakin/utils/transform.py:621-631."""
    if bbox_center is None:
        bbox_center = (vertices.min(0) + vertices.max(0)) / 2
    vertices = vertices - bbox_center
    if scale:
        if bbox_scale is None:
            bbox_scale = np.linalg.norm(vertices, 2, 1).max()
        vertices = vertices / bbox_scale
    else:
        bbox_scale = 1
    return vertices, bbox_center, bbox_scale


def bounds_corners(vmin, vmax) -> np.ndarray:
    """trimesh.bounds.corners ordering: the 8 corners of the axis-aligned box [recalled]."""
    (x0, y0, z0), (x1, y1, z1) = vmin, vmax
    return np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0],
                     [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float64)


def _obj_record(mesh, verts_can, corners_can) -> dict:
    col = mesh.visual.vertex_colors
    if col is None:
        col = np.full((len(verts_can), 3), 77, np.uint8)  # pyrender's default base colour 0.3
    return {"vertices": verts_can.astype(np.float32), "faces": mesh.faces.astype(np.int32), "colors": col,
            "corners_can": np.asarray(corners_can, np.float32)}


def load_ho3d_objects(query_obj: List[str], obj_root="./data/YCB_models_process", corner_file="assets/ho3d_corners.pkl") -> Dict[str, dict]:
    """HO3DObjEngine (object_engine.py:30-65): flip into the camera convention, bbox-centre, corners from the pickle."""
    with open(corner_file, "rb") as f:
        obj_corners = pickle.load(f)
    out = {}
    for name in query_obj:
        mesh = load_textured_mesh(os.path.join(obj_root, name, "ds_textured.obj"))
        verts = CAM_EXTR.dot(mesh.vertices.transpose()).transpose()
        verts_can, center, scale = center_vert_bbox(verts, scale=False)
        corners = CAM_EXTR.dot(np.asarray(obj_corners[name]).transpose()).transpose()
        out[name] = _obj_record(mesh, verts_can, (corners - center) / scale)
    return out


def load_dexycb_objects(query_obj: List[str], obj_root="./data/DexYCB/models") -> Dict[str, dict]:
    """DexYCBObjEngine (object_engine.py:68-91)."""
    out = {}
    for name in query_obj:
        mesh = load_textured_mesh(os.path.join(obj_root, name, "textured_simple.obj"))
        verts_can, center, _ = center_vert_bbox(mesh.vertices, scale=False)
        corners = bounds_corners(mesh.vertices.min(0), mesh.vertices.max(0))
        out[name] = _obj_record(mesh, verts_can, corners - center)
    return out


def load_html_hands(root="data/HTML_supp") -> List[SimpleNamespace]:
    """HTMLHand.get_HTML_mesh (hand_texture.py:5-12): 51 textured hand meshes (html_003 is skipped)."""
    return [load_textured_mesh(os.path.join(root, f"html_{i + 1:03d}", "hand.obj")) for i in range(52) if i != 2]


def load_backgrounds(path="assets/synth_bg", size: Optional[int] = None) -> List[np.ndarray]:
    """Renderer.setup's `backgrounds` (renderer.py:60-67 reads every image of BGS_PATH) as RGB u8 arrays."""
    from PIL import Image
    out = []
    for fn in sorted(os.listdir(path)):
        if fn.lower().endswith((".jpg", ".jpeg", ".png")):
            im = Image.open(os.path.join(path, fn)).convert("RGB")
            if size:
                im = im.resize((size, size))
            out.append(np.asarray(im))
    return out
